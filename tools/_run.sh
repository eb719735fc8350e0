timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
ncu --set full --clock-control none --import-source on -k regex:"k_bwt_local|k_mtf_ranks|k_bwt_gather_local|k_rle2_emit" -c 8 -f -o gpurun_out/prof_r01_v8_local_mtf python tools/ncu_probe.py 256 320 > gpurun_out/ncu_v8_full2.log 2>&1; echo ncu_full_exit=$?; tail -2 gpurun_out/ncu_v8_full2.log
