python tools/ncu_probe.py 256 320 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_radix_onesweep|k_bwt_regroup|k_bwt_local|k_rle_walk_prefix|k_rle_emit|k_rle_btile_seg|k_rle_tile_seg" -s 0 -c 40 -f -o /tmp/r02_full python tools/ncu_probe.py 256 320 > /dev/null 2>&1
python tools/ncu_summary.py /tmp/r02_full.ncu-rep > gpurun_out/r02_ncu_full_summary2.md 2>&1
python tools/ncu_hot.py /tmp/r02_full.ncu-rep k_rle_walk_prefix 0 1.5 > gpurun_out/r02_ncu_hot_walk.txt 2>&1
python tools/ncu_hot.py /tmp/r02_full.ncu-rep k_radix_onesweep 6 1.5 > gpurun_out/r02_ncu_hot_onesweep.txt 2>&1
ls -la /tmp/r02_full.ncu-rep; du -sh gpurun_out
