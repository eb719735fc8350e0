python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_pre.json 2> /dev/null || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_launches_bench_1GiB.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_list.json 2> /dev/null
python tools/ncu_probe.py 256 320 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_radix_onesweep|k_bwt_local|k_bwt_regroup|k_bwt_gather_local|k_mtf_ranks|k_rle_emit|k_rle_btile_seg|k_rle_walk_prefix|k_rle_tile_seg" -s 40 -c 28 -f -o /tmp/r02_full python tools/ncu_probe.py 256 320 > /dev/null 2>&1
python tools/ncu_summary.py /tmp/r02_full.ncu-rep > gpurun_out/r02_ncu_full_summary.md 2>&1
for k in k_radix_onesweep k_bwt_local k_bwt_regroup k_rle_walk_prefix; do python tools/ncu_hot.py /tmp/r02_full.ncu-rep $k 0 1.5 > gpurun_out/r02_ncu_hot_$k.txt 2>&1; done
ls -la /tmp/r02_full.ncu-rep; du -sh gpurun_out
