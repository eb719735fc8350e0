python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_v8.json 2> gpurun_out/bench_v8.err; echo bench_exit=$?; tail -2 gpurun_out/bench_v8.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_v8_ref.json 2> gpurun_out/bench_v8_ref.err; echo ref_exit=$?; cat gpurun_out/bench_v8_ref.json | cut -c1-400
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r01_v8.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_v8_list.log 2>&1; echo ncu_list_exit=$?
ncu --set full --clock-control none --import-source on -k regex:"k_radix_onesweep|k_bwt_local|k_bwt_regroup|k_bwt_count_packed" -c 14 -f -o gpurun_out/prof_r01_v8_top python tools/ncu_probe.py 256 320 > gpurun_out/ncu_v8_full.log 2>&1; echo ncu_full_exit=$?; tail -2 gpurun_out/ncu_v8_full.log
ls -la gpurun_out/*v8*
