python bench.py > gpurun_out/bench26.json 2> gpurun_out/bench26.err; echo bench_exit=$?; cat gpurun_out/bench26.json | cut -c1-1500
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu7.log 2>&1; echo ncu_exit=$?
ncu --set full --clock-control none --import-source on -k regex:'k_bwt_local' --launch-skip 4 --launch-count 2 -o gpurun_out/prof_r01_local_v2 -f python tools/ncu_probe.py 256 320 > gpurun_out/ncu8.log 2>&1; echo ncu_exit=$?
