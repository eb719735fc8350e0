timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_v9.json 2> gpurun_out/bench_v9.err; echo bench_exit=$?; tail -2 gpurun_out/bench_v9.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r01_v9.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_v9_list.log 2>&1; echo ncu_list_exit=$?
