# scratch command file for `gpurun -- 'bash tools/_run.sh'` (development aid): the round's usual check
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo bench_exit=$?; tail -2 gpurun_out/bench.err
