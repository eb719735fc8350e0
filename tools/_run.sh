python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench25.json 2> gpurun_out/bench25.err; echo bench_exit=$?; python -c "
import json; d=json.load(open('gpurun_out/bench25.json')); print(d['value'], d['e2e'], d['ms_per_step']); print(d['roofline']['stage_ms_last_step'])"
