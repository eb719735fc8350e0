python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench23.json 2> gpurun_out/bench23.err; echo bench_exit=$?; python -c "
import json; d=json.load(open('gpurun_out/bench23.json')); print(d['value'], d['e2e']['value'], d['ms_per_step']); print(d['roofline']['kernels_ms_per_step']); print(d['roofline']['stage_ms_last_step'])"
