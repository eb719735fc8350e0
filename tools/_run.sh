python tools/perf_probe.py mixed256 rand256 rep64 > gpurun_out/probe4.log 2>&1; cat gpurun_out/probe4.log
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench19.json 2> gpurun_out/bench19.err; echo bench_exit=$?; python -c "
import json; d=json.load(open('gpurun_out/bench19.json')); print(d['value'], d['e2e']['value'], d['ms_per_step']); print(d['roofline']['kernels_ms_per_step'])"
