timeout 900 python tools/gpu_debug.py > gpurun_out/debug35.log 2>&1; echo exit=$?; tail -1 gpurun_out/debug35.log
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench34.json 2> gpurun_out/bench34.err; echo bench_exit=$?; python -c "
import json; d=json.load(open('gpurun_out/bench34.json')); print(d['value'], d['e2e']['value'], d['ms_per_step']); print(list(d['roofline']['kernels_ms_per_step'].items())[:5])"
python tools/perf_probe.py rep64 2>&1 | tail -3
