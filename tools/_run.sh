timeout 900 python tools/gpu_debug.py > gpurun_out/debug45.log 2>&1; echo exit=$?; tail -1 gpurun_out/debug45.log; grep -v "^   -> OK\|^==" gpurun_out/debug45.log | head -5
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench45.json 2> gpurun_out/bench45.err; echo "exit=$?"; tail -3 gpurun_out/bench45.err; python -c "
import json; d=json.load(open('gpurun_out/bench45.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['rho_prefix_doubling']); k=d['roofline']['kernels_ms_per_step']; print(sorted(k.items(), key=lambda x:-x[1])[:12])"
python tools/perf_probe.py text64 rand256 2>&1 | grep -v roundtrip
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
