timeout 900 python tools/gpu_debug.py > gpurun_out/debug46.log 2>&1; echo exit=$?; tail -1 gpurun_out/debug46.log; grep -v "^   -> OK\|^==" gpurun_out/debug46.log | head -5
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench46.json 2> gpurun_out/bench46.err; echo "exit=$?"; tail -3 gpurun_out/bench46.err; python -c "
import json; d=json.load(open('gpurun_out/bench46.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['rho_prefix_doubling']); k=d['roofline']['kernels_ms_per_step']; print(sorted(k.items(), key=lambda x:-x[1])[:12]); print(d['roofline']['traffic'])"
timeout 900 python -m pytest tests -m gpu -x -q -k "golden or boundary or ragged or alternative or lanes or config3 or config5" 2>&1 | tail -3
