timeout 900 python tools/gpu_debug.py > gpurun_out/debug34.log 2>&1; echo exit=$?; tail -1 gpurun_out/debug34.log; grep -v "^   -> OK\|^==" gpurun_out/debug34.log | head
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench32.json 2> gpurun_out/bench32.err; echo bench_exit=$?; python -c "
import json; d=json.load(open('gpurun_out/bench32.json')); print(d['value'], d['e2e']['value'], d['ms_per_step']); print({k:v for k,v in d['roofline']['kernels_ms_per_step'].items() if 'mtf' in k or 'rle2' in k}); print(d['roofline']['stage_ms_last_step'])"
