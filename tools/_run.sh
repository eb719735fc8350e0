timeout 900 python tools/gpu_debug.py > gpurun_out/debug38.log 2>&1; echo exit=$?; tail -1 gpurun_out/debug38.log; grep -v "^   -> OK\|^==" gpurun_out/debug38.log | head
timeout 900 python -m pytest tests -m gpu -x -q -k "boundary or golden or streaming or ragged or sharded or segments or alternative or stage" 2>&1 | tail -3
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench38.json 2> gpurun_out/bench38.err; echo bench_exit=$?; python -c "
import json; d=json.load(open('gpurun_out/bench38.json')); print(d['value'], d['e2e']['value'], d['ms_per_step']); k=d['roofline']['kernels_ms_per_step']; print({n:v for n,v in k.items() if 'rle' in n or 'crc' in n}); print(d['roofline']['stage_ms_last_step'])"
