# scratch command file for `gpurun -- 'bash tools/_run.sh'` (development aid): the round's usual check
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -8
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -15
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5
python tools/scan_probe.py 2>&1 | tail -8
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo bench_exit=$?; tail -2 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print(d['value'], d['e2e']['value'], d['roofline']['kernels_ms_per_step'].get('k_rle_walk'), d['roofline']['stage_ms_last_step'])
for k in ('config3_repetitive','strong','e2e_facade','decompress'):
    print(k, json.dumps(d.get(k))[:1500])
PY
