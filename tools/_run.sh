# scratch command file for `gpurun -- 'bash tools/_run.sh'` (development aid): the round's usual check
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -8
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo bench_exit=$?; tail -2 gpurun_out/bench.err
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo ref_exit=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print(d['value'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'], d['cpu_baseline'])
for k in ('config3_repetitive','strong','e2e_facade'):
    print(k, json.dumps(d.get(k))[:900])
print('decompress', json.dumps({k:v for k,v in d['decompress'].items() if k!='roofline'})[:900])
print(open('gpurun_out/bench_ref.json').read()[:600])
PY
