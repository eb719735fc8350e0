nvidia-smi --query-gpu=name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 4 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo exit=$?; tail -5 gpurun_out/bench_n4.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_n4.json'))
    print(json.dumps(d)[:3000])
except Exception as e: print("no json", e)
PY
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n4.json'))
for r,p in enumerate(d['config']['phases_ms_per_rank']): print('weak rank',r,p)
for r,p in enumerate(d['strong']['phases_ms_per_rank']): print('strong rank',r,p)
PY
