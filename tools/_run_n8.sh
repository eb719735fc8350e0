nvidia-smi --query-gpu=name --format=csv,noheader | wc -l
nvidia-smi topo -m 2>/dev/null | head -12
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 4 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo exit=$?; tail -3 gpurun_out/bench_n8.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_n8.json'))
    print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['parity_check'], d['config']['blocks_per_rank'])
    s=d['strong']; print('strong', s['value'], s['ms_per_step'], s['e2e'], s['parity_check'], s['blocks_per_rank'])
except Exception as e: print("no json", e)
PY
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n8.json'))
for r,p in enumerate(d['config']['phases_ms_per_rank']): print('weak rank',r,p)
PY
