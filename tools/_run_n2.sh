nvidia-smi --query-gpu=name --format=csv,noheader | head -8
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --size-mb 512 --strong-mb 2048 --steps 2 --warmup 1 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo exit=$?; tail -5 gpurun_out/bench_n2.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_n2.json'))
    print(json.dumps(d)[:3000])
except Exception as e: print("no json", e)
PY
