timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --size-mb 256 --strong-mb 512 --steps 2 --warmup 1 > gpurun_out/bench_n2c.json 2> gpurun_out/bench_n2c.err; echo exit=$?; tail -3 gpurun_out/bench_n2c.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n2c.json'))
r=d['roofline']
print(d['value'], d['config']['parity_check'], d['strong']['parity_check'])
print({k:r[k] for k in ('kernel','achieved','frac','share_of_step','what')})
PY
