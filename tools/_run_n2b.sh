for v in "BZ2B200_CHAIN_OVERLAP=0 BZ2B200_WALK_SMEM=0" "BZ2B200_CHAIN_OVERLAP=1 BZ2B200_WALK_SMEM=0" "BZ2B200_CHAIN_OVERLAP=1 BZ2B200_WALK_SMEM=190"; do
env $v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --strong-mb 0 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
python - "$v" <<'PY'
import json,sys
d=json.load(open('gpurun_out/bench_n2.json'))
p=d['config']['phases_ms_per_rank']
print(sys.argv[1], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'walked', [x.get('walked_at') for x in p], 'compress', [x['compress'] for x in p])
PY
done
