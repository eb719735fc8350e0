for cfg in "6 48 600" "6 48 1200" "8 32 1200" "6 96 1200"; do set -- $cfg
BZ2B200_WALK_GROWTH=$1 BZ2B200_WALK_FIRST=$2 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --parallel $3 > gpurun_out/bench_w.json 2> gpurun_out/bench_w.err; python -c "
import json; d=json.load(open('gpurun_out/bench_w.json')); print('$cfg', d['value'], d['e2e']['value'], d['ms_per_step'])"
done
