timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_v10.json 2> gpurun_out/bench_v10.err; echo bench_exit=$?; tail -2 gpurun_out/bench_v10.err; python -c "
import json; d=json.load(open('gpurun_out/bench_v10.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic'], d['cpu_baseline']['value'], d['decompress']['value'])"
