python bench.py > gpurun_out/bench_v12.json 2> gpurun_out/bench_v12.err; echo bench_exit=$?; python -c "
import json; d=json.load(open('gpurun_out/bench_v12.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['cpu_baseline']['value'], d['decompress']['value'], d['config']['parity_check'])"
