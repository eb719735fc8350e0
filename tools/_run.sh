python bench.py > gpurun_out/bench_v13.json 2> gpurun_out/bench_v13.err; echo bench_exit=$?; tail -2 gpurun_out/bench_v13.err; python -c "
import json; d=json.load(open('gpurun_out/bench_v13.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['cpu_baseline'], d['decompress']['value'], d['config']['parity_check'])"
