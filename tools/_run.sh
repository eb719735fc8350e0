timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu10.log 2>&1; echo pytest_exit=$?; tail -4 gpurun_out/pytest_gpu10.log
python bench.py > gpurun_out/bench33.json 2> gpurun_out/bench33.err; echo bench_exit=$?; python -c "
import json; d=json.load(open('gpurun_out/bench33.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['decompress'], d['cpu_baseline'], d['roofline']['kernel'], d['roofline']['frac'])"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
