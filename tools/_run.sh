python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench51.json 2> gpurun_out/bench51.err; echo "exit=$?"; tail -3 gpurun_out/bench51.err; python -c "
import json; d=json.load(open('gpurun_out/bench51.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'])"
timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
