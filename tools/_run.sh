# scratch command file for `gpurun -- 'bash tools/_run.sh'` (development aid): the round's usual check
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -8
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo bench_exit=$?; tail -1 gpurun_out/bench.err
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo ref_exit=$?
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['e2e']['value'], d['config']['parity_check'])"
