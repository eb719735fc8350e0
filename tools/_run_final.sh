# final confirmation of the tree (development aid): GPU suite, smoke, one bench line
nvidia-smi --query-gpu=name --format=csv,noheader | head -1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; echo bench_exit=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_final2.json')); print(d['value'], d['e2e']['value'], d['config']['parity_check'], d.get('e2e_facade'))"
