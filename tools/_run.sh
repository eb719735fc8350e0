timeout 900 python tools/gpu_debug.py > gpurun_out/debug47.log 2>&1; echo exit=$?; tail -1 gpurun_out/debug47.log; grep -v "^   -> OK\|^==" gpurun_out/debug47.log | head -5
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench47.json 2> gpurun_out/bench47.err; echo "exit=$?"; tail -3 gpurun_out/bench47.err; python -c "
import json; d=json.load(open('gpurun_out/bench47.json')); print(d['value'], d['e2e']['value'], d['ms_per_step']); k=d['roofline']['kernels_ms_per_step']; print(k['k_mtf_ranks'], k['k_bwt_local'], k['k_radix_onesweep'])"
timeout 900 python -m pytest tests -m gpu -x -q -k "golden or stage or ragged or alphabet or config4" 2>&1 | tail -3
