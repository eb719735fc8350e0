timeout 900 python tools/gpu_debug.py > gpurun_out/debug49.log 2>&1; echo exit=$?; tail -1 gpurun_out/debug49.log; grep -v "^   -> OK\|^==" gpurun_out/debug49.log | head -5
python bench.py --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/bench49.json 2> gpurun_out/bench49.err; echo "exit=$?"; tail -3 gpurun_out/bench49.err; python -c "
import json; d=json.load(open('gpurun_out/bench49.json')); print(d['value'], d['e2e']['value'], d['ms_per_step']); k=d['roofline']['kernels_ms_per_step']; print(k['k_rle2_emit'], k['k_mtf_ranks'], k['k_bwt_local'], k['k_radix_onesweep'])"
timeout 900 python -m pytest tests -m gpu -x -q -k "golden or stage or ragged or alphabet or config4 or config3" 2>&1 | tail -3
