timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "golden or stage or alphabet or ragged or config1 or lanes" 2>&1 | tail -3
for b in 8 10; do
BZ2B200_R0_BITS=$b python bench.py --no-extras --no-cpu-baseline --steps 4 --warmup 2 > gpurun_out/exp_r0b$b.json 2>/dev/null
python - $b <<'PY'
import json,sys
d=json.load(open(f'gpurun_out/exp_r0b{sys.argv[1]}.json'))
k=d['roofline']['kernels_ms_per_step']
print('r0_bits', sys.argv[1], d['value'], d['e2e']['value'], 'onesweep', k['k_radix_onesweep'], 'keys0', k['k_bwt_keys0'], 'regroup', k['k_bwt_regroup'], 'bases', k['k_radix_bases'])
PY
done
