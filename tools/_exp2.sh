for b in 1 4 8 16; do
BZ2B200_LIB=$PWD/bzip2_opencl_b200/libbz2b200_mb$b.so python bench.py --no-extras --no-cpu-baseline --steps 4 --warmup 2 > gpurun_out/exp_mb$b.json 2>/dev/null
python - $b <<'PY'
import json,sys
d=json.load(open(f'gpurun_out/exp_mb{sys.argv[1]}.json'))
k=d['roofline']['kernels_ms_per_step']
print('mbatch', sys.argv[1], d['value'], d['e2e']['value'], 'onesweep', k['k_radix_onesweep'], 'local', k['k_bwt_local'])
PY
done
