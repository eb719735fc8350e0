for v in "" _sf16 _wm128 _wm256 _sf16wm128 _sf8; do
BZ2B200_LIB=$PWD/bzip2_opencl_b200/libbz2b200$v.so python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v.err; python -c "
import json; d=json.load(open('gpurun_out/bench_v$v.json')); k=d['roofline']['kernels_ms_per_step']; print('$v', d['value'], d['e2e']['value'], d['ms_per_step'], k['k_bwt_local'], d['config']['parity_check'])"
done
