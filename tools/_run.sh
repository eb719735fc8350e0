python tools/dec_probe.py 1024 1200 mixed 2>&1 | tail -2
timeout 300 python -m pytest tests -m gpu -x -q -k "decompress or concatenated" 2>&1 | tail -2
for v in "" _sf20 _sf10; do
BZ2B200_LIB=$PWD/bzip2_opencl_b200/libbz2b200$v.so python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v.err; python -c "
import json; d=json.load(open('gpurun_out/bench_v$v.json')); k=d['roofline']['kernels_ms_per_step']; print('$v', d['value'], d['e2e']['value'], d['ms_per_step'], k['k_bwt_local'])"
BZ2B200_LIB=$PWD/bzip2_opencl_b200/libbz2b200$v.so python tools/perf_probe.py rep64 2>&1 | grep -v roundtrip
done
python bench.py --steps 3 --warmup 2 --no-cpu-baseline --parallel 1200 > gpurun_out/bench_p1200.json 2> gpurun_out/bench_v.err; python -c "
import json; d=json.load(open('gpurun_out/bench_p1200.json')); print('P1200', d['value'], d['e2e']['value'], d['ms_per_step'])"
