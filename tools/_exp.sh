run() { # name, env..., args
  name=$1; shift
  env "$@" python bench.py --no-extras --no-cpu-baseline --steps 4 --warmup 2 $ARGS > gpurun_out/exp_$name.json 2> gpurun_out/exp_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.load(open(f'gpurun_out/exp_{n}.json'))
    k=d['roofline']['kernels_ms_per_step']
    top=sorted(k.items(), key=lambda kv:-kv[1])[:10]
    print(f"{n:14s} value {d['value']:9.1f} e2e {d['e2e']['value']:9.1f}  ", ' '.join(f"{a[2:]}={b:.2f}" for a,b in top))
except Exception as e: print(n, "FAILED", e)
PY
}
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -x -q 2>&1 | tail -5
ARGS=""
run default X=1
ARGS="--size-mb 64"
run c1 X=1
run c1_first BZ2B200_WALK_FIRST=600
