for sg in "16,128" "16,32" "32,32" "16,64"; do
BZ2B200_SEG_MB=$sg python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_sg.json 2> gpurun_out/bench_sg.err; python -c "
import json; d=json.load(open('gpurun_out/bench_sg.json')); print('$sg', d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'])"
done
BZ2B200_TRACE=1 python tools/e2e_probe.py 1024 600 > gpurun_out/e2e_probe2.log 2> gpurun_out/e2e_trace2.log; cat gpurun_out/e2e_probe2.log
timeout 600 python -m pytest tests -m gpu -x -q -k "segments or lanes or streaming or boundary" 2>&1 | tail -3
