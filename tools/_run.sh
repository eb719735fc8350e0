timeout 600 python tools/dec_debug.py > gpurun_out/dec2.log 2>&1; echo exit=$?; grep -v " OK " gpurun_out/dec2.log | tail -12
python tools/dec_probe.py 256 600 mixed 2>&1 | tail -4
