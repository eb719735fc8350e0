timeout 600 python tools/dec_debug.py > gpurun_out/dec4.log 2>&1; echo exit=$?; grep -v " OK " gpurun_out/dec4.log | tail -8
python tools/dec_probe.py 1024 1200 mixed 2>&1 | tail -3
python tools/dec_probe.py 512 1200 rand 2>&1 | tail -2
