timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -x -q 2>&1 | tail -5
python tools/scan_probe.py 2>&1 | tail -8
