timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -x -q 2>&1 | tail -8
