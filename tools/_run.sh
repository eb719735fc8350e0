timeout 1500 python -m pytest tests -m gpu -x -q -k "decompress or input_stream" > gpurun_out/pytest_gpu9.log 2>&1; echo pytest_exit=$?; tail -15 gpurun_out/pytest_gpu9.log
