timeout 900 python -m pytest tests -m gpu -x -q -k "cli or input_stream or facade" > gpurun_out/pytest_gpu11.log 2>&1; echo pytest_exit=$?; tail -6 gpurun_out/pytest_gpu11.log
