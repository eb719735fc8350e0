timeout 1500 python -m pytest tests -m gpu -x -q -k "segments or alternative or stage_dumps or streaming" > gpurun_out/pytest_gpu8.log 2>&1; echo pytest_exit=$?; tail -5 gpurun_out/pytest_gpu8.log
