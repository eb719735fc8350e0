timeout 1500 python tools/run_configs.py --out gpurun_out/r01_configs_full_size_v8.json > gpurun_out/configs_v8.log 2>&1; echo exit=$?; tail -30 gpurun_out/configs_v8.log
