timeout 2400 python tools/run_configs.py --out gpurun_out/configs_full.json > gpurun_out/configs_full.log 2>&1; echo exit=$?; tail -5 gpurun_out/configs_full.log | cut -c1-600
