python tools/ncu_probe.py 256 320 > gpurun_out/probe.log 2>&1 || { tail -3 gpurun_out/probe.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_bwt_local -s 2 -c 4 -f -o gpurun_out/r02_local python tools/ncu_probe.py 256 320 > gpurun_out/ncu_local.log 2>&1
tail -2 gpurun_out/ncu_local.log
ls -la gpurun_out/r02_local.ncu-rep
