timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -x -q -k "decompress or facade or input_stream or alternative or concatenated or cli" 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
