python tools/dec_probe.py 1024 1200 mixed 2>&1 | tail -2
timeout 300 python -m pytest tests -m gpu -x -q -k "decompress or concatenated or input_stream" 2>&1 | tail -2
