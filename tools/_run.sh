timeout 140 python -m pytest tests -m gpu -x -q -k "input_stream_facade or cli_front" 2>&1 | tail -3
