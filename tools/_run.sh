timeout 80 python -m pytest tests -m gpu -x -q -k "cli_front" 2>&1 | tail -3
