python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python -m pytest tests -m gpu -x -q -k "lanes or empty or stage_dumps or streaming or output_too_small" 2>&1 | tail -2
