python tools/kprobe.py rep 256 300 2>&1 | tail -2
