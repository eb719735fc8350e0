#!/usr/bin/env python3
"""Host-buffer (pinned) compress call with the library's timeline trace on (development aid): BZ2B200_TRACE=1 python tools/e2e_probe.py [MiB] [P]"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bzip2_opencl_b200 import api, corpus

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
par = int(sys.argv[2]) if len(sys.argv) > 2 else 600
data = corpus.mixed_corpus(mib << 20)
n = len(data)
host = torch.frombuffer(bytearray(data), dtype=torch.uint8).pin_memory()
c = api.Compressor(9, par, 100000)
cap = c.max_output_bytes(n) + 64
hout = torch.empty(cap, dtype=torch.uint8).pin_memory()
d_in = host.cuda()
d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
for mode in ("device", "host"):
    for r in range(3):
        torch.cuda.synchronize()
        if r == 2:
            print(f"---- {mode} buffers, traced call", file=sys.stderr, flush=True)
            os.environ["BZ2B200_TRACE_NOW"] = "1"
        t = time.perf_counter()
        if mode == "device":
            c.compress_blocks_ptr(d_in.data_ptr(), n, d_out.data_ptr(), cap)
        else:
            c.compress_blocks_ptr(host.data_ptr(), n, hout.data_ptr(), cap)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        st = c.stats()
    print(f"{mode}: wall {dt*1e3:.1f} ms  dev {st['ms_total']:.1f}  h2d {st['ms_h2d']:.1f} d2h {st['ms_d2h']:.2f} walk {st['ms_walk']:.1f}", flush=True)
