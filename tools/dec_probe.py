#!/usr/bin/env python3
"""Decompress throughput + per-kernel times on a GPU box (development aid): python tools/dec_probe.py [MiB] [P] [kind]"""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bzip2_opencl_b200 import api, corpus
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 256
par = int(sys.argv[2]) if len(sys.argv) > 2 else 600
kind = sys.argv[3] if len(sys.argv) > 3 else "mixed"
data = {"mixed": lambda: corpus.mixed_corpus(mb << 20, workers=16), "rand": lambda: corpus.rand_bytes(mb << 20, 4),
        "text": lambda: corpus.word_text(mb << 20, 101, markov=True)}[kind]()
c = api.Compressor(9, par, 100000)
s = c.compress(data)
d_in = torch.frombuffer(bytearray(s), dtype=torch.uint8).cuda()
d_out = torch.empty(len(data) + 64, dtype=torch.uint8, device="cuda")
for rep in range(3):
    if rep == 2: c.set_profiling(True)
    torch.cuda.synchronize(); t = time.perf_counter()
    n = c.decompress_ptr(d_in.data_ptr(), len(s), d_out.data_ptr(), len(data) + 64)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"{kind} {mb} MiB: {len(s)} -> {n} bytes in {dt*1e3:.1f} ms = {n/dt/1e6:.1f} MB/s (dev {c.stats()['ms_total']:.1f} ms, blocks {c.stats()['n_blocks']})", flush=True)
assert bytes(d_out[:n].cpu().numpy()) == data
kt = c.kernel_times()
print({k: round(v[0], 2) for k, v in sorted(kt.items(), key=lambda kv: -kv[1][0]) if v[1]})
