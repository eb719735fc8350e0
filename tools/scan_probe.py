"""standalone boundary scan (summary tree + chained walk) on device-resident input: ms per GiB, per corpus"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bzip2_opencl_b200 import api, corpus
import importlib.util
spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
c = api.Compressor(9, 600, 100000)
def probe(name, t):
    n = t.numel()
    best = None
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        starts = c.scan_blocks(ptr=t.data_ptr(), nbytes=n)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    st = c.stats()
    print(f"{name}: {n >> 20} MiB, {len(starts)} blocks, scan {best * 1e3:.2f} ms = {best * 1e3 / (n / 2**30):.2f} ms/GiB, {best * 1e6 / max(1, len(starts)):.2f} us/block, ms_walk(events) {st['ms_walk']:.2f}")
data = corpus.mixed_corpus(1 << 30, workers=16)
probe("mixed", torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda())
third = len(data) // 64
for k, nm in ((0, "text"), (1, "code"), (2, "binary")):
    seg = data[k * (16 << 20):(k + 1) * (16 << 20)] * 16
    probe(nm, torch.frombuffer(bytearray(seg), dtype=torch.uint8).cuda())
probe("random", torch.randint(0, 256, (1 << 30,), dtype=torch.uint8, device="cuda"))
probe("zeros", torch.zeros(1 << 30, dtype=torch.uint8, device="cuda"))
