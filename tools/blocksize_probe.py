#!/usr/bin/env python3
"""BASELINE config 5 shape on one GPU: block sizes -1..-9 on 1 GiB of the C5 text, device-resident, with the batch size the facade picks
(bz2b200_recommended_blocks: the same BYTES in flight at every block size) and with the reference CLI's fixed -p 600."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bzip2_opencl_b200 import api, corpus
n = 1 << 30
data = corpus.word_text(n, 5, markov=True)
d_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
lib = api.load_library()
res = {}
for mult in (1, 2, 3, 5, 7, 9):
    row = {}
    for label, par in (("recommended", lib.bz2b200_recommended_blocks(0, 100000, mult)), ("p600", 600)):
        c = api.Compressor(mult, par, 100000)
        cap = c.max_output_bytes(n) + 64
        out = torch.empty(cap, dtype=torch.uint8, device="cuda")
        best = None
        for _ in range(3):
            torch.cuda.synchronize(); t = time.perf_counter()
            r = c.compress_blocks_ptr(d_in.data_ptr(), n, out.data_ptr(), cap)
            torch.cuda.synchronize(); dt = time.perf_counter() - t
            best = dt if best is None else min(best, dt)
        row[label] = {"blocks_in_flight": par, "MB/s": round(n / 1e6 / best, 1), "blocks": len(r["crcs"])}
        c.close(); del out
        torch.cuda.empty_cache()
    res[f"-{mult}"] = row
print(json.dumps(res, indent=1))
