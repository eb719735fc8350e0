#!/usr/bin/env python3
"""One compress call over a mixed corpus (development aid for ncu captures): python tools/ncu_probe.py [MiB] [P] [kind]"""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bzip2_opencl_b200 import api, corpus

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 128
par = int(sys.argv[2]) if len(sys.argv) > 2 else 160
kind = sys.argv[3] if len(sys.argv) > 3 else "mixed"
data = {"mixed": lambda: corpus.mixed_corpus(mb << 20), "rand": lambda: corpus.rand_bytes(mb << 20, 4),
        "text": lambda: corpus.word_text(mb << 20, 101, markov=True), "rep": lambda: corpus.repetitive_corpus(mb << 20)}[kind]()
c = api.Compressor(9, par, 100000)
d_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
cap = c.max_output_bytes(len(data)) + 64
d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
for _ in range(2):
    res = c.compress_blocks_ptr(d_in.data_ptr(), len(data), d_out.data_ptr(), cap)
torch.cuda.synchronize()
print("ok", res["nbits"], c.stats()["kernel_launches"])
