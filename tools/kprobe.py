#!/usr/bin/env python3
"""Per-kernel times of one compress call (development aid): python tools/kprobe.py kind MiB [P]"""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bzip2_opencl_b200 import api, corpus
kind, mb = sys.argv[1], int(sys.argv[2])
par = int(sys.argv[3]) if len(sys.argv) > 3 else 600
data = {"mixed": lambda: corpus.mixed_corpus(mb << 20, workers=16), "rand": lambda: corpus.rand_bytes(mb << 20, 4),
        "text": lambda: corpus.word_text(mb << 20, 101, markov=True), "rep": lambda: corpus.repetitive_corpus(mb << 20)}[kind]()
c = api.Compressor(9, par, 100000)
d_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
cap = c.max_output_bytes(len(data)) + 64
d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
for rep in range(3):
    if rep == 2: c.set_profiling(True)
    c.compress_blocks_ptr(d_in.data_ptr(), len(data), d_out.data_ptr(), cap)
torch.cuda.synchronize()
st = c.stats()
print(f"{kind} {mb} MiB: dev {st['ms_total']:.1f} ms = {len(data)/1e6/(st['ms_total']/1e3):.0f} MB/s rounds {st['bwt_rounds']} rho {st['bwt_active_volume']/max(1,st['n_rle']):.2f} blocks {st['n_blocks']}")
kt = c.kernel_times()
print({k: (round(v[0], 2), v[1]) for k, v in sorted(kt.items(), key=lambda kv: -kv[1][0])[:12] if v[1]})
