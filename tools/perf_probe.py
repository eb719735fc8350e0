#!/usr/bin/env python3
"""Throughput / stage-time probe on a GPU box (development aid)."""
import sys, os, time, bz2, hashlib
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bzip2_opencl_b200 import api, corpus


def probe(name, data, mult=9, base=100000, parallel=64, reps=2, verify=False):
    n = len(data)
    c = api.Compressor(mult, parallel, base)
    d_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    cap = c.max_output_bytes(n) + 64
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    best = None
    for r in range(reps + 1):
        torch.cuda.synchronize()
        t = time.time()
        res = c.compress_blocks_ptr(d_in.data_ptr(), n, d_out.data_ptr(), cap)
        torch.cuda.synchronize()
        dt = time.time() - t
        st = c.stats()
        if r > 0 and (best is None or dt < best[0]):
            best = (dt, st)
    dt, st = best
    print(f"{name:14s} n={n/1e6:8.1f}MB P={parallel:4d} blocks={st['n_blocks']:5d} wall {dt*1e3:8.1f} ms = {n/dt/1e6:8.1f} MB/s | dev {st['ms_total']:8.1f} "
          f"rle {st['ms_rle']:6.1f} sort0 {st['ms_bwt_sort0']:6.1f} rounds {st['ms_bwt_rounds']:7.1f}[{st['bwt_rounds']:2d}, rho={st['bwt_active_volume']/max(1,st['n_rle']):.2f}] "
          f"fin {st['ms_bwt_finish']:5.1f} mtf {st['ms_mtf']:6.1f} huf {st['ms_huffman']:6.1f} emit {st['ms_emit']:6.1f} walk {st['ms_walk']:6.1f} d2h {st['ms_d2h']:5.1f} launches {st['kernel_launches']}",
          flush=True)
    if verify:
        nbits = res["nbits"]
        bits = d_out[:(nbits + 7) // 8].cpu().numpy().tobytes()
        full = api.stream_from_blocks(mult, [(bits, nbits, res["crcs"])])
        t = time.time()
        ok = bz2.decompress(full) == data
        print(f"   bz2 roundtrip {'OK' if ok else 'FAIL'} ({time.time()-t:.1f}s) size {len(full)} sha {hashlib.sha256(full).hexdigest()[:16]}", flush=True)
    c.close()
    del d_in, d_out
    torch.cuda.empty_cache()


if __name__ == "__main__":
    which = sys.argv[1:] or ["text64", "mixed256", "rand256", "rep64"]
    if "text64" in which:
        d = corpus.word_text(64 << 20, 101, markov=True)
        probe("text64", d, parallel=16)
        probe("text64", d, parallel=80, verify=True)
    if "mixed256" in which:
        d = corpus.mixed_corpus(256 << 20)
        probe("mixed256", d, parallel=64)
        probe("mixed256", d, parallel=320, verify=True)
    if "rand256" in which:
        d = corpus.rand_bytes(256 << 20, 4)
        probe("rand256", d, parallel=320, verify=True)
    if "rep64" in which:
        d = corpus.repetitive_corpus(64 << 20)
        probe("rep64", d, parallel=80, verify=True)
