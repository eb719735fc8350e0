#!/usr/bin/env python3
"""Full-size runs of the five BASELINE.json configs on ONE B200 (run on a GPU box), with bounded host-side verification.

  python tools/run_configs.py [c1 c2 c3 c4 c5] [--out profiles/rNN_configs.json] [--quick]

Per config: device-resident throughput (CUDA events inside the library), end-to-end throughput from host memory, and parity:
  * byte-for-byte against the reference built for the CPU (oracle/_ref, all host threads) on the whole input where that takes
    seconds, else on a stated prefix / sample of blocks;
  * every stream decodes with libbzip2 (python `bz2`, incremental, bounded prefix for the multi-GiB runs) and, on a prefix,
    with the reference's own decoder (InputStream/BlockDecompressor, base-100000 build);
  * C3's non-primitive blocks are compared with the origPtr field masked;
  * C4 is also compressed through the multi-device handle (every GPU of the box; one GPU named four times on a one-GPU box) and must
    give the same file.
The oracle is used only as the checker; the product path is libbz2b200.so.
"""
from __future__ import annotations

import argparse
import bz2
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bzip2_opencl_b200 import api, corpus, multi  # noqa: E402
from oracle import oracle, refimpl  # noqa: E402

BASE = 100000
THREADS = os.cpu_count() or 1


def pinned(data: bytes) -> torch.Tensor:
    t = torch.empty(len(data), dtype=torch.uint8, pin_memory=True)
    t.numpy()[:] = np.frombuffer(data, dtype=np.uint8)
    return t


def gpu_compress(comp: api.Compressor, host: torch.Tensor, reps: int = 2):
    """Returns (stream bytes, result dict, device-resident MB/s, e2e MB/s, stats of the device-resident run)."""
    n = host.numel()
    cap = comp.max_output_bytes(n) + 64
    d_in = host.cuda()
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    h_out = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
    best_dev, st_dev, res = None, None, None
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        t = time.perf_counter()
        res = comp.compress_blocks_ptr(d_in.data_ptr(), n, d_out.data_ptr(), cap)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        if best_dev is None or dt < best_dev:
            best_dev, st_dev = dt, comp.stats()
    del d_in, d_out
    best_e2e = None
    for _ in range(reps):
        t = time.perf_counter()
        res2 = comp.compress_blocks_ptr(host.data_ptr(), n, h_out.data_ptr(), cap)
        dt = time.perf_counter() - t
        best_e2e = dt if best_e2e is None else min(best_e2e, dt)
    assert res2["nbits"] == res["nbits"] and (res2["crcs"] == res["crcs"]).all()
    bits = h_out[:(res["nbits"] + 7) // 8].numpy().tobytes()
    stream = api.stream_from_blocks(comp.multiplier, [(bits, res["nbits"], res["crcs"])])
    torch.cuda.empty_cache()
    return stream, res, n / 1e6 / best_dev, n / 1e6 / best_e2e, st_dev


def bz2_prefix_ok(stream: bytes, data: bytes, limit: int) -> tuple[bool, int]:
    """Decode (at most) the first `limit` bytes with libbzip2 and compare them with the input."""
    d = bz2.BZ2Decompressor()
    got = 0
    pos = 0
    step = 1 << 22
    while got < min(limit, len(data)) and pos < len(stream) and not d.eof:
        out = d.decompress(stream[pos:pos + step])
        pos += step
        if out:
            if data[got:got + len(out)] != out:
                return False, got
            got += len(out)
    return (got >= min(limit, len(data))), got


def ref_decoder_ok(comp_mult: int, data_prefix: bytes, parallel: int = 64) -> bool:
    """The reference's own decoder on a freshly compressed prefix (its decoder reads whole streams only)."""
    c = api.Compressor(comp_mult, parallel, BASE)
    s = c.compress(data_prefix)
    c.close()
    return refimpl.decompress(s, len(data_prefix) + 16, BASE) == data_prefix


def stats_row(st: dict) -> dict:
    keys = ("ms_total", "ms_rle", "ms_bwt_sort0", "ms_bwt_rounds", "ms_mtf", "ms_huffman", "ms_emit", "ms_walk")
    r = {k: round(float(st[k]), 2) for k in keys}
    r["bwt_rounds"] = int(st["bwt_rounds"])
    r["rho"] = round(st["bwt_active_volume"] / max(1, st["n_rle"]), 3)
    r["blocks"] = int(st["n_blocks"])
    return r


def run_c1(quick):
    data = corpus.word_text((16 if quick else 64) << 20, 101, markov=True)
    comp = api.Compressor(9, 80, BASE)
    stream, res, dev, e2e, st = gpu_compress(comp, pinned(data))
    t = time.perf_counter()
    ref = refimpl.compress(data, 9, BASE, threads=THREADS)
    tref = time.perf_counter() - t
    ok_rt, got = bz2_prefix_ok(stream, data, len(data))
    row = dict(config="C1 64 MiB Markov text, -9", input_bytes=len(data), out_bytes=len(stream), device_MBps=round(dev, 1), e2e_MBps=round(e2e, 1),
               identical_to_reference=stream == ref, reference_cpu_MBps=round(len(data) / 1e6 / tref, 1), reference_threads=THREADS,
               bz2_roundtrip_bytes=got, bz2_roundtrip_ok=ok_rt, reference_decoder_ok=ref_decoder_ok(9, data[:8 << 20]),
               sha256=hashlib.sha256(stream).hexdigest()[:16], stages=stats_row(st))
    comp.close()
    return row


def run_c2(quick):
    data = corpus.mixed_corpus((128 if quick else 1024) << 20, workers=min(32, THREADS))
    comp = api.Compressor(9, 600, BASE)
    stream, res, dev, e2e, st = gpu_compress(comp, pinned(data))
    t = time.perf_counter()
    ref = refimpl.compress(data, 9, BASE, threads=THREADS)
    tref = time.perf_counter() - t
    ok_rt, got = bz2_prefix_ok(stream, data, 256 << 20)
    row = dict(config="C2 1 GiB mixed corpus, -9, 1 GPU", input_bytes=len(data), out_bytes=len(stream), device_MBps=round(dev, 1), e2e_MBps=round(e2e, 1),
               identical_to_reference=stream == ref, reference_cpu_MBps=round(len(data) / 1e6 / tref, 1), reference_threads=THREADS,
               bz2_roundtrip_bytes=got, bz2_roundtrip_ok=ok_rt, sha256=hashlib.sha256(stream).hexdigest()[:16], stages=stats_row(st))
    comp.close()
    return row


def run_c3(quick):
    data = corpus.repetitive_corpus((32 if quick else 256) << 20)
    comp = api.Compressor(9, 300, BASE)
    stream, res, dev, e2e, st = gpu_compress(comp, pinned(data))
    ok_rt, got = bz2_prefix_ok(stream, data, len(data))
    # origPtr-masked comparison with the oracle on a slice of every quarter (the oracle's CPU BWT needs seconds per periodic block)
    q = len(data) // 4
    oracle_ok, masked_ok, checked, nonprim = True, True, 0, 0
    c2 = api.Compressor(9, 16, BASE)
    for k in range(4):
        part = data[k * q:k * q + (3 << 20)]
        exp, offs, flags = oracle.compress_ex(part, 9, BASE)
        got_s = c2.compress(part)
        oracle_ok &= got_s == exp                       # the oracle and the CUDA path share the first-of-tie-group origPtr rule
        ref = refimpl.compress(part, 9, BASE, threads=THREADS)
        masked_ok &= oracle.mask_origptr(got_s, offs, flags) == oracle.mask_origptr(ref, offs, flags)
        nonprim += int(flags.sum())
        checked += len(part)
    c2.close()
    row = dict(config="C3 256 MiB repetitive (zeros / short periods / long-LCP text / near-periodic), -9", input_bytes=len(data), out_bytes=len(stream),
               device_MBps=round(dev, 1), e2e_MBps=round(e2e, 1), bz2_roundtrip_bytes=got, bz2_roundtrip_ok=ok_rt,
               identical_to_oracle_on_slices=bool(oracle_ok), origptr_masked_equal_to_reference_on_slices=bool(masked_ok), nonprimitive_blocks_in_slices=nonprim,
               slices_checked_bytes=checked, sha256=hashlib.sha256(stream).hexdigest()[:16],
               stages=stats_row(st))
    comp.close()
    return row


def run_c4(quick):
    n = (512 if quick else 8192) << 20
    data = corpus.rand_bytes(n, 4)
    host = pinned(data)
    comp = api.Compressor(9, 600, BASE)
    d_all = host.cuda()
    torch.cuda.synchronize()
    t = time.perf_counter()
    starts = comp.scan_blocks(ptr=d_all.data_ptr(), nbytes=n)          # the standalone boundary scan (summary tree + walk)
    torch.cuda.synchronize()
    t_scan = time.perf_counter() - t
    cap = comp.max_output_bytes(n) + 64
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    best = None
    for _ in range(2):
        torch.cuda.synchronize()
        t = time.perf_counter()
        r = comp.compress_blocks_ptr(d_all.data_ptr(), n, d_out.data_ptr(), cap)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    bits = d_out[:(r["nbits"] + 7) // 8].cpu().numpy().tobytes()
    stream = api.stream_from_blocks(9, [(bits, r["nbits"], r["crcs"])])
    comp.close()
    # the same bytes through the multi-device handle (every GPU of the box; on a one-GPU box the device four times): slices, chained
    # walk, per-device bit shift -- the stream must be the same file
    ndev = api.load_library().bz2b200_device_count()
    devs = list(range(ndev)) if ndev > 1 else [0, 0, 0, 0]
    many = api.Compressor(9, 600 if ndev > 1 else 150, BASE, devices=devs)
    torch.cuda.synchronize()
    t = time.perf_counter()
    r2 = many.compress_blocks_ptr(d_all.data_ptr(), n, d_out.data_ptr(), cap)
    torch.cuda.synchronize()
    t_multi = time.perf_counter() - t
    bits2 = d_out[:(r2["nbits"] + 7) // 8].cpu().numpy().tobytes()
    same = r2["nbits"] == r["nbits"] and bits2 == bits and (r2["crcs"] == r["crcs"]).all()
    many.close()
    del d_all, d_out
    torch.cuda.empty_cache()
    lim = (64 if quick else 512) << 20
    ok_rt, got = bz2_prefix_ok(stream, data, lim)
    t = time.perf_counter()
    ref = refimpl.compress(data[:lim], 9, BASE, threads=THREADS)
    tref = time.perf_counter() - t
    ours_prefix = api.Compressor(9, 64, BASE)
    mine = ours_prefix.compress(data[:lim])
    ours_prefix.close()
    row = dict(config="C4 8 GiB LCG random bytes, -9: one GPU, and the multi-device handle over " + (f"{ndev} GPUs" if ndev > 1 else "one GPU named four times"),
               input_bytes=n, out_bytes=len(stream), blocks=len(starts), device_MBps_one_gpu=round(n / 1e6 / best, 1),
               multi_handle_MBps=round(n / 1e6 / t_multi, 1), multi_handle_devices=devs, multi_handle_stream_identical=bool(same),
               boundary_scan_ms=round(1e3 * t_scan, 1), ratio=round(len(stream) / n, 5),
               bz2_roundtrip_bytes=got, bz2_roundtrip_ok=ok_rt, prefix_identical_to_reference=mine == ref, prefix_bytes=lim,
               reference_cpu_MBps=round(lim / 1e6 / tref, 1), reference_threads=THREADS, sha256=hashlib.sha256(stream).hexdigest()[:16])
    return row


def _text_seg(args):
    k, m = args
    return corpus.word_text(m, 5 + k, markov=True)


def run_c5(quick):
    from concurrent.futures import ProcessPoolExecutor
    n = (256 if quick else 4096) << 20
    seg = 64 << 20
    jobs = [(k, min(seg, n - k * seg)) for k in range((n + seg - 1) // seg)]
    with ProcessPoolExecutor(max_workers=min(32, THREADS)) as ex:
        data = b"".join(ex.map(_text_seg, jobs))
    host = pinned(data)
    rows = []
    lim = 64 << 20
    for mult in range(1, 10):
        comp = api.Compressor(mult, 600 * 9 // mult, BASE)
        stream, res, dev, e2e, st = gpu_compress(comp, host, reps=1)
        ok_rt, got = bz2_prefix_ok(stream, data, lim)
        ref = refimpl.compress(data[:lim], mult, BASE, threads=THREADS)
        small = api.Compressor(mult, 64, BASE)
        mine = small.compress(data[:lim])
        small.close()
        rows.append(dict(multiplier=mult, block_size=mult * BASE, out_bytes=len(stream), device_MBps=round(dev, 1), e2e_MBps=round(e2e, 1),
                         bz2_roundtrip_bytes=got, bz2_roundtrip_ok=ok_rt, prefix_identical_to_reference=mine == ref, prefix_bytes=lim,
                         reference_decoder_ok=refimpl.decompress(mine[:], lim + 16, BASE) == data[:lim] if mult in (1, 5, 9) else None,
                         rho=round(st["bwt_active_volume"] / max(1, st["n_rle"]), 3), blocks=int(st["n_blocks"])))
        comp.close()
        print("  C5", rows[-1], flush=True)
    return dict(config="C5 block-size sweep -1..-9 on 4 GiB Markov text (64 MiB segments, seeds 5+k), 1 GPU", input_bytes=n, sweep=rows)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.json"))
    ap.add_argument("--quick", action="store_true", help="scaled-down sizes (development)")
    args = ap.parse_args()
    oracle.build()
    rows = []
    for name in args.which:
        t = time.perf_counter()
        row = {"c1": run_c1, "c2": run_c2, "c3": run_c3, "c4": run_c4, "c5": run_c5}[name](args.quick)
        row["wall_s"] = round(time.perf_counter() - t, 1)
        rows.append(row)
        print(json.dumps(row), flush=True)
        json.dump({"gpu": torch.cuda.get_device_name(0), "host_threads": THREADS, "quick": args.quick, "rows": rows}, open(args.out, "w"), indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main())
