#!/usr/bin/env python3
"""Stage-by-stage comparison of the CUDA path against the oracle (development aid; run on a GPU box)."""
import sys, os, time, bz2, traceback
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bzip2_opencl_b200 import api, corpus
from oracle import oracle


def first_diff(a, b):
    n = min(len(a), len(b))
    d = np.nonzero(np.asarray(a[:n]) != np.asarray(b[:n]))[0]
    return int(d[0]) if len(d) else (n if len(a) != len(b) else -1)


def check(name, data, mult, base, parallel=64, verbose=True):
    ok = True
    c = api.Compressor(mult, parallel, base)
    t = time.time()
    r = c.compress_blocks(data)
    dt = time.time() - t
    st = c.stats()
    nb = len(r["crcs"])
    starts, lens, crcs = oracle.rle1_blocks(data, mult, base)
    print(f"== {name}: n={len(data)} base={base} mult={mult} blocks gpu={nb} oracle={len(starts)}  wall {dt*1e3:.1f} ms  dev {st['ms_total']:.2f} ms "
          f"(rle {st['ms_rle']:.2f} sort0 {st['ms_bwt_sort0']:.2f} rounds {st['ms_bwt_rounds']:.2f}[{st['bwt_rounds']}] fin {st['ms_bwt_finish']:.2f} "
          f"mtf {st['ms_mtf']:.2f} huf {st['ms_huffman']:.2f} emit {st['ms_emit']:.2f}) launches {st['kernel_launches']}", flush=True)
    if nb != len(starts):
        print("   BLOCK COUNT MISMATCH"); ok = False
    gs = c.scan_blocks(data)
    if len(gs) != len(starts) or (len(gs) and first_diff(gs, starts) >= 0):
        print("   STARTS MISMATCH", gs[:5], starts[:5]); ok = False
    if nb and first_diff(r["crcs"], crcs) >= 0:
        i = first_diff(r["crcs"], crcs)
        print(f"   CRC MISMATCH at block {i}: {r['crcs'][i]:08x} vs {crcs[i]:08x}"); ok = False
    # stage dumps for blocks of the last batch
    b_lo = (nb - 1) // parallel * parallel if nb else 0
    def resident(b):
        try:
            c.stage(b, api.STAGE_ORIGPTR)
            return True
        except api.Bz2Error:
            return False            # batches are cut by the lanes / walk chunks, not at multiples of `parallel`
    for b in sorted(set([x for x in (b_lo, nb - 1) if resident(x)])) if nb else []:
        o = oracle.block_stages(data, b, mult, base)
        for stage, key, nm in ((api.STAGE_RLE1, "rle", "RLE1"), (api.STAGE_PRESENT, "present", "PRESENT"), (api.STAGE_BWT, "L", "BWT L"),
                               (api.STAGE_MTF, "mtf", "MTF"), (api.STAGE_FREQ, "freq", "FREQ"), (api.STAGE_SELECTORS, "selectors", "SELECTORS")):
            g = c.stage(b, stage)
            e = o[key].astype(g.dtype) if key != "present" else o[key].astype(np.uint8)
            d = first_diff(g, e)
            if d >= 0:
                print(f"   block {b}: {nm} mismatch at {d} (gpu len {len(g)} oracle len {len(e)}): gpu {g[max(0,d-2):d+6]} oracle {e[max(0,d-2):d+6]}"); ok = False
        g = c.stage(b, api.STAGE_ORIGPTR)[0]
        if g != o["orig_ptr"]:
            print(f"   block {b}: ORIGPTR gpu {g} oracle {o['orig_ptr']}"); ok = False
        gl = c.stage(b, api.STAGE_LENGTHS).reshape(-1, 258)
        el = o["code_lengths"]
        if gl.shape != el.shape or (gl != el).any():
            print(f"   block {b}: LENGTHS mismatch shapes {gl.shape} {el.shape}");
            if gl.shape == el.shape:
                t_, s_ = np.nonzero(gl != el); print("      first at table", t_[0], "sym", s_[0], gl[t_[0], s_[0]], el[t_[0], s_[0]])
            ok = False
        if int(r["block_nbits"][b]) != o["block_bits"]:
            print(f"   block {b}: NBITS gpu {int(r['block_nbits'][b])} oracle {o['block_bits']}"); ok = False
    full = c.compress(data)
    exp = oracle.compress(data, mult, base)
    if full != exp:
        d = first_diff(np.frombuffer(full, np.uint8), np.frombuffer(exp, np.uint8))
        print(f"   STREAM MISMATCH: len gpu {len(full)} oracle {len(exp)} first diff byte {d}"); ok = False
    try:
        if base == 100000 and bz2.decompress(full) != data:
            print("   bz2 ROUNDTRIP MISMATCH"); ok = False
    except Exception as e:
        print("   bz2 decode failed:", e); ok = False
    print("   ->", "OK" if ok else "FAIL", flush=True)
    c.close()
    return ok


def main():
    cases = [
        ("one", b"x", 9, 100000),
        ("abc", b"abcabcabd" * 3, 9, 100000),
        ("hello", b"hello hello world, hello bzip2 on b200! " * 40, 9, 100000),
        ("runs", corpus.g_runs(), 9, 100000),
        ("text200k", corpus.word_text(200000, 2), 1, 100000),
        ("rand300k", corpus.rand_bytes(300000, 1), 1, 100000),
        ("zeros300k", bytes(300000), 9, 100000),
        ("text3m", corpus.word_text(3000000, 2), 9, 100000),
        ("text3m-10k", corpus.word_text(3000000, 2), 9, 10000),
        ("rand1m", corpus.rand_bytes(1000003, 1), 9, 100000),
        ("ababx", b"ab" * 400000 + b"x", 9, 100000),
        ("abab", b"ab" * 5000, 9, 100000),
        ("zeros5m", bytes(5_000_000), 1, 100000),
        ("mixed4m", corpus.mixed_corpus(4 << 20, 1 << 20), 9, 100000),
    ]
    sel = sys.argv[1:] 
    allok = True
    for name, data, mult, base in cases:
        if sel and name not in sel:
            continue
        try:
            allok &= check(name, data, mult, base)
        except Exception:
            traceback.print_exc()
            allok = False
            break
    print("ALL OK" if allok else "SOME FAILED")
    return 0 if allok else 1


if __name__ == "__main__":
    sys.exit(main())
