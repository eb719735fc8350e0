#!/usr/bin/env python3
"""GPU decompressor against libbzip2 / the reference decoder on the debug corpus (development aid; run on a GPU box)."""
import sys, os, time, bz2, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bzip2_opencl_b200 import api, corpus
from oracle import refimpl

def main():
    cases = [
        ("one", b"x", 9), ("abc", b"abcabcabd" * 3, 9), ("hello", b"hello hello world, hello bzip2 on b200! " * 40, 9),
        ("runs", corpus.g_runs(), 9), ("text200k", corpus.word_text(200000, 2), 1), ("rand300k", corpus.rand_bytes(300000, 1), 1),
        ("zeros300k", bytes(300000), 9), ("text3m", corpus.word_text(3000000, 2), 9), ("rand1m", corpus.rand_bytes(1000003, 1), 9),
        ("ababx", b"ab" * 400000 + b"x", 9), ("zeros5m", bytes(5_000_000), 1), ("mixed8m", corpus.mixed_corpus(8 << 20, 1 << 20), 9),
        ("empty", b"", 9),
    ]
    ok = True
    for name, data, mult in cases:
        try:
            c = api.Compressor(mult, 8, 100000)
            ours = c.compress(data)
            ext = bz2.compress(data, mult)          # a stream made by libbzip2 (17-bit lengths, its own tables)
            for label, s in (("ours", ours), ("libbzip2", ext)):
                t = time.time()
                got = c.decompress(s)
                dt = time.time() - t
                good = got == data
                ok &= good
                print(f"{name:10s} {label:8s} {len(s):9d} -> {len(got):9d} {'OK' if good else 'MISMATCH'}  {dt*1e3:7.1f} ms dev {c.stats()['ms_total']:.1f} ms", flush=True)
            c.close()
        except Exception:
            traceback.print_exc(); ok = False
    # error paths
    c = api.Compressor(9, 8, 100000)
    s = bytearray(c.compress(corpus.word_text(300000, 3)))
    for label, mut in (("bad header", lambda b: b.__setitem__(1, 0x5b)), ("flip data bit", lambda b: b.__setitem__(len(b) // 2, b[len(b) // 2] ^ 0x10)),
                       ("truncated", lambda b: b.__delitem__(slice(len(b) - 20, len(b)))), ("bad stream crc", lambda b: b.__setitem__(len(b) - 2, b[len(b) - 2] ^ 1))):
        b = bytearray(s); mut(b)
        try:
            c.decompress(bytes(b)); print(f"{label}: NOT DETECTED"); ok = False
        except api.Bz2Error as e:
            print(f"{label}: {e}")
    print("ALL OK" if ok else "SOME FAILED")
    return 0 if ok else 1

if __name__ == "__main__":
    sys.exit(main())
