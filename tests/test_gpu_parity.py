"""Parity tests proper: the CUDA path, called through the C ABI (libbz2b200.so), against the oracle on the same inputs.
Bar: bit-exact.  Every produced stream must also decode with libbzip2 (Python bz2) and the reference's own decoder."""
import bz2
import ctypes as C
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from bzip2_opencl_b200 import api, corpus, multi
from oracle import oracle, refimpl
from conftest import have_ref, ROOT

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "appendix_e.json")))
CONFIGS = {"A": (10000, 9), "B": (100000, 9), "C": (100000, 1)}
NAMES = [k for k in GOLD if not k.startswith("_")]

_comp = {}


def comp(mult=9, base=100000, parallel=24):
    key = (mult, base, parallel)
    if key not in _comp:
        if len(_comp) > 6:
            for c in _comp.values():
                c.close()
            _comp.clear()
        _comp[key] = api.Compressor(mult, parallel, base)
    return _comp[key]


def test_native_library_is_what_runs():
    lib = api.load_library()
    assert lib.bz2b200_device_count() >= 1
    c = comp()
    c.compress(b"hello")
    assert c.stats()["kernel_launches"] > 20  # the CUDA kernels ran; there is no other path


@pytest.mark.parametrize("cfg", list(CONFIGS))
@pytest.mark.parametrize("name", NAMES)
def test_golden_vectors(name, cfg):
    base, mult = CONFIGS[cfg]
    d = corpus.GOLDEN_INPUTS[name]()
    out = comp(mult, base).compress(d)
    size, sha = GOLD[name][cfg]
    assert len(out) == size
    assert out == oracle.compress(d, mult, base)            # bit-exact vs the oracle (same origPtr convention)
    if GOLD[name]["primitive"]:
        assert hashlib.sha256(out).hexdigest()[:16] == sha  # and vs the committed reference hash
    if base == 100000:
        assert bz2.decompress(out) == d
    if have_ref(base):
        assert refimpl.decompress(out, len(d) + 16, base) == d


def _stage_check(c, d, mult, base, blocks):
    for b in blocks:
        o = oracle.block_stages(d, b, mult, base)
        assert (c.stage(b, api.STAGE_RLE1) == o["rle"]).all(), f"RLE1 block {b}"
        assert (c.stage(b, api.STAGE_PRESENT).astype(bool) == o["present"]).all(), f"presence block {b}"
        assert (c.stage(b, api.STAGE_BWT) == o["L"]).all(), f"BWT block {b}"
        assert int(c.stage(b, api.STAGE_ORIGPTR)[0]) == o["orig_ptr"], f"origPtr block {b}"
        assert (c.stage(b, api.STAGE_MTF) == o["mtf"]).all(), f"MTF block {b}"
        assert (c.stage(b, api.STAGE_FREQ) == o["freq"]).all(), f"freq block {b}"
        assert (c.stage(b, api.STAGE_SELECTORS) == o["selectors"]).all(), f"selectors block {b}"
        assert (c.stage(b, api.STAGE_LENGTHS).reshape(-1, 258) == o["code_lengths"]).all(), f"lengths block {b}"


def test_stage_dumps_match_oracle():
    d = corpus.word_text(250000, 77) + bytes(5000) + corpus.rand_bytes(140000, 3) + corpus.g_runs()[:100000]
    for mult, base in ((1, 100000), (9, 10000)):
        c = comp(mult, base, 64)
        r = c.compress_blocks(d)
        starts, lens, crcs = oracle.rle1_blocks(d, mult, base)
        assert (c.scan_blocks(d) == starts).all()
        assert (r["crcs"] == crcs).all()
        _stage_check(c, d, mult, base, range(len(starts)))
        # per-block bit lengths add up and match the oracle's
        for b in range(len(starts)):
            assert int(r["block_nbits"][b]) == oracle.block_stages(d, b, mult, base)["block_bits"]
        assert int(r["block_nbits"].sum()) == r["nbits"]


def test_empty_and_tiny_inputs():
    c = comp(9, 100000)
    assert c.compress(b"") == oracle.compress(b"", 9, 100000) and len(c.compress(b"")) == 14
    for d in (b"a", b"ab", b"aa", b"aaaa", b"aaaaa", b"\x00", b"\xff" * 3, bytes(range(256)), b"abracadabra"):
        out = c.compress(d)
        assert out == oracle.compress(d, 9, 100000)
        assert bz2.decompress(out) == d
    r = c.compress_blocks(b"")
    assert r["nbits"] == 0 and len(r["crcs"]) == 0 and r["consumed"] == 0


@pytest.mark.parametrize("base", [1000, 3000])
def test_block_boundary_edges(base):
    # lengths around the refusal limit (blockLength > bs-6), long runs straddling boundaries, 255-run chunking
    c = comp(1, base, 16)
    bs = base
    rng = np.random.default_rng(5)
    cases = [bytes((i * 7 + 3) % 251 for i in range(n)) for n in (bs - 7, bs - 6, bs - 5, bs - 4, bs - 1, bs, bs + 1, 2 * bs + 3)]
    cases += [bytes(n) for n in (254, 255, 256, 509, 510, 511, 5 * bs, 51 * bs + 17)]
    cases += [bytes((i * 7 + 3) % 251 for i in range(bs - 20)) + b"z" * k + b"tail" * 50 for k in (1, 3, 4, 5, 6, 254, 255, 256, 300, 1000)]
    for _ in range(6):
        runs = rng.integers(1, 700, 60)
        vals = rng.integers(0, 4, 60, dtype=np.uint8)
        cases.append(np.repeat(vals, runs).tobytes())
    for d in cases:
        out = c.compress(d)
        exp, offs, nonprim = oracle.compress_ex(d, 1, base)
        assert out == exp, f"n={len(d)}"
        s, l, crc = oracle.rle1_blocks(d, 1, base)
        assert (c.scan_blocks(d) == s).all()


def test_ragged_sizes_random_and_text():
    rng = np.random.default_rng(11)
    c = comp(2, 10000, 8)
    for i in range(12):
        n = int(rng.integers(1, 150000))
        d = rng.integers(0, 256 if i % 2 else 3, n, dtype=np.uint8).tobytes() if i % 3 else corpus.word_text(n, i)
        assert c.compress(d) == oracle.compress(d, 2, 10000), f"case {i} n={n}"


def test_output_independent_of_parallel_block_count():
    # the reference's output depends on -p; ours must not
    d = corpus.word_text(700000, 21) + corpus.rand_bytes(300000, 22)
    outs = {p: api.Compressor(1, p, 100000).compress(d) for p in (1, 3, 16)}
    assert outs[1] == outs[3] == outs[16] == oracle.compress(d, 1, 100000)


def test_streaming_calls_with_carry_bits():
    # not-final calls return only complete blocks and how much input they consumed; start_bit carries the bit offset
    d = corpus.word_text(430000, 31) + bytes(300000) + corpus.rand_bytes(150000, 32)
    mult, base = 1, 100000
    c = comp(mult, base, 4)
    parts, pos, startbit = [], 0, 0
    chunk = 260000
    buf = b""
    fed = 0
    while fed < len(d) or buf:
        take = d[fed:fed + chunk]
        fed += len(take)
        buf += take
        final = fed >= len(d)
        r = c.compress_blocks(buf, is_final=final, start_bit=startbit)
        bits = np.unpackbits(np.frombuffer(r["bits"], dtype=np.uint8))[startbit:startbit + r["nbits"]]
        assert not np.unpackbits(np.frombuffer(r["bits"], dtype=np.uint8))[:startbit].any()
        parts.append((np.packbits(bits).tobytes(), int(r["nbits"]), r["crcs"]))
        startbit = (startbit + r["nbits"]) % 8
        buf = buf[r["consumed"]:]
        if final:
            assert not buf
            break
    assert api.stream_from_blocks(mult, parts) == oracle.compress(d, mult, base)


def test_device_pointers_in_and_out():
    import torch
    d = corpus.mixed_corpus(3 << 20, 1 << 20)
    c = comp(9, 100000)
    t_in = torch.frombuffer(bytearray(d), dtype=torch.uint8).cuda()
    cap = c.max_output_bytes(len(d)) + 64
    t_out = torch.zeros(cap, dtype=torch.uint8, device="cuda")
    r = c.compress_blocks_ptr(t_in.data_ptr(), len(d), t_out.data_ptr(), cap)
    bits = t_out[:(r["nbits"] + 7) // 8].cpu().numpy().tobytes()
    assert api.stream_from_blocks(9, [(bits, r["nbits"], r["crcs"])]) == oracle.compress(d, 9, 100000)
    # misaligned device input is staged internally
    r2 = c.compress_blocks_ptr(t_in.data_ptr() + 1, len(d) - 1, t_out.data_ptr(), cap)
    bits2 = t_out[:(r2["nbits"] + 7) // 8].cpu().numpy().tobytes()
    assert api.stream_from_blocks(9, [(bits2, r2["nbits"], r2["crcs"])]) == oracle.compress(d[1:], 9, 100000)


def test_output_too_small_is_reported():
    c = comp(9, 100000)
    d = corpus.rand_bytes(100000, 1)
    src = np.frombuffer(d, dtype=np.uint8)
    out = np.zeros(1000, dtype=np.uint8)
    n = C.c_size_t(0)
    rc = c.lib.bz2b200_compress_stream(c.h, src.ctypes.data_as(C.c_void_p), len(d), out.ctypes.data_as(C.c_void_p), 1000, C.byref(n))
    assert rc == api.E_OUTPUT


def test_sharded_block_ranges_reassemble_bit_exactly():
    # the N-GPU path on one GPU: every rank's byte range goes through the C ABI separately, rank 0 concatenates
    d = corpus.word_text(900000, 41) + bytes(200000) + corpus.rand_bytes(500000, 42)
    mult, base = 2, 100000
    c = comp(mult, base, 4)
    starts = c.scan_blocks(d)
    whole = oracle.compress(d, mult, base)
    for world in (2, 3, 8):
        parts = [multi.compress_shard(lambda buf, final: c.compress_blocks(buf, final), d, rng) for rng in multi.shard_byte_ranges(starts, len(d), world)]
        assert api.stream_from_blocks(mult, parts) == whole


def test_host_input_arrives_in_segments():
    # host input above the first 16 MiB segment: H2D copies, the tree of summaries and the walk are extended segment by
    # segment while earlier blocks are already being compressed (pinned and pageable memory take different copy schedules);
    # a run crossing the segment edge and a block boundary near it must not show
    import torch
    head = corpus.word_text((16 << 20) - 1000, 61)
    d = head + bytes(5000) + corpus.mixed_corpus(20 << 20, 4 << 20) + b"z" * 700
    c = comp(9, 100000, 16)
    exp = refimpl.compress(d, 9, 100000, threads=os.cpu_count() or 8) if have_ref(100000) else oracle.compress(d, 9, 100000)
    assert c.compress(d) == exp                                  # pageable host memory
    t = torch.empty(len(d), dtype=torch.uint8, pin_memory=True)
    t.numpy()[:] = np.frombuffer(d, dtype=np.uint8)
    cap = c.max_output_bytes(len(d)) + 64
    out = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
    r = c.compress_blocks_ptr(t.data_ptr(), len(d), out.data_ptr(), cap)   # pinned: every copy queued up front
    bits = out[:(r["nbits"] + 7) // 8].numpy().tobytes()
    assert api.stream_from_blocks(9, [(bits, r["nbits"], r["crcs"])]) == exp
    assert (c.scan_blocks(d) == oracle.rle1_blocks(d, 9, 100000)[0]).all()


def test_alternative_sort_paths_give_the_same_bits(monkeypatch):
    # development switches: three-kernel radix passes instead of the single sweep; every group through the global sort
    d = corpus.mixed_corpus(3 << 20, 1 << 20) + b"ab" * 300000 + bytes(100000)
    exp = oracle.compress(d, 9, 100000)
    # byte keys instead of keys over the block's alphabet (5 and 4 bytes)
    for env in ({"BZ2B200_ONESWEEP": "0"}, {"BZ2B200_LOCALSORT": "0"}, {"BZ2B200_ONESWEEP": "0", "BZ2B200_LOCALSORT": "0"},
                {"BZ2B200_PACK": "0"}, {"BZ2B200_PACK": "0", "BZ2B200_KEY0_BYTES": "4"},
                # round-2 switches: 10-bit digits / grouped sweeps in round 0, blocks never sorted through their root, byte loads
                # instead of the TMA bulk copy, the tree-descent walk
                {"BZ2B200_R0_BITS": "10"}, {"BZ2B200_R0_GROUP": "2"}, {"BZ2B200_R0_BITS": "10", "BZ2B200_R0_GROUP": "3"}, {"BZ2B200_ROOT": "0"},
                {"BZ2B200_BULK": "0"}, {"BZ2B200_WALK": "0"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        c = api.Compressor(9, 8, 100000)
        try:
            assert c.compress(d) == exp, env
        finally:
            c.close()
            for k in env:
                monkeypatch.delenv(k)


def test_alphabet_sizes_and_mixed_batches():
    # round-0 keys are packed over each block's own alphabet: bits per symbol change at 2/3, 4/5, 8/9, 16/17, 32/33, 64/65, 128/129 used
    # values, the number of symbols per key (16 .. 5) with them, and every block of a batch doubles from its own depth
    rng = np.random.default_rng(20261018)
    parts = []
    for alpha in (1, 2, 3, 4, 5, 8, 9, 16, 17, 32, 33, 64, 65, 128, 129, 200, 256):
        vals = np.sort(rng.choice(256, size=alpha, replace=False)).astype(np.uint8)
        body = vals[rng.integers(0, alpha, size=23_000)]
        body[5000:5600] = body[4000:4600]                       # a repeat, so that some rotations stay tied for a few rounds
        parts.append(body.tobytes())
    d = b"".join(parts)
    for mult, base, par in ((2, 10000, 64), (1, 100000, 3), (9, 100000, 8)):   # blocks of one alphabet each / of several / one block
        out = comp(mult, base, par).compress(d)
        assert out == oracle.compress(d, mult, base), (mult, base)
        if base == 100000:
            assert bz2.decompress(out) == d
    c = comp(2, 10000, 64)                                       # per-block stages of a block with a 3-symbol and one with a 129-symbol alphabet
    r = c.compress_blocks(d, True)
    nb = len(r["crcs"])
    for b in (nb - 1, nb - 2):
        o = oracle.block_stages(d, b, 2, 10000)
        assert (c.stage(b, api.STAGE_BWT) == o["L"]).all() and c.stage(b, api.STAGE_ORIGPTR)[0] == o["orig_ptr"]


def test_table_count_thresholds_and_unused_tables():
    # one small block each: MTF lengths on both sides of 200 / 600 / 1200 / 2400 symbols (2 .. 6 tables, kernel.cpp:2808-2818) for alphabets
    # of 2 .. 256 values; with few symbols most tables get no group at all.  Same list as tests/test_oracle.py::small_block_cases, where the
    # oracle is pinned to the reference's own code on it.
    c = comp(9, 100000, 2)
    for alpha in (2, 3, 5, 30, 256):
        for n in (1, 2, 3, 49, 50, 51, 99, 100, 101, 150, 197, 198, 199, 200, 201, 202, 597, 598, 599, 600, 601, 1197, 1198, 1199, 1200, 1201,
                  2397, 2398, 2399, 2400, 2401, 2402, 3000, 5000):
            raw = np.frombuffer(corpus.rand_bytes(n, 1000 * alpha + n), dtype=np.uint8).astype(np.int64)
            d = (raw % alpha).astype(np.uint8).tobytes()
            assert c.compress(d) == oracle.compress(d, 9, 100000), (alpha, n)
    # and all of them as the blocks of ONE stream (base 1000: the C ABI's smallest block): batches of blocks with 2 .. 6 tables side by side
    raw = np.frombuffer(corpus.rand_bytes(150_000, 4242), dtype=np.uint8).astype(np.int64)
    d = b"".join((raw[i:i + 5000] % a).astype(np.uint8).tobytes() for i, a in zip(range(0, 150_000, 5000), [2, 3, 5, 30, 256, 4] * 5))
    assert comp(1, 1000, 64).compress(d) == oracle.compress(d, 1, 1000)
    assert comp(3, 1000, 7).compress(d) == oracle.compress(d, 3, 1000)


def test_run_structured_fuzz():
    # corpus.run_fuzz (run lengths on the RLE1 edges 3/4/5, 254..262, 510..516, small alphabets, short periods, copies of earlier stretches):
    # one-kilobyte to 90k blocks, so that block ends fall into every phase of a run and the sort meets every kind of tie.  The oracle is pinned
    # to the reference's own code on the same generator in tests/test_oracle.py.
    blocks = 0
    for i in range(60):
        d = corpus.run_fuzz(20000 + 1777 * i, 9000 + i)
        for mult, base, par in ((1, 1000, 64), (1 + i % 9, 1000, 5), (1 + (i * 5) % 9, 10000, 3)):
            c = comp(mult, base, par)
            r = c.compress_blocks(d)
            out = c.compress(d)
            assert out == oracle.compress(d, mult, base), (i, mult, base)
            starts, lens, crcs = oracle.rle1_blocks(d, mult, base)
            assert (r["crcs"] == crcs).all()
            blocks += len(crcs)
    assert blocks > 2500
    d = corpus.run_fuzz(3_000_000, 77)                           # and at real block sizes
    for mult in (1, 9):
        assert comp(mult, 100000, 8).compress(d) == oracle.compress(d, mult, 100000)


def test_code_lengths_limited_to_20_bits():
    # symbol counts that fall off like Fibonacci numbers make a plain Huffman code 21-22 bits deep; the allocator's
    # relocation path (kernel.cpp:2748-2805, huff.cuh mk_allocate) caps the lengths at 20.  tests/test_oracle.py shows that this input
    # hits the limit and that the oracle equals the reference's own code on it.
    d = corpus.skewed_bytes(880000, 1, 48, 13, 8) + corpus.skewed_bytes(880000, 2) + corpus.skewed_bytes(880000, 7)
    c = comp(9, 100000, 4)
    out = c.compress(d)
    assert out == oracle.compress(d, 9, 100000)
    assert bz2.decompress(out) == d
    c.compress_blocks(d, True)
    _stage_check(c, d, 9, 100000, [2])                          # the call's last block is resident: every stage, code lengths included
    assert int(c.stage(2, api.STAGE_LENGTHS).max()) == 20
    out2 = comp(9, 10000, 16).compress(d)                       # the stock base: 90k blocks, 30 of them
    assert out2 == oracle.compress(d, 9, 10000)


def test_lanes_never_change_the_stream(monkeypatch):
    # batches of a call run on 1..3 lanes (own stream + host thread each); the bit position passes from batch to batch in order,
    # so the stream, the per-block tables and the bytes drained so far are the same whatever the lane count or batch size
    d = corpus.mixed_corpus(9 << 20, 3 << 20) + bytes(1_200_000) + corpus.rand_bytes(2_000_000, 97) + b"ab" * 700_000
    exp = oracle.compress(d, 2, 100000)                       # 200k blocks: ~70 blocks through batches of 1..12
    ref_crcs = None
    for lanes, par in (("1", 5), ("2", 5), ("2", 2), ("3", 7), ("4", 24), ("2", 1)):
        monkeypatch.setenv("BZ2B200_LANES", lanes)
        c = api.Compressor(2, par, 100000)
        try:
            assert c.compress(d) == exp, (lanes, par)
            r = c.compress_blocks(d, True)
            if ref_crcs is None:
                ref_crcs, ref_nbits = r["crcs"], r["block_nbits"]
            assert (r["crcs"] == ref_crcs).all() and (r["block_nbits"] == ref_nbits).all()
            assert c.stats()["n_blocks"] == len(ref_crcs)
            last = len(ref_crcs) - 1                             # the last batch stays resident for the stage dumps, on whichever lane ran it
            o = oracle.block_stages(d, last, 2, 100000)
            assert (c.stage(last, api.STAGE_BWT) == o["L"]).all() and c.stage(last, api.STAGE_ORIGPTR)[0] == o["orig_ptr"]
            assert c.decompress(exp) == d                        # the handle's own per-batch buffers (decompress side) appear on first use
            src = np.frombuffer(d, dtype=np.uint8)               # a failure in the middle of the batches comes back as a code, nothing hangs
            out = np.zeros(200_000, dtype=np.uint8)
            n = C.c_size_t(0)
            rc = c.lib.bz2b200_compress_stream(c.h, src.ctypes.data_as(C.c_void_p), len(d), out.ctypes.data_as(C.c_void_p), out.size, C.byref(n))
            assert rc == api.E_OUTPUT
            assert c.compress(d[:3_000_000]) == oracle.compress(d[:3_000_000], 2, 100000)   # and the handle works afterwards
        finally:
            c.close()
    monkeypatch.delenv("BZ2B200_LANES")


@pytest.mark.skipif(not have_ref(100000), reason="oracle/_ref not built")
def test_config1_64mb_markov_text_vs_real_reference():
    # BASELINE configs[0]: 64 MiB English-like Markov text, -9 (900k), against the reference's own code (oracle/_ref)
    d = corpus.word_text(64 << 20, 101, markov=True)
    out = api.Compressor(9, 80, 100000).compress(d)
    ref = refimpl.compress(d, 9, 100000, threads=os.cpu_count() or 8)
    assert out == ref
    assert bz2.decompress(out) == d


def test_config3_repetitive_deep_prefix_doubling():
    # BASELINE configs[2] shapes: a full all-zero block ((0,0,0,0,251)^179999, non-primitive), period-2 blocks ((ab)^449998,
    # non-primitive), period-3 (primitive at 899996), long-LCP text (100 kB segment repeated), near-periodic "ab" + noise
    base_text = corpus.word_text(100_000, 500)
    ab = np.frombuffer(b"ab" * 1_000_000, dtype=np.uint8).copy()
    ab[65535::65536] = np.frombuffer(corpus.rand_bytes(len(ab[65535::65536]), 7), dtype=np.uint8)
    d = bytes(46_000_000) + b"ab" * 1_400_000 + b"xyz" * 700_000 + base_text * 25 + ab.tobytes()
    c = api.Compressor(9, 16, 100000)
    out = c.compress(d)
    assert c.stats()["bwt_rounds"] >= 15
    exp, offs, nonprim = oracle.compress_ex(d, 9, 100000)
    assert nonprim.sum() >= 3
    assert out == exp
    assert bz2.decompress(out) == d
    if have_ref(100000):
        ref = refimpl.compress(d, 9, 100000, threads=os.cpu_count() or 8)
        assert oracle.mask_origptr(out, offs, nonprim) == oracle.mask_origptr(ref, offs, nonprim)
        assert refimpl.decompress(out, len(d) + 16, 100000) == d


def test_config5_block_size_sweep_roundtrip():
    # BASELINE configs[4] scaled: every multiplier 1..9 on text; bz2 -d and the reference decoder both accept it
    d = corpus.word_text(6 << 20, 5, markov=True)
    for mult in range(1, 10):
        out = comp(mult, 100000, 64).compress(d)
        assert out == oracle.compress(d, mult, 100000), f"-{mult}"
        assert bz2.decompress(out) == d
        if have_ref(100000) and mult in (1, 5, 9):
            assert refimpl.decompress(out, len(d) + 16, 100000) == d


def test_config4_random_bytes_full_blocks_properties():
    # BASELINE configs[3] shape at 128 MiB: incompressible bytes; size-independent properties instead of the oracle
    d = corpus.rand_bytes(128 << 20, 4)
    c = api.Compressor(9, 160, 100000)
    r = c.compress_blocks(d)
    starts = c.scan_blocks(d)
    assert len(starts) == len(r["crcs"]) and starts[0] == 0 and (np.diff(starts.astype(np.int64)) > 0).all()
    for b in (0, len(starts) // 2, len(starts) - 1):  # checksum of checksums: block CRCs are CRCs of the raw input ranges
        lo = int(starts[b]); hi = int(starts[b + 1]) if b + 1 < len(starts) else len(d)
        assert int(r["crcs"][b]) == oracle.crc32(d[lo:hi])
    full = api.stream_from_blocks(9, [(r["bits"], r["nbits"], r["crcs"])])
    assert bz2.decompress(full) == d
    assert 1.0 < len(full) / len(d) < 1.01


@pytest.mark.skipif(not have_ref(100000), reason="oracle/_ref not built")
def test_config4_random_bytes_vs_real_reference_and_sharded_over_devices():
    # random data at 900k against the reference itself on more than two blocks -- 24 MiB of the C4 generator through one
    # device and through a multi-device handle (on a one-GPU box the same device several times), and the first 96 MiB of BASELINE
    # configs[3] compared between the two paths bit for bit
    d = corpus.rand_bytes(24 << 20, 4)
    ref = refimpl.compress(d, 9, 100000, threads=os.cpu_count() or 8)
    one = api.Compressor(9, 32, 100000)
    assert one.compress(d) == ref
    ndev = api.load_library().bz2b200_device_count()
    many = api.Compressor(9, 16, 100000, devices=[i % ndev for i in range(4)])
    big = corpus.rand_bytes(96 << 20, 4)
    a, b = one.compress(big), many.compress(big)
    assert a == b
    assert bz2.decompress(b) == big
    one.close()
    many.close()


@pytest.mark.skipif(not have_ref(100000), reason="oracle/_ref not built")
def test_config2_mixed_corpus_prefix_and_config5_block_sizes_vs_real_reference():
    # bounded forms of BASELINE configs[1] and [4] inside the driver-visible suite (tools/run_configs.py runs them at full size):
    # the first 96 MiB of the mixed corpus at -9, and every block size -1..-9 on 24 MiB of the C5 text, against the reference itself
    threads = os.cpu_count() or 8
    d = corpus.mixed_corpus(96 << 20, workers=min(16, threads))
    c = api.Compressor(9, 64, 100000)
    assert c.compress(d) == refimpl.compress(d, 9, 100000, threads=threads)
    c.close()
    t = corpus.word_text(24 << 20, 5, markov=True)
    for mult in range(1, 10):
        out = comp(mult, 100000, 64).compress(t)
        assert out == refimpl.compress(t, mult, 100000, threads=threads), f"-{mult}"
        if mult in (1, 9):
            assert bz2.decompress(out) == t


def test_cpp_facade_end_to_end(tmp_path):
    # the reference-named C++ surface (OutputStream.hpp ...) driven exactly like app.cpp:103-117, both write() overloads
    exe = str(tmp_path / "facade")
    subprocess.check_call(["g++", "-std=c++17", "-O2", os.path.join(ROOT, "tests", "host", "facade_smoke.cpp"), "-o", exe,
                           "-DBZ2B200_BLOCKSIZE_DEFAULT=100000", "-I" + os.path.join(ROOT, "bzip2_opencl_b200", "include"),
                           "-I" + os.path.join(ROOT, "include"), "-L" + os.path.join(ROOT, "bzip2_opencl_b200"), "-lbz2b200"])
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "bzip2_opencl_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    d = corpus.word_text(1500000, 61) + bytes(700000) + corpus.rand_bytes(400000, 62)
    src = tmp_path / "in.bin"
    src.write_bytes(d)
    for mult, par, bulk in ((1, 3, False), (1, 64, True), (9, 2, True)):
        dst = tmp_path / f"out_{mult}_{par}.bz2"
        subprocess.check_call([exe, str(src), str(dst), str(mult), str(par)] + (["bulk"] if bulk else []), env=env)
        assert dst.read_bytes() == oracle.compress(d, mult, 100000)


# ---------------------------------------------------------------------------------------------------- decompress side (8f N1)
def test_decompress_matches_libbzip2_and_reference_decoder():
    # streams made by the CUDA compressor, by libbzip2 (17-bit codes, its own tables) and of every block size
    c = comp(9, 100000, 16)
    cases = [b"", b"x", corpus.g_runs(), corpus.word_text(2_500_000, 71), corpus.rand_bytes(1_000_003, 72), bytes(3_000_000),
             b"ab" * 500_000 + b"x", corpus.mixed_corpus(5 << 20, 1 << 20)]
    for d in cases:
        for s in (c.compress(d), bz2.compress(d, 9), bz2.compress(d, 1)):
            assert c.decompress(s) == d
            assert c.decompress(s, out_capacity=len(d)) == d      # exact capacity
    d = corpus.word_text(3_000_000, 73)
    for mult in (1, 4, 9):
        s = comp(mult, 100000, 16).compress(d)
        assert comp(mult, 100000, 16).decompress(s) == d
        if have_ref(100000):
            assert refimpl.decompress(s, len(d) + 16, 100000) == d   # and the reference decoder agrees on the same stream
    # stock base 10000 (-9 = 90 kB blocks)
    s = comp(9, 10000, 16).compress(d)
    assert comp(9, 10000, 16).decompress(s) == d


def test_decompress_long_codes_and_degenerate_tables():
    # the decoder's lookup window is 10 bits: 20-bit codes (the allocator's limit, above libbzip2's 17) go through the limit search; small
    # blocks carry tables nobody selects (all-zero counts) next to used ones.  Streams of the oracle (= the reference's own bytes).
    c = comp(9, 100000, 4)
    d = corpus.skewed_bytes(880000, 1, 48, 13, 8) + corpus.skewed_bytes(880000, 2) + corpus.skewed_bytes(880000, 7)
    assert c.decompress(oracle.compress(d, 9, 100000)) == d
    if have_ref(100000):
        assert refimpl.decompress(c.compress(d), len(d) + 16, 100000) == d
    for alpha in (2, 3, 5, 30, 256):
        for n in (1, 3, 50, 199, 200, 201, 599, 600, 601, 1199, 1200, 1201, 2399, 2400, 2401, 5000):
            raw = np.frombuffer(corpus.rand_bytes(n, 1000 * alpha + n), dtype=np.uint8).astype(np.int64)
            x = (raw % alpha).astype(np.uint8).tobytes()
            assert c.decompress(oracle.compress(x, 9, 100000)) == x, (alpha, n)
    raw = np.frombuffer(corpus.rand_bytes(150_000, 4242), dtype=np.uint8).astype(np.int64)
    x = b"".join((raw[i:i + 5000] % a).astype(np.uint8).tobytes() for i, a in zip(range(0, 150_000, 5000), [2, 3, 5, 30, 256, 4] * 5))
    c1 = comp(1, 1000, 64)                                      # 152 one-kilobyte blocks, 2 .. 6 tables side by side in the batches
    assert c1.decompress(oracle.compress(x, 1, 1000)) == x


def test_decompress_run_structured_fuzz():
    # the inverse of RLE1 on run lengths around its edges (count bytes 0, 1, 250, 251 and chunks of 255 + 4), small blocks and real ones
    for i in range(25):
        d = corpus.run_fuzz(30_000 + 4111 * i, 300 + i)
        for mult, base, par in ((1, 1000, 64), (1 + i % 9, 10000, 4)):
            assert comp(mult, base, par).decompress(oracle.compress(d, mult, base)) == d, (i, mult, base)
    d = corpus.run_fuzz(2_500_000, 78)
    assert comp(9, 100000, 4).decompress(oracle.compress(d, 9, 100000)) == d
    assert comp(9, 100000, 4).decompress(bz2.compress(d, 5)) == d


def test_decompress_device_pointers_and_batches():
    import torch
    d = corpus.mixed_corpus(12 << 20, 2 << 20)
    c = comp(9, 100000, 4)                     # 13+ blocks through batches of 4
    s = c.compress(d)
    t_in = torch.frombuffer(bytearray(s), dtype=torch.uint8).cuda()
    t_out = torch.zeros(len(d) + 64, dtype=torch.uint8, device="cuda")
    n = c.decompress_ptr(t_in.data_ptr(), len(s), t_out.data_ptr(), len(d) + 64)
    assert n == len(d) and t_out[:n].cpu().numpy().tobytes() == d
    assert c.stats()["n_blocks"] == len(c.scan_blocks(d))


def test_decompress_error_texts_are_the_references():
    c = comp(9, 100000, 16)
    d = corpus.word_text(400_000, 74)
    s = bytearray(c.compress(d))

    def err(b):
        with pytest.raises(api.Bz2Error) as e:
            c.decompress(bytes(b))
        assert e.value.code == api.E_FORMAT
        return str(e.value)

    b = bytearray(s); b[2] = ord("x")
    assert "Invalid BZip2 header" in err(b)
    b = bytearray(s); b[len(b) // 2] ^= 0x04
    assert any(t in err(b) for t in ("BZip2 block CRC error", "Error decoding  block", "block Huffman tables invalid",
                                     "BZip2 block exceeds declared block size", "BZip2 stream format error", "BZip2 start pointer invalid"))
    b = bytearray(s); b[-1] ^= 0x80; b[-3] ^= 1
    assert "BZip2 stream CRC error" in err(b)
    assert "Insufficient data" in err(s[:len(s) - 30])
    assert "BZip2 stream format error" in err(bytes(s[:4]) + b"\x00" * 40)
    with pytest.raises(api.Bz2Error) as e:       # output too small is not a format error
        c.decompress(bytes(s), out_capacity=1000)
    assert e.value.code == api.E_OUTPUT
    # a stream of a larger block size than the handle's
    with pytest.raises(api.Bz2Error) as e:
        comp(1, 100000, 16).decompress(bytes(s))
    assert e.value.code == api.E_BLOCKSIZE


def test_cpp_input_stream_facade(tmp_path):
    exe = str(tmp_path / "facade_dec")
    subprocess.check_call(["g++", "-std=c++17", "-O2", os.path.join(ROOT, "tests", "host", "facade_decompress.cpp"), "-o", exe,
                           "-DBZ2B200_BLOCKSIZE_DEFAULT=100000", "-I" + os.path.join(ROOT, "bzip2_opencl_b200", "include"),
                           "-I" + os.path.join(ROOT, "include"), "-L" + os.path.join(ROOT, "bzip2_opencl_b200"), "-lbz2b200"])
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "bzip2_opencl_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    d = corpus.word_text(1_200_000, 75) + bytes(400_000) + corpus.rand_bytes(300_000, 76)
    src = tmp_path / "in.bz2"
    src.write_bytes(bz2.compress(d, 9))
    for bulk in (False, True):
        dst = tmp_path / f"out_{bulk}"
        subprocess.check_call([exe, str(src), str(dst)] + (["bulk"] if bulk else []), env=env)
        assert dst.read_bytes() == d
    cat = tmp_path / "cat.bz2"                                 # two streams: the reference's behaviour (first only) and the opt-in for all
    cat.write_bytes(bz2.compress(d[:500_000], 1) + bz2.compress(d[500_000:], 9))
    subprocess.check_call([exe, str(cat), str(tmp_path / "first"), "bulk"], env=env)
    assert (tmp_path / "first").read_bytes() == d[:500_000]
    subprocess.check_call([exe, str(cat), str(tmp_path / "all"), "bulk", "multi"], env=env)
    assert (tmp_path / "all").read_bytes() == d
    bad = tmp_path / "bad.bz2"
    b = bytearray(src.read_bytes()); b[-2] ^= 1
    bad.write_bytes(b)
    r = subprocess.run([exe, str(bad), str(tmp_path / "x")], env=env, capture_output=True)
    assert r.returncode == 4 and b"BZip2 stream CRC error" in r.stderr


def test_cli_front_end_round_trip(tmp_path):
    # app.cpp's flags and file rules -- compress removes the input unless -k, -d strips ".bz2" (".out" if taken), -c checks
    exe = str(tmp_path / "bzip2_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", os.path.join(ROOT, "bzip2_opencl_b200", "cli", "bzip2_b200.cpp"), "-o", exe,
                           "-DBZ2B200_BLOCKSIZE_DEFAULT=100000", "-I" + os.path.join(ROOT, "bzip2_opencl_b200", "include"),
                           "-I" + os.path.join(ROOT, "include"), "-L" + os.path.join(ROOT, "bzip2_opencl_b200"), "-lbz2b200"])
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "bzip2_opencl_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    d = corpus.word_text(1_300_000, 81) + bytes(200_000) + corpus.rand_bytes(100_000, 82)
    f = tmp_path / "data.bin"
    f.write_bytes(d)
    subprocess.check_call([exe, str(f), "-s", "5", "-p", "4", "-k"], env=env)
    z = tmp_path / "data.bin.bz2"
    assert f.exists() and z.read_bytes() == oracle.compress(d, 5, 100000)
    r = subprocess.run([exe, str(z), "--check"], env=env, capture_output=True)
    assert r.returncode == 0 and b"Integrity check passed!" in r.stdout and z.exists()
    r = subprocess.run([exe, str(z), "-d", "-k"], env=env, capture_output=True)      # data.bin exists -> data.bin.out
    assert r.returncode == 0 and b"Warning: Output file exists" in r.stderr
    assert (tmp_path / "data.bin.out").read_bytes() == d and z.exists()
    f.unlink()
    subprocess.check_call([exe, str(z), "--dec"], env=env)                            # input removed, name restored
    assert f.read_bytes() == d and not z.exists()
    subprocess.check_call([exe, str(f)], env=env)                                     # default -9, input removed
    assert not f.exists() and bz2.decompress(z.read_bytes()) == d
    two = tmp_path / "two.bin.bz2"                                                   # concatenated streams: first one only, or all with -a
    two.write_bytes(z.read_bytes() + bz2.compress(b"tail" * 1000, 1))
    subprocess.check_call([exe, str(two), "-d", "-k"], env=env)
    assert (tmp_path / "two.bin").read_bytes() == d
    (tmp_path / "two.bin").unlink()
    subprocess.check_call([exe, str(two), "-d", "-k", "-a"], env=env)
    assert (tmp_path / "two.bin").read_bytes() == d + b"tail" * 1000
    r = subprocess.run([exe, str(two), "-c", "--all-streams"], env=env, capture_output=True)
    assert r.returncode == 0 and b"Integrity check passed!" in r.stdout
    b = bytearray(z.read_bytes()); b[len(b) // 2] ^= 1
    z.write_bytes(b)
    r = subprocess.run([exe, str(z), "-c"], env=env, capture_output=True)
    assert r.returncode == 2 and b"BZip2" in r.stderr or b"block" in r.stderr


# ------------------------------------------------------------------------------------------- on-disk framing extras (8f N4)
def test_concatenated_streams_decode_like_bzip2():
    # `cat a.bz2 b.bz2 | bzip2 -d`: streams of different block sizes and makers, an empty stream in between, trailing bytes ignored
    c = comp(9, 100000, 8)
    a, b, e = corpus.word_text(1_900_000, 91), corpus.rand_bytes(950_000, 92) + bytes(70_000), corpus.mixed_corpus(3 << 20, 1 << 20)
    parts = [c.compress(a), bz2.compress(b, 3), c.compress(b""), comp(2, 100000, 8).compress(e), bz2.compress(a[:10], 1)]
    cat = b"".join(parts)
    want = a + b + e + a[:10]
    assert bz2.decompress(cat) == want                      # libbzip2's own multi-stream reading agrees on what the file means
    assert c.decompress(cat, multi=True) == want and c.last_streams == 5
    assert c.decompress(cat) == a                            # the reference's InputStream stops behind the first stream (InputStream.hpp:146-153)
    assert c.decompress(cat + b"\x00garbage", multi=True) == want and c.last_streams == 5
    assert c.decompress(parts[0], multi=True) == a and c.last_streams == 1
    bad = bytearray(cat); bad[len(parts[0]) + len(parts[1]) // 2] ^= 0x10      # damage inside the second stream fails the call
    with pytest.raises(api.Bz2Error) as ex:
        c.decompress(bytes(bad), multi=True)
    assert ex.value.code == api.E_FORMAT
    with pytest.raises(api.Bz2Error) as ex:                  # a later stream with larger blocks than the handle allows
        comp(2, 100000, 8).decompress(parts[3] + parts[0], multi=True)
    assert ex.value.code == api.E_BLOCKSIZE


def test_bzip2_cli_accepts_every_block_size(tmp_path):
    # `bzip2 -t` / `bzip2 -dc` sweep: header digit '0'+mult (OutputStream.hpp:128) with blocks of mult x base bytes.  With the stock
    # base 10000 the digit promises ten times the real block size, which every decoder accepts (blocks may be shorter than declared).
    import shutil
    exe = shutil.which("bzip2")
    if exe is None:
        pytest.skip("no bzip2 binary in this image")
    d = corpus.word_text(700_000, 93) + bytes(300_000) + corpus.rand_bytes(250_000, 94) + b"ab" * 200_000
    for base in (100000, 10000):
        for mult in range(1, 10):
            s = comp(mult, base, 16).compress(d)
            assert s[:4] == b"BZh" + bytes([ord("0") + mult])
            f = tmp_path / f"s_{base}_{mult}.bz2"
            f.write_bytes(s)
            assert subprocess.run([exe, "-t", str(f)], capture_output=True).returncode == 0
            if mult in (1, 9):
                assert subprocess.run([exe, "-dc", str(f)], capture_output=True, check=True).stdout == d
    cat = tmp_path / "cat.bz2"                               # multi-stream file written by two compressors
    cat.write_bytes(comp(9, 100000, 16).compress(d) + comp(1, 100000, 16).compress(d[::-1]))
    assert subprocess.run([exe, "-dc", str(cat)], capture_output=True, check=True).stdout == d + d[::-1]
