"""The oracle (oracle/bz2_oracle.c) is pinned here: golden hashes of SURVEY Appendix E, the real reference built for the
CPU (oracle/_ref), and two independent decoders (Python's bz2 = libbzip2, and the reference's own InputStream)."""
import bz2
import hashlib
import json
import os

import numpy as np
import pytest

from bzip2_opencl_b200 import corpus
from oracle import oracle, refimpl
from conftest import have_ref

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "appendix_e.json")))
CONFIGS = {"A": (10000, 9), "B": (100000, 9), "C": (100000, 1)}
NAMES = [k for k in GOLD if not k.startswith("_")]


@pytest.mark.parametrize("name", NAMES)
def test_corpus_generators_are_pinned(name):
    d = corpus.GOLDEN_INPUTS[name]()
    assert len(d) == GOLD[name]["bytes"]
    assert hashlib.sha256(d).hexdigest()[:16] == GOLD[name]["sha"]


@pytest.mark.parametrize("cfg", list(CONFIGS))
@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_golden(name, cfg):
    base, mult = CONFIGS[cfg]
    d = corpus.GOLDEN_INPUTS[name]()
    out, offs, nonprim = oracle.compress_ex(d, mult, base)
    size, sha = GOLD[name][cfg]
    assert len(out) == size
    if not nonprim.any():
        assert hashlib.sha256(out).hexdigest()[:16] == sha  # bit-identical to the reference's output
    elif have_ref(base):
        # non-primitive blocks: only the 24-bit origPtr may differ
        ref = refimpl.compress(d, mult, base, threads=4)
        assert hashlib.sha256(ref).hexdigest()[:16] == sha
        assert oracle.mask_origptr(out, offs, nonprim) == oracle.mask_origptr(ref, offs, nonprim)
    if base == 100000:
        assert bz2.decompress(out) == d
    if have_ref(base):
        assert refimpl.decompress(out, len(d) + 16, base) == d


def _seeded_cases():
    rng = np.random.default_rng(7)
    cases = []
    for i in range(10):
        n = int(rng.integers(1, 400000))
        kind = i % 5
        if kind == 0:
            d = rng.integers(0, 256, n, dtype=np.uint8)
        elif kind == 1:
            d = rng.integers(0, 4, n, dtype=np.uint8)
        elif kind == 2:
            d = np.repeat(rng.integers(0, 256, n // 50 + 1, dtype=np.uint8), rng.integers(1, 400, n // 50 + 1))[:n]
        elif kind == 3:
            d = np.frombuffer(corpus.word_text(n, 1000 + i), dtype=np.uint8)
        else:
            d = np.frombuffer(corpus.binary_records(n, 2000 + i), dtype=np.uint8)
        cases.append(d.tobytes())
    return cases


@pytest.mark.skipif(not have_ref(100000) or not have_ref(10000), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("idx", range(10))
def test_oracle_equals_real_reference_on_seeded_inputs(idx):
    d = _seeded_cases()[idx]
    for base, mult in ((100000, 9), (100000, 1 + idx % 9), (10000, 9), (10000, 1 + (idx * 3) % 9)):
        out, offs, nonprim = oracle.compress_ex(d, mult, base)
        ref = refimpl.compress(d, mult, base, threads=2)
        assert oracle.mask_origptr(out, offs, nonprim) == oracle.mask_origptr(ref, offs, nonprim)
        if not nonprim.any():
            assert out == ref
        s1, l1, c1 = oracle.rle1_blocks(d, mult, base)
        s2, l2, c2 = refimpl.rle1_blocks(d, mult, base)
        assert (s1 == s2).all() and (l1 == l2).all() and (c1 == c2).all()


def alphabet_sweep_input() -> bytes:
    # the input of tests/test_gpu_parity.py::test_alphabet_sizes_and_mixed_batches (stretches over 1 .. 256 used byte values)
    rng = np.random.default_rng(20261018)
    parts = []
    for alpha in (1, 2, 3, 4, 5, 8, 9, 16, 17, 32, 33, 64, 65, 128, 129, 200, 256):
        vals = np.sort(rng.choice(256, size=alpha, replace=False)).astype(np.uint8)
        body = vals[rng.integers(0, alpha, size=23_000)]
        body[5000:5600] = body[4000:4600]
        parts.append(body.tobytes())
    return b"".join(parts)


@pytest.mark.skipif(not have_ref(100000) or not have_ref(10000), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_equals_real_reference_on_the_alphabet_sweep():
    # what the GPU path is compared with on that input is the oracle: pin the oracle to the reference's own code there as well
    d = alphabet_sweep_input()
    for mult, base in ((2, 10000), (1, 100000), (9, 100000)):
        out, offs, nonprim = oracle.compress_ex(d, mult, base)
        ref = refimpl.compress(d, mult, base, threads=2)
        assert oracle.mask_origptr(out, offs, nonprim) == oracle.mask_origptr(ref, offs, nonprim)
        if not nonprim.any():
            assert out == ref
        if base == 100000:
            assert bz2.decompress(out) == d


def small_block_cases():
    # one small block each: MTF lengths on both sides of the table-count thresholds (selectTableCount, kernel.cpp:2808-2818: 200 / 600 /
    # 1200 / 2400 symbols) for alphabets of 2 .. 256 values -- with few symbols most of the 2 .. 6 tables end up without a single group
    # (all-zero counts).  tests/test_gpu_parity.py builds the same list.
    cases = []
    for alpha in (2, 3, 5, 30, 256):
        for n in (1, 2, 3, 49, 50, 51, 99, 100, 101, 150, 197, 198, 199, 200, 201, 202, 597, 598, 599, 600, 601, 1197, 1198, 1199, 1200, 1201,
                  2397, 2398, 2399, 2400, 2401, 2402, 3000, 5000):
            raw = np.frombuffer(corpus.rand_bytes(n, 1000 * alpha + n), dtype=np.uint8).astype(np.int64)
            cases.append((raw % alpha).astype(np.uint8).tobytes())
    return cases


@pytest.mark.skipif(not have_ref(100000), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_equals_real_reference_around_the_table_count_thresholds():
    seen, lengths = set(), set()
    for d in small_block_cases():
        assert oracle.compress(d, 9, 100000) == refimpl.compress(d, 9, 100000)
        st = oracle.block_stages(d, 0, 9, 100000)
        seen.add((st["n_tables"], len(set(st["selectors"].tolist()))))
        lengths.add(st["mtf_length"])
    assert {2, 3, 4, 5, 6} <= {t for t, _ in seen}
    assert any(used < t for t, used in seen if t == 6) and any(used == 1 for _, used in seen)      # tables nobody selects
    for edge in (200, 600, 1200, 2400):
        assert {edge - 1, edge, edge + 1} <= lengths


@pytest.mark.skipif(not have_ref(10000) or not have_ref(100000), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_equals_real_reference_on_run_structured_fuzz():
    # corpus.run_fuzz: run lengths on the RLE1 edges, small alphabets, short periods, copies -- block ends in every phase of a run
    # (tests/test_gpu_parity.py::test_run_structured_fuzz compares the CUDA path with the oracle on the same generator)
    blocks = 0
    for i in range(40):
        d = corpus.run_fuzz(60000 + 3001 * i, 9000 + i)
        for mult, base in ((1, 10000), (1 + i % 9, 10000), (1, 100000)):
            out, offs, nonprim = oracle.compress_ex(d, mult, base)
            ref = refimpl.compress(d, mult, base, threads=2)
            assert oracle.mask_origptr(out, offs, nonprim) == oracle.mask_origptr(ref, offs, nonprim), (i, mult, base)
            s1, l1, c1 = oracle.rle1_blocks(d, mult, base)
            s2, l2, c2 = refimpl.rle1_blocks(d, mult, base)
            assert (s1 == s2).all() and (l1 == l2).all() and (c1 == c2).all()
            blocks += len(offs)
    assert blocks > 350


def length_limit_input() -> bytes:
    # three 900k blocks whose MTF symbol counts fall off like Fibonacci numbers (tests/test_gpu_parity.py uses the same input)
    return corpus.skewed_bytes(880000, 1, 48, 13, 8) + corpus.skewed_bytes(880000, 2) + corpus.skewed_bytes(880000, 7)


def _unrestricted_depth(freqs) -> int:
    import heapq
    h = [(int(f), i, 0) for i, f in enumerate(freqs)]
    heapq.heapify(h)
    c = len(h)
    while len(h) > 1:
        a, b = heapq.heappop(h), heapq.heappop(h)
        heapq.heappush(h, (a[0] + b[0], c, max(a[2], b[2]) + 1))
        c += 1
    return h[0][2]


def test_length_limit_input_really_hits_the_20_bit_limit():
    # the allocator caps code lengths at 20 (HUFFMAN_ENCODE_MAXIMUM_CODE_LENGTH, kernel.cpp:2748-2805).  On this input
    # a plain Huffman code over a table's counts would be 21-22 bits deep, the coded lengths stop at 20, and the codes stay complete.
    d = length_limit_input()
    limited = 0
    for b in range(3):
        st = oracle.block_stages(d, b, 9, 100000)
        a, mtf, sel, cl = st["alphabet_size"], st["mtf"], st["selectors"], st["code_lengths"]
        for t in range(st["n_tables"]):
            f = np.zeros(a, dtype=np.int64)
            for g in np.nonzero(sel == t)[0]:
                np.add.at(f, mtf[g * 50:(g + 1) * 50], 1)
            lens = cl[t, :a].astype(np.int64)
            assert lens.max() <= 20 and lens.min() >= 1
            assert sum(1 << (20 - int(l)) for l in lens) == 1 << 20          # Kraft equality: a complete prefix code
            if _unrestricted_depth(f) > 20:
                limited += 1
                assert lens.max() == 20
    assert limited >= 3
    assert bz2.decompress(oracle.compress(d, 9, 100000)) == d


@pytest.mark.skipif(not have_ref(100000) or not have_ref(10000), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_equals_real_reference_where_code_lengths_are_limited():
    d = length_limit_input()
    for mult, base in ((9, 100000), (9, 10000)):
        out, offs, nonprim = oracle.compress_ex(d, mult, base)
        ref = refimpl.compress(d, mult, base, threads=4)
        assert not nonprim.any() and out == ref
    for b in range(3):
        o, r = oracle.block_stages(d, b, 9, 100000), refimpl.block_stages(d, b, 9, 100000)
        assert (o["mtf"] == r["mtf"]).all() and (o["freq"] == r["freq"]).all() and (o["selectors"] == r["selectors"]).all()


@pytest.mark.skipif(not have_ref(100000), reason="oracle/_ref not built")
def test_oracle_stage_dumps_equal_reference():
    d = corpus.word_text(250000, 77) + bytes(5000) + corpus.rand_bytes(40000, 3)
    for b in range(3):
        o = oracle.block_stages(d, b, 1, 100000)
        r = refimpl.block_stages(d, b, 1, 100000)
        assert o["length"] == r["length"] and o["orig_ptr"] == r["orig_ptr"] and o["crc"] == r["crc"]
        assert (o["rle"] == r["rle"]).all() and (o["L"] == r["L"]).all() and (o["present"] == r["present"]).all()
        assert (o["mtf"] == r["mtf"]).all() and (o["freq"] == r["freq"]).all() and (o["selectors"] == r["selectors"]).all()


def test_oracle_block_boundary_edges():
    # blocks of exactly bs, bs-1 ... bytes; refused byte starts the next block with a fresh run
    for base in (1000, 4000):
        bs = base
        for n in (bs - 7, bs - 6, bs - 5, bs - 4, bs, bs + 1, 2 * bs):
            d = bytes((i * 7 + 3) % 251 for i in range(n))
            starts, lens, _ = oracle.rle1_blocks(d, 1, base)
            assert int(lens.max()) <= bs
            assert starts[0] == 0
            out = oracle.compress(d, 1, base)
            if base == 100000:
                assert bz2.decompress(out) == d
        z = bytes(5 * bs)
        starts, lens, _ = oracle.rle1_blocks(z, 1, base)
        assert int(lens.max()) <= bs and len(starts) >= 1
