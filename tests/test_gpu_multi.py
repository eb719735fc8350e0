"""Round-2 GPU tests: one stream over several handles / devices, the reference's unmodified app.cpp on the facade,
decoder hardening against crafted streams.  All calls go through the C ABI; the oracle is the checker."""
import bz2
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from bzip2_opencl_b200 import api, corpus, multi
from oracle import oracle, refimpl
from conftest import have_ref, ROOT

pytestmark = pytest.mark.gpu


def _devices(n):
    """n device ids: the box's GPUs round robin (a multi handle may name a device more than once, so one GPU covers the logic)."""
    have = api.load_library().bz2b200_device_count()
    return [i % have for i in range(n)]


def _runs_and_text(n, seed):
    parts, k = [], 0
    while sum(len(p) for p in parts) < n:
        parts.append(corpus.word_text(37000 + 991 * k, seed + k))
        parts.append(bytes([k % 3]) * (2000 + 5000 * (k % 4)))          # runs: blocks that hold far more raw bytes than bs
        parts.append(corpus.rand_bytes(21000, seed + 7 * k))
        parts.append(b"ab" * 3000)
        k += 1
    return b"".join(parts)[:n]


def test_slice_primitives_reproduce_the_stream():
    """the process-per-GPU flow on one GPU: slices of one stream through bz2b200_scan_prepare / _scan_blocks_ex / _compress_scanned /
    _shift_bits (one handle per 'rank'), assembled as multi.py does"""
    import torch
    for mult, base, n in ((1, 1000, 300_000), (3, 10000, 4_000_000), (9, 100000, 60_000_000)):
        d = _runs_and_text(n, 5) if n < 10_000_000 else corpus.mixed_corpus(n, 1 << 22)
        exp = refimpl.compress(d, mult, base, threads=os.cpu_count() or 8) if (n > 10_000_000 and have_ref(base)) else oracle.compress(d, mult, base)
        bs = mult * base
        src = np.frombuffer(d, dtype=np.uint8)
        for world in (2, 3):
            plan = multi.plan_slices(len(d), world, bs, ratio=0.8, min_slice=2 * bs)
            handles = [api.Compressor(mult, 8, base) for _ in range(world)]
            for r, (b0, lo, hi) in enumerate(plan):                     # every rank prepares its slice first (asynchronously)
                handles[r].scan_prepare(src.ctypes.data + b0, hi - b0)
            nxt = {0: 0}
            res = []
            for r, (b0, lo, hi) in enumerate(plan):
                c = handles[r]
                box = {}

                def scan(first_rel, final, c=c, b0=b0, hi=hi, box=box):
                    nb, used = c.scan_blocks_ex(src.ctypes.data + b0, hi - b0, final, first_rel, want_starts=False)
                    box["nb"] = nb
                    return nb, used
                s, nb, nx = multi.chain_walk(scan, plan[r], r, world, True, lambda r=r: nxt[r], lambda v, r=r: nxt.__setitem__(r + 1, v))
                if nb:
                    cap = c.max_output_bytes(hi - b0) + 64
                    out = torch.zeros(cap, dtype=torch.uint8, device="cuda")
                    q = c.compress_scanned_ptr(out.data_ptr(), cap, nb)
                    e = (q["nbits"] - 1) >> 3
                    tail = multi.tail_of(bytes(out[max(0, e - 1):e + 1].cpu().numpy()), q["nbits"])
                    res.append((q["nbits"], out, q["crcs"], tail))
                else:
                    res.append((0, None, [], 0))
            nbits_all = [x[0] for x in res]
            placement = multi.place_parts(nbits_all, [x[3] for x in res], 32)
            shifted = []
            for r, (x, pl) in enumerate(zip(res, placement)):
                if not x[0]:
                    shifted.append(b"")
                    continue
                n_dst = (pl[1] + x[0] + 7) // 8
                dst = torch.empty(n_dst + 16, dtype=torch.uint8, device="cuda")
                handles[r].shift_bits_ptr(x[1].data_ptr(), x[0], pl[1], pl[2], dst.data_ptr(), n_dst + 16)
                shifted.append(bytes(dst[:n_dst].cpu().numpy()))
                assert shifted[-1] == multi.shift_bits_host(bytes(x[1][:(x[0] + 7) // 8].cpu().numpy()), x[0], pl[1], pl[2])
            got = multi.assemble_stream(mult, shifted, placement, nbits_all, multi.fold_crcs([x[2] for x in res]))
            assert got == exp, (mult, base, world)
            for c in handles:
                c.close()


def test_compress_blocks_from_hands_the_chain_on_early():
    """bz2b200_compress_blocks_from: scan + compress of a slice in one call; the callback delivers the next slice's start from inside the
    call (before the blocks are compressed); prepared and unprepared slices, host and device memory"""
    import torch
    for mult, base, n in ((2, 10000, 5_000_000), (9, 100000, 40_000_000)):
        d = _runs_and_text(n, 31) if n < 10_000_000 else corpus.mixed_corpus(n, 1 << 22)
        exp = refimpl.compress(d, mult, base, threads=os.cpu_count() or 8) if (n > 10_000_000 and have_ref(base)) else oracle.compress(d, mult, base)
        bs = mult * base
        src = np.frombuffer(d, dtype=np.uint8)
        dev = torch.frombuffer(bytearray(d), dtype=torch.uint8).cuda()
        for world, on_device in ((3, False), (2, True)):
            plan = multi.plan_slices(len(d), world, bs, ratio=0.85, min_slice=2 * bs)
            handles = [api.Compressor(mult, 8, base) for _ in range(world)]
            ptr0 = dev.data_ptr() if on_device else src.ctypes.data
            for r, (b0, lo, hi) in enumerate(plan):
                if r > 0 and r % 2 == 1:                       # some slices are prepared, some are not
                    handles[r].scan_prepare(ptr0 + b0, hi - b0)
            s, res = 0, []
            for r, (b0, lo, hi) in enumerate(plan):
                c = handles[r]
                got = {}
                cap = c.max_output_bytes(hi - b0) + 64
                out = torch.zeros(cap, dtype=torch.uint8, device="cuda")
                order = []

                def cb(nxt, got=got, order=order):
                    got["next"] = nxt
                    order.append("walked")
                assert b0 <= s < hi
                q = c.compress_blocks_from_ptr(ptr0 + b0, hi - b0, out.data_ptr(), cap, r == world - 1, s - b0, cb)
                order.append("returned")
                assert order == ["walked", "returned"] and got["next"] == q["consumed"]
                e = (q["nbits"] - 1) >> 3
                res.append((q["nbits"], bytes(out[:(q["nbits"] + 7) // 8].cpu().numpy()), q["crcs"],
                            multi.tail_of(bytes(out[max(0, e - 1):e + 1].cpu().numpy()), q["nbits"])))
                s = b0 + got["next"]
            assert s == len(d)
            nbits_all = [x[0] for x in res]
            placement = multi.place_parts(nbits_all, [x[3] for x in res], 32)
            shifted = [multi.shift_bits_host(x[1], x[0], pl[1], pl[2]) for x, pl in zip(res, placement)]
            assert multi.assemble_stream(mult, shifted, placement, nbits_all, multi.fold_crcs([x[2] for x in res])) == exp, (mult, base, world)
            for c in handles:
                c.close()


def test_scan_blocks_ex_from_an_offset_and_not_final():
    d = _runs_and_text(2_500_000, 9)
    mult, base = 2, 10000
    c = api.Compressor(mult, 8, base)
    starts = oracle.rle1_blocks(d, mult, base)[0]
    src = np.frombuffer(d, dtype=np.uint8)
    # not final: the last block (its refusing byte has not been seen) is left to the caller
    got, used = c.scan_blocks_ex(src.ctypes.data, len(d), False, 0)
    assert (got == starts[:-1]).all() and used == int(starts[-1])
    # from a block start in the middle, with the bytes in front as context only
    k = len(starts) // 2
    got, used = c.scan_blocks_ex(src.ctypes.data, len(d), True, int(starts[k]))
    assert (got == starts[k:]).all() and used == len(d)
    # a slice that begins at an arbitrary earlier byte
    off = int(starts[k]) - 12345
    got, used = c.scan_blocks_ex(src.ctypes.data + off, len(d) - off, True, 12345)
    assert (got + off == starts[k:]).all()
    # the block table is good for exactly one compress call
    cap = c.max_output_bytes(len(d)) + 64
    out = np.zeros(cap, dtype=np.uint8)
    q = c.compress_scanned_ptr(out.ctypes.data, cap, len(got))
    assert len(q["crcs"]) == len(starts) - k
    assert api.stream_from_blocks(mult, [(out.tobytes(), q["nbits"], q["crcs"])]) == oracle.compress(d[int(starts[k]):], mult, base)
    with pytest.raises(api.Bz2Error):
        c.compress_scanned_ptr(out.ctypes.data, cap, len(got))
    c.close()


@pytest.mark.parametrize("ndev", [2, 3, 8])
def test_multi_device_handle_equals_single_handle(ndev, monkeypatch):
    monkeypatch.setenv("BZ2B200_MULTI_RATIO", "0.9")
    for mult, base, n in ((1, 1000, 700_000), (2, 10000, 6_000_000)):
        d = _runs_and_text(n, 21 + ndev)
        exp = oracle.compress(d, mult, base)
        m = api.Compressor(mult, 6, base, devices=_devices(ndev))
        assert m.lib.bz2b200_device_count_of(m.h) == ndev
        assert m.compress(d) == exp
        assert m.stats()["n_blocks"] == len(oracle.rle1_blocks(d, mult, base)[0]) and m.stats()["kernel_launches"] > 0
        assert (m.scan_blocks(d) == oracle.rle1_blocks(d, mult, base)[0]).all()
        # streaming calls: not final + carry bits, through the multi handle
        cut = len(d) * 2 // 3
        a = m.compress_blocks(d[:cut], is_final=False)
        b = m.compress_blocks(d[a["consumed"]:], is_final=True, start_bit=a["nbits"] % 8)
        first = np.frombuffer(a["bits"], dtype=np.uint8).copy()
        second = np.frombuffer(b["bits"], dtype=np.uint8).copy()
        if a["nbits"] % 8:
            second[0] |= first[-1]
            first = first[:-1]
        bits = first.tobytes() + second.tobytes()
        assert api.stream_from_blocks(mult, [(bits, a["nbits"] + b["nbits"], list(a["crcs"]) + list(b["crcs"]))]) == exp
        assert m.decompress(exp) == d                               # the decompress side runs on the first device
        m.close()


def test_multi_device_handle_long_runs_and_tiny_inputs():
    mult, base = 1, 1000
    m = api.Compressor(mult, 6, base, devices=_devices(4))
    for d in (bytes(3_000_000), b"\x07" * 1_000_001 + b"xyz" + bytes(900_000), b"", b"a", corpus.word_text(5000, 1),
              bytes(200_000) + corpus.word_text(200_000, 3) + bytes(300_000)):
        assert m.compress(d) == oracle.compress(d, mult, base), len(d)
    m.close()


def test_multi_device_handle_on_run_structured_fuzz(monkeypatch):
    # corpus.run_fuzz through the slice chain: slice edges (multiples of 4096) fall into runs, count bytes and period stretches, and every
    # device starts its walk from the state the prefix tables give it at the position the device in front handed on
    for ratio in ("0.95", "0.6"):
        monkeypatch.setenv("BZ2B200_MULTI_RATIO", ratio)
        for ndev in (2, 5):
            m = api.Compressor(1, 6, 1000, devices=_devices(ndev))
            for i in range(12):
                d = corpus.run_fuzz(90_000 + 7919 * i, 500 + i)
                assert m.compress(d) == oracle.compress(d, 1, 1000), (ratio, ndev, i)
                assert (m.scan_blocks(d) == oracle.rle1_blocks(d, 1, 1000)[0]).all()
            m.close()


def test_multi_device_handle_900k_blocks_device_pointers():
    import torch
    d = corpus.mixed_corpus(96 << 20, 1 << 22)
    single = api.Compressor(9, 64, 100000)
    exp = single.compress(d)
    single.close()
    assert bz2.decompress(exp) == d
    m = api.Compressor(9, 32, 100000, devices=_devices(3))
    assert m.compress(d) == exp                                       # pageable host input
    t_in = torch.frombuffer(bytearray(d), dtype=torch.uint8).cuda()
    cap = m.max_output_bytes(len(d)) + 64
    t_out = torch.zeros(cap, dtype=torch.uint8, device="cuda")
    r = m.compress_blocks_ptr(t_in.data_ptr(), len(d), t_out.data_ptr(), cap)   # device in, device out
    bits = t_out[:(r["nbits"] + 7) // 8].cpu().numpy().tobytes()
    assert api.stream_from_blocks(9, [(bits, r["nbits"], r["crcs"])]) == exp
    pin = torch.empty(len(d), dtype=torch.uint8, pin_memory=True)
    pin.numpy()[:] = np.frombuffer(d, dtype=np.uint8)
    out = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
    r = m.compress_blocks_ptr(pin.data_ptr(), len(d), out.data_ptr(), cap)      # pinned host in and out
    assert api.stream_from_blocks(9, [(out[:(r["nbits"] + 7) // 8].numpy().tobytes(), r["nbits"], r["crcs"])]) == exp
    m.close()


def _app(base):
    return os.path.join(ROOT, "oracle", "_ref", f"app_b200_{base}")


@pytest.mark.skipif(not os.path.exists(_app(100000)), reason="oracle/_ref/app_b200_* not built (needs /root/reference at build time)")
@pytest.mark.parametrize("devices", ["0", "0,0,0"])
def test_unmodified_reference_app_runs_on_the_facade(tmp_path, devices):
    """/root/reference/app.cpp, compiled as it is against bzip2_opencl_b200/include (oracle/build_ref.py build_app): its per-byte
    compress loop (app.cpp:105-116), -d and -c on the CUDA path; BZ2B200_DEVICES=0,0,0 drives the multi-device handle through it"""
    env = dict(os.environ, BZ2B200_DEVICES=devices, BZ2B200_FACADE_BLOCKS="12")
    d = corpus.word_text(2_000_000, 77) + bytes(3_000_000) + corpus.rand_bytes(1_500_000, 78) + bytes([5]) * 64_000_000 + corpus.word_text(500_000, 79)
    for base, mult in ((100000, 9), (10000, 2)):
        data = d if base == 100000 else d[:9_000_000]
        f = tmp_path / f"in_{base}.bin"
        f.write_bytes(data)
        subprocess.run([_app(base), str(f), "-k", "-s", str(mult), "-p", "3"], env=env, check=True, timeout=600)
        out = (tmp_path / f"in_{base}.bin.bz2").read_bytes()
        exp = oracle.compress(data, mult, base)    # (the run blocks are non-primitive: the oracle's origPtr convention)
        assert out == exp
        r = subprocess.run([_app(base), str(tmp_path / f"in_{base}.bin.bz2"), "-c"], env=env, capture_output=True, timeout=600)
        assert r.returncode == 0 and b"Integrity check passed!" in r.stdout
        os.rename(f, tmp_path / "orig")
        subprocess.run([_app(base), str(tmp_path / f"in_{base}.bin.bz2"), "-d", "-k"], env=env, check=True, timeout=600)
        assert f.read_bytes() == data
        os.remove(f)
        os.remove(tmp_path / f"in_{base}.bin.bz2")


def test_facade_zero_runs_written_byte_by_byte(tmp_path):
    """a batch that does not complete a block (long zero runs at a small -p) must not turn into one device call per byte"""
    exe = str(tmp_path / "facade")
    subprocess.check_call(["g++", "-std=c++17", "-O2", os.path.join(ROOT, "tests", "host", "facade_smoke.cpp"), "-o", exe,
                           "-DBZ2B200_BLOCKSIZE_DEFAULT=100000", "-I" + os.path.join(ROOT, "bzip2_opencl_b200", "include"),
                           "-I" + os.path.join(ROOT, "include"), "-L" + os.path.join(ROOT, "bzip2_opencl_b200"), "-lbz2b200", "-pthread"])
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "bzip2_opencl_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""),
               BZ2B200_FACADE_BLOCKS="1")
    d = bytes(120_000_000) + b"tail"
    f = tmp_path / "zeros.bin"
    f.write_bytes(d)
    subprocess.run([exe, str(f), str(tmp_path / "zeros.bz2"), "9", "1"], env=env, check=True, timeout=300)
    assert bz2.decompress((tmp_path / "zeros.bz2").read_bytes()) == d
    assert (tmp_path / "zeros.bz2").read_bytes() == oracle.compress(d, 9, 100000)


def test_decompress_reports_the_size_it_needs():
    """an output buffer that is too small costs ONE more call, not a doubling loop -- BZ2B200_E_OUTPUT carries the number of bytes
    the stream decodes to (the decoder goes on without writing)"""
    d = bytes(30_000_000) + corpus.word_text(2_000_000, 3) + bytes(10_000_000)
    c = api.Compressor(9, 8, 100000)
    stream = c.compress(d) + c.compress(b"second stream " * 1000)
    src = np.frombuffer(stream, dtype=np.uint8)
    for cap in (0, 1000, 29_000_000, len(d) - 1):
        out = np.zeros(max(cap, 1), dtype=np.uint8)
        n = C.c_size_t(0)
        rc = c.lib.bz2b200_decompress_stream(c.h, src.ctypes.data_as(C.c_void_p), len(stream), out.ctypes.data_as(C.c_void_p), cap, C.byref(n))
        assert rc == api.E_OUTPUT and n.value == len(d), (cap, n.value)
    ns = C.c_uint64(0)
    n = C.c_size_t(0)
    out = np.zeros(100, dtype=np.uint8)
    rc = c.lib.bz2b200_decompress_concat(c.h, src.ctypes.data_as(C.c_void_p), len(stream), out.ctypes.data_as(C.c_void_p), 100, C.byref(n), C.byref(ns))
    assert rc == api.E_OUTPUT and n.value == len(d) + 14000
    assert c.decompress(stream) == d                                  # (the binding retries with the reported size)
    assert c.decompress(stream, multi=True) == d + b"second stream " * 1000
    c.close()


# ---------------------------------------------------------------------------------------------------------------- crafted streams
class _Bits:
    def __init__(self):
        self.b = []

    def put(self, n, v):
        for i in range(n - 1, -1, -1):
            self.b.append((v >> i) & 1)

    def bytes(self):
        pad = (-len(self.b)) % 8
        return np.packbits(np.array(self.b + [0] * pad, dtype=np.uint8)).tobytes()


def _crafted(lengths, nsel=1, selectors=None, ntab=2, digit=1, nsyms_used=2, body_bits=64):
    """a one-block stream with hand-made tables: `lengths` for every table and symbol (alphabet = nsyms_used + 2)"""
    w = _Bits()
    w.put(24, 0x425A68); w.put(8, ord("0") + digit)
    w.put(24, 0x314159); w.put(24, 0x265359); w.put(32, 0x12345678); w.put(1, 0); w.put(24, 0)
    ranges = (nsyms_used + 15) // 16          # byte values 0 .. nsyms_used-1 are in use
    w.put(16, (0xffff << (16 - ranges)) & 0xffff)
    for r in range(ranges):
        k = min(16, nsyms_used - 16 * r)
        w.put(16, (0xffff << (16 - k)) & 0xffff)
    w.put(3, ntab); w.put(15, nsel)
    for i in range(nsel):
        s = selectors[i] if selectors else 0
        w.put(s + 1, ((1 << s) - 1) << 1)   # unary
    alpha = nsyms_used + 2
    for t in range(ntab):
        w.put(5, lengths[0])
        cur = lengths[0]
        for j in range(alpha):
            while cur < lengths[j]:
                w.put(2, 2); cur += 1
            while cur > lengths[j]:
                w.put(2, 3); cur -= 1
            w.put(1, 0)
    for _ in range(body_bits):
        w.put(1, 1)
    w.put(24, 0x177245); w.put(24, 0x385090); w.put(32, 0)
    return w.bytes()


def test_crafted_tables_never_reach_outside_the_decoder_tables():
    """over-subscribed code lengths (Kraft sum > 1), selectors naming a table the block does not define and selector counts
    above the handle's slot must end in a format error (or decode like the reference), never in a device fault -- and the handle
    must stay usable"""
    good = corpus.word_text(300_000, 5)
    for mult in (1, 9):
        c = api.Compressor(mult, 4, 100000)
        ok = c.compress(good)
        for stream in (
            _crafted([1] * 258, nsyms_used=256),                       # 258 codes of length 1
            _crafted([2] * 40, nsyms_used=38),                         # 40 codes of length 2
            _crafted([1, 1, 1, 1], nsyms_used=2),
            _crafted([3, 3, 3, 3], nsel=3, selectors=[0, 1, 2], ntab=2),   # MTF index 2 -> table 2 of 2
            _crafted([3, 3, 3, 3], nsel=18001, ntab=2),                # more selectors than a 100k slot holds
            _crafted([20] * 4, nsel=2),
        ):
            with pytest.raises(api.Bz2Error) as e:
                c.decompress(stream, out_capacity=1 << 20)
            assert e.value.code == api.E_FORMAT
            assert c.decompress(ok) == good                            # the context is alive and the scratch intact
        c.close()
