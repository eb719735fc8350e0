"""CPU-only checks of the host side: the RLE1 segment algebra + block walk, CRC combine arithmetic, the C++ facade surface,
the C ABI's symbols and argument errors, and bit-granular stream concatenation (single process and world_size 2 over gloo)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from bzip2_opencl_b200 import api, corpus, multi
from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "tests", "host")


def _compile(src, out, extra=()):
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", os.path.join(HOST, src), "-o", out, *extra]
    subprocess.check_call(cmd)


def test_rle_segment_algebra_and_walk(tmp_path):
    exe = str(tmp_path / "rle_seg")
    _compile("test_rle_seg.cpp", exe)
    out = subprocess.check_output([exe]).decode()
    assert out.startswith("OK")


def test_crc_combine_math(tmp_path):
    exe = str(tmp_path / "crc_math")
    _compile("test_crc_math.cpp", exe)
    assert subprocess.check_output([exe]).decode().strip() == "OK"


def test_packed_round0_key_math(tmp_path):
    # csrc/packed_math.h: keys over the block's alphabet order the rotations by their first h0 bytes and never contradict the full order
    exe = str(tmp_path / "packed_keys")
    _compile("test_packed_keys.cpp", exe)
    assert subprocess.check_output([exe]).decode().strip() == "OK"


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "bz2b200.h")).read()
    declared = set(re.findall(r"\b(bz2b200_[a-z_0-9]+)\s*\(", hdr))
    declared.discard("bz2b200_stage_dump)")
    lib = api.load_library()
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert set(api.EXPORTS) <= declared


def test_header_is_plain_c99_and_links(tmp_path):
    # the boundary is a C ABI: a C99 translation unit (-pedantic -Werror) includes the header, binds the new entry points and links
    exe = str(tmp_path / "c_abi_check")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(HOST, "c_abi_check.c"), "-L" + os.path.join(ROOT, "bzip2_opencl_b200"), "-lbz2b200", "-o", exe])
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "bzip2_opencl_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    assert subprocess.run([exe], env=env).returncode in (0, 1)   # (1: a device exists and pinned memory could be allocated)


def test_c_abi_argument_errors_without_gpu():
    lib = api.load_library()
    h = C.c_void_p()
    assert lib.bz2b200_create(C.byref(h), 0, 100000, 0, 4) == api.E_BLOCKSIZE
    assert lib.bz2b200_create(C.byref(h), 0, 100000, 10, 4) == api.E_BLOCKSIZE
    assert lib.bz2b200_create(C.byref(h), 0, 5, 9, 4) == api.E_BLOCKSIZE
    assert lib.bz2b200_create(C.byref(h), 0, 100000, 9, 0) == api.E_PARALLEL
    assert lib.bz2b200_strerror(api.E_BLOCKSIZE) == b"Invalid block size"
    assert lib.bz2b200_strerror(api.E_PARALLEL) == b"Invalid parallel block count"
    lib.bz2b200_destroy(None)
    if lib.bz2b200_device_count() == 0:
        # the product must fail loudly without a CUDA device -- no CPU fallback
        assert lib.bz2b200_create(C.byref(h), 0, 100000, 9, 4) == api.E_CUDA
        with pytest.raises(api.Bz2Error):
            api.Compressor()
    # the multi-device and slice entry points check their arguments before they touch a device
    ids = (C.c_int * 2)(0, 0)
    assert lib.bz2b200_create_multi(C.byref(h), None, 2, 100000, 9, 4) == api.E_ARG
    assert lib.bz2b200_create_multi(C.byref(h), ids, 0, 100000, 9, 4) == api.E_ARG
    assert lib.bz2b200_create_multi(C.byref(h), ids, 2, 100000, 0, 4) == api.E_BLOCKSIZE
    assert lib.bz2b200_create_multi(C.byref(h), ids, 2, 100000, 9, 0) == api.E_PARALLEL
    assert lib.bz2b200_device_count_of(None) == 0
    assert lib.bz2b200_recommended_blocks(0, 100000, 0) == 0 and lib.bz2b200_recommended_blocks(-1, 100000, 9) == 0
    n = C.c_uint64(0)
    assert lib.bz2b200_scan_blocks_ex(None, None, 0, 1, 0, None, 0, C.byref(n), None) == api.E_ARG
    assert lib.bz2b200_compress_scanned(None, 0, None, 0, None, None, None, 0, None) == api.E_ARG
    assert lib.bz2b200_shift_bits(None, None, 0, 0, 0, None, 0) == api.E_ARG
    lib.bz2b200_host_free(None)
    if lib.bz2b200_device_count() == 0:
        assert lib.bz2b200_create_multi(C.byref(h), ids, 2, 100000, 9, 4) == api.E_CUDA
        assert lib.bz2b200_recommended_blocks(0, 100000, 9) == 0
        assert not lib.bz2b200_host_alloc(4096)   # pinned memory needs the CUDA runtime: no silent malloc fallback
    with pytest.raises(ValueError, match="Invalid block size"):
        api.Compressor(multiplier=0)
    with pytest.raises(ValueError, match="Invalid parallel block count"):
        api.Compressor(parallel=0)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "bzip2_opencl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in text and "import oracle" not in text and "from oracle" not in text, f
                assert "libref_" not in text, f


def test_facade_compiles_and_keeps_the_reference_error_behaviour(tmp_path):
    exe = str(tmp_path / "facade")
    _compile("facade_smoke.cpp", exe, ["-I" + os.path.join(ROOT, "bzip2_opencl_b200", "include"), "-I" + os.path.join(ROOT, "include"),
                                      "-L" + os.path.join(ROOT, "bzip2_opencl_b200"), "-lbz2b200"])
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "bzip2_opencl_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    out = subprocess.check_output([exe, "errors"], env=env).decode()
    assert out.startswith("errors-ok 5")
    assert "f5d3bf25" in out and oracle.crc32(b"abbb") == 0xf5d3bf25


def _split_blocks(stream, offs):
    """cut an oracle stream into per-block (bits, nbits) using the oracle's block bit offsets"""
    bits = np.unpackbits(np.frombuffer(stream, dtype=np.uint8))
    ends = list(offs[1:]) + [len(bits) - ((len(bits) - int(offs[-1])) % 8 if False else 0)]
    return bits, ends


def test_stream_from_blocks_bit_concatenation():
    d = corpus.word_text(350000, 5) + corpus.rand_bytes(120000, 9)
    for mult, base in ((1, 100000), (2, 10000)):
        stream, offs, _ = oracle.compress_ex(d, mult, base)
        _, _, crcs = oracle.rle1_blocks(d, mult, base)
        bits = np.unpackbits(np.frombuffer(stream, dtype=np.uint8))
        nb = len(offs)
        # the trailer starts 80 bits + padding before the end; find it: last block ends where the end magic begins
        magic = np.unpackbits(np.frombuffer((0x177245385090).to_bytes(6, "big"), dtype=np.uint8))
        end = None
        for pad in range(8):
            e = len(bits) - pad - 80
            if e >= 0 and (bits[e:e + 48] == magic).all():
                end = e
        assert end is not None
        bounds = [int(o) for o in offs] + [end]
        for groups in (1, 2, 3, nb):
            parts = []
            per = -(-nb // groups)
            for g in range(groups):
                lo, hi = min(g * per, nb), min((g + 1) * per, nb)
                seg = bits[bounds[lo]:bounds[hi]]
                parts.append((np.packbits(seg).tobytes(), len(seg), crcs[lo:hi]))
            assert api.stream_from_blocks(mult, parts) == stream
    assert api.stream_from_blocks(9, []) == oracle.compress(b"", 9, 100000)


def test_shard_ranges_cover_every_block_once():
    starts = np.array([0, 100, 250, 400, 900], dtype=np.uint64)
    for world in (1, 2, 3, 4, 8):
        r = multi.shard_byte_ranges(starts, 1000, world)
        assert len(r) == world and r[0][0] == 0 and r[-1][1] == 1000
        assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
        assert sum(x[3] - x[2] for x in r) == len(starts)


WORKER = r'''
import os, sys
sys.path.insert(0, {root!r})
import numpy as np
import torch.distributed as dist
from bzip2_opencl_b200 import multi, corpus
from oracle import oracle
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
data = corpus.word_text(300000, 11) + bytes(70000) + corpus.rand_bytes(90000, 12)
mult, base = 1, 100000
# rank 0 scans the boundaries (on a GPU box: bz2b200_scan_blocks) and broadcasts them: metadata only
box = [oracle.rle1_blocks(data, mult, base)[0] if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
starts = box[0]
rng = multi.shard_byte_ranges(starts, len(data), world)[rank]
def compress_blocks(buf, final):  # stand-in for Compressor.compress_blocks: the oracle, block by block (TEST ONLY)
    stream, offs, _ = oracle.compress_ex(buf, mult, base)
    bits = np.unpackbits(np.frombuffer(stream, dtype=np.uint8))
    end = max(e for e in range(len(bits) - 87, len(bits) - 79) if e >= 32 and bits[e:e+24].tolist() == [int(c) for c in format(0x177245, "024b")])
    seg = bits[32:end]
    return dict(bits=np.packbits(seg).tobytes(), nbits=len(seg), crcs=oracle.rle1_blocks(buf, mult, base)[2])
part = multi.compress_shard(compress_blocks, data, rng)
stream = multi.gather_stream(mult, part, rank, world)
if rank == 0:
    assert stream == oracle.compress(data, mult, base), "sharded stream differs"
    print("SHARDED-OK", len(stream))
dist.destroy_process_group()
'''


def test_world_size_2_sharding_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29531", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "SHARDED-OK" in out.stdout


# ---- one stream over several ranks by byte slices (round 2): the host logic of multi.py with the oracle standing in for the device calls
def _oracle_block_bits(buf, mult, base):
    stream, offs, _ = oracle.compress_ex(buf, mult, base)
    bits = np.unpackbits(np.frombuffer(stream, dtype=np.uint8))
    magic = [int(c) for c in format(0x177245385090, "048b")]
    end = max(e for e in range(max(32, len(bits) - 87), len(bits) - 79) if bits[e:e + 48].tolist() == magic)
    return bits[32:end]


def _oracle_scan(buf, final, mult, base):
    """stand-in for bz2b200_scan_blocks_ex (TEST ONLY): blocks of buf; not final = the last one is unfinished"""
    st, _, crcs = oracle.rle1_blocks(buf, mult, base)
    if final or len(st) == 0:
        return len(st), len(buf), crcs
    return len(st) - 1, int(st[-1]), crcs[:len(st) - 1]


def _sliced_stream(data, mult, base, world, ratio=0.9, min_slice=None):
    plan = multi.plan_slices(len(data), world, base * mult, ratio, min_slice)
    nxt, res = {0: 0}, []
    for r in range(world):
        b0, lo, hi = plan[r]
        box = {}

        def scan(first_rel, final):
            nb, used, crcs = _oracle_scan(data[b0 + first_rel:hi], final, mult, base)
            box.update(a=b0 + first_rel, e=b0 + first_rel + used, crcs=crcs)
            return nb, first_rel + used
        s, nb, nx = multi.chain_walk(scan, plan[r], r, world, True, lambda: nxt[r], lambda v: nxt.__setitem__(r + 1, v))
        if nb:
            bits = _oracle_block_bits(data[box["a"]:box["e"]], mult, base)
            res.append((len(bits), np.packbits(bits).tobytes(), box["crcs"]))
        else:
            res.append((0, b"", []))
    nbits_all = [x[0] for x in res]
    tails = [multi.tail_of(x[1][max(0, ((x[0] - 1) >> 3) - 1):((x[0] - 1) >> 3) + 1], x[0]) if x[0] else 0 for x in res]
    placement = multi.place_parts(nbits_all, tails, 32)
    shifted = [multi.shift_bits_host(x[1], x[0], pl[1], pl[2]) for x, pl in zip(res, placement)]
    return multi.assemble_stream(mult, shifted, placement, nbits_all, multi.fold_crcs([x[2] for x in res]))


def test_byte_slices_chain_walk_and_bit_placement():
    d = (corpus.word_text(300000, 11) + bytes(70000) + corpus.rand_bytes(90000, 12) + b"ab" * 30000 + bytes([7]) * 200000 +
         corpus.word_text(100000, 3))
    for mult, base in ((1, 10000), (3, 10000), (1, 100000)):
        exp = oracle.compress(d, mult, base)
        for world in (1, 2, 3, 5, 8):
            assert _sliced_stream(d, mult, base, world, min_slice=2 * base * mult) == exp, (mult, base, world)
    # a run longer than a slice: the chain passes straight through ranks that own no block start
    z = bytes(40000) + b"q" + bytes(9000)
    assert _sliced_stream(z, 1, 1000, 6, min_slice=2000) == oracle.compress(z, 1, 1000)
    plan = multi.plan_slices(10_000_000, 4, 900000, 0.9)
    assert plan[0][1] == 0 and plan[-1][2] == 10_000_000 and all(plan[i][2] == plan[i + 1][1] for i in range(3))
    assert all(b <= lo and lo - b <= multi.head_bytes(900000) for b, lo, hi in plan)
    assert multi.head_bytes(900000) >= 255 * ((900000 - 6) // 5 + 1)


def test_bit_placement_of_parts_against_plain_bit_concatenation():
    # property: parts of arbitrary bit lengths (some empty), packed from bit 0 each, shifted and placed as multi.place_parts says, give the
    # bytes of the plain concatenation of all bits behind the 32 header bits
    rng = np.random.default_rng(7)
    for trial in range(300):
        k = int(rng.integers(1, 9))
        lens = [0 if rng.random() < 0.2 else int(rng.integers(9, 400)) for _ in range(k)]
        bits = [rng.integers(0, 2, n).astype(np.uint8) for n in lens]
        packed = [np.packbits(b).tobytes() for b in bits]
        tails = [multi.tail_of(p[max(0, ((n - 1) >> 3) - 1):((n - 1) >> 3) + 1], n) if n else 0 for p, n in zip(packed, lens)]
        placement = multi.place_parts(lens, tails, 32)
        shifted = [multi.shift_bits_host(p, n, pl[1], pl[2]) for p, n, pl in zip(packed, lens, placement)]
        body = bytearray(4)
        for (off, shift, carry, byte0, n_copy), sh in zip(placement, shifted):
            if n_copy:
                assert byte0 == len(body)
                body += sh[:n_copy]
        allbits = np.concatenate([np.zeros(32, dtype=np.uint8)] + bits) if sum(lens) else np.zeros(32, dtype=np.uint8)
        assert bytes(body) == np.packbits(allbits).tobytes(), (trial, lens)
        # and the trailer continues the last partial byte
        if sum(lens):
            end_bit = 32 + sum(lens)
            last = body[-1] if end_bit & 7 else 0
            tr = multi.trailer_bytes(end_bit, last, 0xdeadbeef)
            tail_bits = np.unpackbits(np.frombuffer(tr, dtype=np.uint8))
            used = end_bit & 7
            want = np.concatenate([allbits[len(allbits) - used:] if used else np.zeros(0, dtype=np.uint8),
                                   np.unpackbits(np.frombuffer((0x177245385090).to_bytes(6, "big") + (0xdeadbeef).to_bytes(4, "big"), dtype=np.uint8))])
            assert (tail_bits[:len(want)] == want).all() and not tail_bits[len(want):].any()


WORKER2 = r'''
import os, sys
sys.path.insert(0, {root!r})
sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np
import torch, torch.distributed as dist
from bzip2_opencl_b200 import multi, corpus
from oracle import oracle
from test_host_logic import _oracle_scan, _oracle_block_bits
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
data = corpus.word_text(300000, 11) + bytes(70000) + corpus.rand_bytes(90000, 12) + bytes(150000) + corpus.word_text(200000, 13)
mult, base = 1, 100000
plan = multi.plan_slices(len(data), world, mult * base, 0.9, min_slice=2 * base)
b0, lo, hi = plan[rank]
box = {{}}
def scan(first_rel, final):
    nb, used, crcs = _oracle_scan(data[b0 + first_rel:hi], final, mult, base)
    box.update(a=b0 + first_rel, e=b0 + first_rel + used, crcs=crcs)
    return nb, first_rel + used
def recv_start():
    t = torch.zeros(1, dtype=torch.int64); dist.recv(t, src=rank - 1); return int(t[0])
def send_start(v):
    dist.send(torch.tensor([v], dtype=torch.int64), dst=rank + 1)
s, nb, nx = multi.chain_walk(scan, plan[rank], rank, world, True, recv_start, send_start)
if nb:
    bits = _oracle_block_bits(data[box["a"]:box["e"]], mult, base)
    part = (len(bits), np.packbits(bits).tobytes(), list(box["crcs"]))
else:
    part = (0, b"", [])
e = (part[0] - 1) >> 3
meta = (part[0], multi.tail_of(part[1][max(0, e - 1):e + 1], part[0]) if part[0] else 0, part[2])
metas = [None] * world
dist.all_gather_object(metas, meta)
placement = multi.place_parts([m[0] for m in metas], [m[1] for m in metas], 32)
pl = placement[rank]
mine = multi.shift_bits_host(part[1], part[0], pl[1], pl[2])[:pl[4]]
parts = [None] * world if rank == 0 else None
dist.gather_object(mine, parts, dst=0)
if rank == 0:
    stream = multi.assemble_stream(mult, parts, placement, [m[0] for m in metas], multi.fold_crcs([m[2] for m in metas]))
    assert stream == oracle.compress(data, mult, base), "sliced stream differs"
    print("SLICED-OK", len(stream))
dist.destroy_process_group()
'''


def test_world_size_2_byte_slices_over_gloo(tmp_path):
    script = tmp_path / "worker2.py"
    script.write_text(WORKER2.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "SLICED-OK" in out.stdout


@pytest.mark.skipif(not os.path.exists("/root/reference/app.cpp"), reason="needs /root/reference")
def test_unmodified_reference_app_compiles_against_the_facade():
    """app.cpp itself, not a re-typed copy -- oracle/build_ref.py feeds it to g++ on stdin from a directory whose include/ is the
    facade (app.cpp:28-29 include "include/OutputStream.hpp" / "include/InputStream.hpp")"""
    from oracle import build_ref
    assert build_ref.build_app()
    for base in (10000, 100000):
        exe = os.path.join(ROOT, "oracle", "_ref", f"app_b200_{base}")
        assert os.path.exists(exe)
        r = subprocess.run([exe, "-h"], capture_output=True)          # argument handling is the reference's own code
        assert r.returncode == 0 and b"--parallel" in r.stdout
        r = subprocess.run([exe, "/nonexistent/file"], capture_output=True)
        assert r.returncode == 1 and b"Failed to open input file." in r.stderr


def test_cli_front_end_argument_handling(tmp_path):
    # the reference's flags and messages (app.cpp:31-92); nothing here reaches the device
    exe = str(tmp_path / "bzip2_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", os.path.join(ROOT, "bzip2_opencl_b200", "cli", "bzip2_b200.cpp"), "-o", exe,
                           "-DBZ2B200_BLOCKSIZE_DEFAULT=100000", "-I" + os.path.join(ROOT, "bzip2_opencl_b200", "include"),
                           "-I" + os.path.join(ROOT, "include"), "-L" + os.path.join(ROOT, "bzip2_opencl_b200"), "-lbz2b200"])
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "bzip2_opencl_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([exe], env=env, capture_output=True)
    assert r.returncode == 1 and b"Usage" in r.stderr
    r = subprocess.run([exe, "-h"], env=env, capture_output=True)
    assert r.returncode == 0 and b"--parallel" in r.stdout and b"--all-streams" in r.stdout
    r = subprocess.run([exe, "file", "--bogus"], env=env, capture_output=True)
    assert r.returncode == 1 and b"Unknown flag!" in r.stderr
    r = subprocess.run([exe, str(tmp_path / "missing")], env=env, capture_output=True)
    assert r.returncode == 1 and b"Failed to open input file." in r.stderr
    f = tmp_path / "plain.txt"
    f.write_bytes(b"x")
    r = subprocess.run([exe, str(f), "-d"], env=env, capture_output=True)
    assert r.returncode == 1 and b".bz2 extenstion" in r.stderr and f.exists()


def _build_stub_facade(tmp_path, sanitizer, base):
    """the OutputStream facade over tests/host/stub_abi.c (the oracle behind the C ABI's names) -- the facade's host pipeline without a GPU"""
    exe = str(tmp_path / f"facade_stub_{sanitizer.split(',')[0]}_{base}")
    obj = []
    for src, cc, std in (("stub_abi.c", "gcc", "-std=c11"), (os.path.join(ROOT, "oracle", "bz2_oracle.c"), "gcc", "-std=c11")):
        o = str(tmp_path / (os.path.basename(src) + f".{sanitizer.split(',')[0]}.o"))
        subprocess.check_call([cc, std, "-O1", "-g", "-fsanitize=" + sanitizer, "-I" + os.path.join(ROOT, "include"), "-c",
                               src if os.path.isabs(src) else os.path.join(HOST, src), "-o", o])
        obj.append(o)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-Wall", "-fsanitize=" + sanitizer, f"-DBZ2B200_BLOCKSIZE_DEFAULT={base}",
                           "-I" + os.path.join(ROOT, "bzip2_opencl_b200", "include"), "-I" + os.path.join(ROOT, "include"),
                           os.path.join(HOST, "facade_stub_driver.cpp"), *obj, "-lpthread", "-o", exe])
    return exe


def _facade_stub_data():
    # several batches at base 2000: text, long runs that cross batch edges (the carried-over bytes), random bytes, a run at the very end
    return (corpus.word_text(30000, 3) + b"\x00" * 70000 + corpus.rand_bytes(20000, 4) + b"ab" * 9000 + b"\xff" * 3000 + corpus.code_text(25000, 5)
            + b"z" * 1021)


@pytest.mark.parametrize("sanitizer", ["thread", "address,undefined"])
def test_facade_host_pipeline_over_a_stub_abi(tmp_path, sanitizer):
    """SURVEY N2 on the CPU: batching, the unfinished block carried into the next batch, the < 8 carried bits, the worker and writer
    threads and readFrom(), with the oracle answering behind the C ABI's names, under ThreadSanitizer and AddressSanitizer/UBSan.
    Every way of feeding the stream gives the oracle's stream, for several batch sizes."""
    base = 2000
    exe = _build_stub_facade(tmp_path, sanitizer, base)
    data = _facade_stub_data()
    src = tmp_path / "in.bin"
    src.write_bytes(data)
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1 exitcode=66", ASAN_OPTIONS="detect_leaks=1", UBSAN_OPTIONS="halt_on_error=1")
    for mult, blocks, mode in ((1, 3, "mixed"), (1, 1, "fd"), (3, 2, "byte"), (9, 4, "vec"), (2, 7, "ptr"), (9, 50, "fd")):
        want = oracle.compress(data, mult, base)
        dst = tmp_path / f"out_{mult}_{blocks}_{mode}.bz2"
        # (readFrom() with several reader threads even on this small file; a trace line per batch must not disturb anything)
        r = subprocess.run([exe, mode, str(src), str(dst), str(mult), "1"], env=dict(env, BZ2B200_FACADE_BLOCKS=str(blocks), STUB_TRACE="1",
                                                                                        BZ2B200_FACADE_READ_MIN="1000", BZ2B200_FACADE_READERS=str(1 + blocks % 5),
                                                                                        BZ2B200_FACADE_TRACE=str(blocks & 1)),
                           capture_output=True)
        assert r.returncode == 0, (mult, blocks, mode, r.stderr.decode()[-2000:])
        assert dst.read_bytes() == want, (mult, blocks, mode)
        calls = int(re.search(rb"stub: (\d+) compress_blocks calls", r.stderr).group(1))
        if blocks * base * mult * 4 < len(data):
            assert calls > 3, (calls, mult, blocks)   # really several batches (1 call is the constructor's empty one)
    # an empty stream and a one-byte stream
    for payload in (b"", b"q"):
        src.write_bytes(payload)
        dst = tmp_path / "small.bz2"
        r = subprocess.run([exe, "byte", str(src), str(dst), "9", "4"], env=env, capture_output=True)
        assert r.returncode == 0, r.stderr.decode()[-2000:]
        assert dst.read_bytes() == oracle.compress(payload, 9, base)


def test_facade_reports_a_failed_batch_and_shuts_down(tmp_path):
    """a batch that fails inside the worker thread surfaces as std::runtime_error in the caller (write or close), and the destructor still
    joins both threads; with the failure in the very first (constructor) call, the constructor throws and leaves nothing behind"""
    exe = _build_stub_facade(tmp_path, "address,undefined", 2000)
    src = tmp_path / "in.bin"
    src.write_bytes(_facade_stub_data())
    for k in (1, 2, 3, 5):
        r = subprocess.run([exe, "fail", str(src), str(tmp_path / "o.bz2"), "1", "1"],
                           env=dict(os.environ, BZ2B200_FACADE_BLOCKS="2", STUB_FAIL_CALL=str(k), ASAN_OPTIONS="detect_leaks=1"), capture_output=True, timeout=120)
        assert r.returncode == 0, (k, r.stdout, r.stderr.decode()[-2000:])
        assert b"injected failure in call %d" % k in r.stdout


def test_facade_read_plan_short_first_batch_then_equal_batches(tmp_path):
    """readFrom() on a regular file: a short first batch (BZ2B200_FACADE_FIRST_MB), the rest of the file in equal batches, several reader
    threads -- the stream does not depend on any of it"""
    base = 2000
    exe = _build_stub_facade(tmp_path, "address,undefined", base)
    data = _facade_stub_data() * 24          # 4.3 MB
    src = tmp_path / "in.bin"
    src.write_bytes(data)
    want = oracle.compress(data, 9, base)
    for first_mb, blocks in ((1, 70), (2, 70), (1, 200)):
        dst = tmp_path / f"o_{first_mb}_{blocks}.bz2"
        r = subprocess.run([exe, "fd", str(src), str(dst), "9", "1"], capture_output=True, timeout=600,
                           env=dict(os.environ, BZ2B200_FACADE_BLOCKS=str(blocks), BZ2B200_FACADE_FIRST_MB=str(first_mb), BZ2B200_FACADE_READ_MIN="65536",
                                    BZ2B200_FACADE_READERS="5", BZ2B200_FACADE_TRACE="1", ASAN_OPTIONS="detect_leaks=1"))
        assert r.returncode == 0, r.stderr.decode()[-2000:]
        assert dst.read_bytes() == want
        sizes = [int(m) for m in re.findall(rb"submit: (\d+) new bytes", r.stderr)]
        assert sizes[0] == min(first_mb << 20, blocks * base * 9), sizes
        rest = sizes[1:-1] if len(sizes) > 2 else []
        assert all(abs(a - b) <= 4096 for a, b in zip(rest, rest[1:])), sizes      # equal batches
        assert sum(sizes) == len(data), sizes
