"""ctypes wrapper over oracle/liboracle.so (bz2_oracle.c, the C restatement of the reference path).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def build(force: bool = False) -> str:
    so = os.path.join(HERE, "liboracle.so")
    src = os.path.join(HERE, "bz2_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-o", so, src])
    return so


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.orc_compress.restype = C.c_int
        L.orc_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.orc_rle1_blocks.restype = C.c_long
        L.orc_rle1_blocks.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long]
        L.orc_block_stages.restype = C.c_int
        L.orc_block_stages.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_long] + [C.c_void_p] * 8
        L.orc_bwt.restype = C.c_int
        L.orc_bwt.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
        L.orc_compress_ex.restype = C.c_int
        L.orc_compress_ex.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                      C.c_void_p, C.c_void_p, C.c_long, C.POINTER(C.c_long)]
        L.orc_crc32.restype = C.c_uint32
        L.orc_crc32.argtypes = [C.c_void_p, C.c_size_t]
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _src(data: bytes) -> np.ndarray:
    return np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, dtype=np.uint8)


def out_capacity(n: int, bs: int) -> int:
    return n + n // 50 + (n // max(bs - 20, 1) + 2) * 2048 + 4096


def compress(data: bytes, mult: int = 9, base: int = 100000) -> bytes:
    L = lib()
    cap = out_capacity(len(data), base * mult)
    out = np.empty(cap, dtype=np.uint8)
    n = C.c_size_t(0)
    rc = L.orc_compress(_ptr(_src(data)), len(data), base, mult, _ptr(out), cap, C.byref(n))
    if rc != 0:
        raise ValueError("Invalid block size" if rc == -1 else "oracle output overflow")
    return out[:n.value].tobytes()


def compress_ex(data: bytes, mult: int = 9, base: int = 100000):
    """Returns (stream, block_bit_offsets, block_nonprimitive_flags)."""
    L = lib()
    cap = out_capacity(len(data), base * mult)
    out = np.empty(cap, dtype=np.uint8)
    maxb = len(data) // max(base * mult - 20, 1) * 2 + 4
    offs = np.zeros(maxb, dtype=np.uint64)
    flags = np.zeros(maxb, dtype=np.uint8)
    n = C.c_size_t(0)
    nb = C.c_long(0)
    rc = L.orc_compress_ex(_ptr(_src(data)), len(data), base, mult, _ptr(out), cap, C.byref(n), _ptr(offs), _ptr(flags), maxb,
                           C.byref(nb))
    if rc != 0:
        raise ValueError("Invalid block size" if rc == -1 else "oracle output overflow")
    return out[:n.value].tobytes(), offs[:nb.value].copy(), flags[:nb.value].astype(bool)


ORIGPTR_BIT = 48 + 32 + 1  # bits from the block magic to the 24-bit origPtr field


def mask_origptr(stream: bytes, bit_offsets, flags) -> bytes:
    """Zero the 24-bit origPtr field of every block flagged non-primitive (SURVEY F4 comparison rule)."""
    a = np.unpackbits(np.frombuffer(stream, dtype=np.uint8))
    for off, f in zip(bit_offsets, flags):
        if f:
            o = int(off) + ORIGPTR_BIT
            a[o:o + 24] = 0
    return np.packbits(a).tobytes()


def rle1_blocks(data: bytes, mult: int = 9, base: int = 100000):
    L = lib()
    maxb = len(data) // max(base * mult - 20, 1) * 2 + 4
    starts = np.zeros(maxb, dtype=np.uint64)
    lens = np.zeros(maxb, dtype=np.int32)
    crcs = np.zeros(maxb, dtype=np.uint32)
    nb = L.orc_rle1_blocks(_ptr(_src(data)), len(data), base, mult, _ptr(starts), _ptr(lens), _ptr(crcs), maxb)
    assert nb <= maxb
    return starts[:nb].copy(), lens[:nb].copy(), crcs[:nb].copy()


def block_stages(data: bytes, index: int, mult: int = 9, base: int = 100000) -> dict:
    L = lib()
    bs = base * mult
    rle = np.zeros(bs + 1, dtype=np.uint8)
    present = np.zeros(256, dtype=np.uint8)
    Lc = np.zeros(bs + 1, dtype=np.uint8)
    mtf = np.zeros(bs + 2, dtype=np.uint16)
    freq = np.zeros(258, dtype=np.int32)
    sel = np.zeros(bs // 50 + 2, dtype=np.uint8)
    lens = np.zeros(6 * 258, dtype=np.uint8)
    info = np.zeros(9, dtype=np.int64)
    rc = L.orc_block_stages(_ptr(_src(data)), len(data), base, mult, index, _ptr(rle), _ptr(present), _ptr(Lc), _ptr(mtf),
                            _ptr(freq), _ptr(sel), _ptr(lens), _ptr(info))
    if rc != 0:
        raise IndexError("block index out of range")
    n, orig, m, alpha, nsel, crc, nbits, ntab, start = (int(v) for v in info)
    return dict(length=n, orig_ptr=orig, mtf_length=m, alphabet_size=alpha, n_selectors=nsel, crc=crc, block_bits=nbits,
                n_tables=ntab, input_start=start, rle=rle[:n].copy(), present=present.astype(bool), L=Lc[:n].copy(),
                mtf=mtf[:m].copy(), freq=freq, selectors=sel[:nsel].copy(), code_lengths=lens.reshape(6, 258)[:ntab].copy())


def bwt(block: bytes):
    L = lib()
    n = len(block)
    out = np.zeros(max(n, 1), dtype=np.uint8)
    npf = C.c_int(0)
    orig = L.orc_bwt(_ptr(_src(block)), n, _ptr(out), C.byref(npf))
    return out[:n].tobytes(), orig, bool(npf.value)


def crc32(data: bytes) -> int:
    return int(lib().orc_crc32(_ptr(_src(data)), len(data)))
