"""ctypes wrapper over oracle/_ref/libref_<base>.so (the real reference built for the CPU by build_ref.py).

TEST / BASELINE INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's CPU legs.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_libs: dict[int, C.CDLL] = {}


def available(base: int = 100000) -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", f"libref_{base}.so"))


def lib(base: int) -> C.CDLL:
    if base not in _libs:
        path = os.path.join(HERE, "_ref", f"libref_{base}.so")
        if not os.path.exists(path):
            from . import build_ref
            build_ref.build()
        L = C.CDLL(path)
        L.ref_compress.restype = C.c_long
        L.ref_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        L.ref_decompress.restype = C.c_long
        L.ref_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.ref_rle1_blocks.restype = C.c_long
        L.ref_rle1_blocks.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long]
        L.ref_block_stages.restype = C.c_int
        L.ref_block_stages.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_long] + [C.c_void_p] * 7
        L.ref_last_error.restype = C.c_char_p
        L.ref_base_block_size.restype = C.c_int
        assert L.ref_base_block_size() == base
        _libs[base] = L
    return _libs[base]


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def compress(data: bytes, mult: int = 9, base: int = 100000, threads: int = 1) -> bytes:
    L = lib(base)
    src = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, dtype=np.uint8)
    cap = len(data) + len(data) // 50 + (len(data) // (base * mult) + 2) * 2048 + 4096
    out = np.empty(cap, dtype=np.uint8)
    n = L.ref_compress(_ptr(src), len(data), mult, threads, _ptr(out), cap)
    if n < 0:
        raise RuntimeError(L.ref_last_error().decode())
    return out[:n].tobytes()


def decompress(data: bytes, cap: int, base: int = 100000) -> bytes:
    L = lib(base)
    src = np.frombuffer(data, dtype=np.uint8)
    out = np.empty(max(cap, 1), dtype=np.uint8)
    n = L.ref_decompress(_ptr(src), len(data), _ptr(out), cap)
    if n < 0:
        raise RuntimeError("reference decoder: " + (L.ref_last_error().decode() if n == -1 else "output overflow"))
    return out[:n].tobytes()


def rle1_blocks(data: bytes, mult: int = 9, base: int = 100000):
    L = lib(base)
    src = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, dtype=np.uint8)
    maxb = len(data) // (base * mult - 20) * 2 + 4
    starts = np.zeros(maxb, dtype=np.uint64)
    lens = np.zeros(maxb, dtype=np.int32)
    crcs = np.zeros(maxb, dtype=np.uint32)
    nb = L.ref_rle1_blocks(_ptr(src), len(data), mult, _ptr(starts), _ptr(lens), _ptr(crcs), maxb)
    assert nb <= maxb
    return starts[:nb].copy(), lens[:nb].copy(), crcs[:nb].copy()


def block_stages(data: bytes, index: int, mult: int = 9, base: int = 100000) -> dict:
    L = lib(base)
    bs = base * mult
    src = np.frombuffer(data, dtype=np.uint8)
    rle = np.zeros(bs + 1, dtype=np.uint8)
    present = np.zeros(256, dtype=np.uint8)
    Lc = np.zeros(bs + 1, dtype=np.uint8)
    mtf = np.zeros(bs + 2, dtype=np.uint16)
    freq = np.zeros(258, dtype=np.int32)
    sel = np.zeros(bs // 50 + 2, dtype=np.uint8)
    info = np.zeros(8, dtype=np.int64)
    rc = L.ref_block_stages(_ptr(src), len(data), mult, index, _ptr(rle), _ptr(present), _ptr(Lc), _ptr(mtf), _ptr(freq),
                            _ptr(sel), _ptr(info))
    if rc != 0:
        raise RuntimeError(L.ref_last_error().decode())
    n, orig, m, alpha, nsel, crc, nbits = (int(v) for v in info[:7])
    return dict(length=n, orig_ptr=orig, mtf_length=m, alphabet_size=alpha, n_selectors=nsel, crc=crc, huff_bits=nbits,
                rle=rle[:n].copy(), present=present.astype(bool), L=Lc[:n].copy(), mtf=mtf[:m].copy(), freq=freq,
                selectors=sel[:nsel].copy())
