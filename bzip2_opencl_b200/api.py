"""ctypes binding of the C ABI in include/bz2b200.h (libbz2b200.so) -- the call a Python user makes.

The product is the CUDA library; this module only marshals pointers.  There is NO CPU fallback: if the library is
missing or no CUDA device is usable, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BZ2B200_LIB") or os.path.join(HERE, "libbz2b200.so")  # BZ2B200_LIB: development variants of the same library

OK, E_BLOCKSIZE, E_PARALLEL, E_ARG, E_CUDA, E_NOMEM, E_OUTPUT, E_STATE, E_FORMAT = 0, -1, -2, -3, -4, -5, -6, -7, -8

STAGE_RLE1, STAGE_PRESENT, STAGE_BWT, STAGE_ORIGPTR, STAGE_MTF, STAGE_FREQ, STAGE_SELECTORS, STAGE_LENGTHS, STAGE_RANK = range(1, 10)

EXPORTS = ["bz2b200_device_count", "bz2b200_create", "bz2b200_destroy", "bz2b200_max_output_bytes", "bz2b200_scan_blocks",
           "bz2b200_compress_blocks", "bz2b200_compress_stream", "bz2b200_decompress_stream", "bz2b200_decompress_concat", "bz2b200_get_stats", "bz2b200_stage_dump", "bz2b200_strerror",
           "bz2b200_last_error", "bz2b200_set_profiling", "bz2b200_kernel_times", "bz2b200_create_multi", "bz2b200_device_count_of",
           "bz2b200_recommended_blocks", "bz2b200_host_alloc", "bz2b200_host_free", "bz2b200_scan_prepare", "bz2b200_scan_blocks_ex",
           "bz2b200_compress_scanned", "bz2b200_shift_bits", "bz2b200_compress_blocks_from"]


WALKED_FN = C.CFUNCTYPE(None, C.c_uint64, C.c_void_p)  # bz2b200_walked_fn


class Stats(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("ms_total", "ms_h2d", "ms_d2h", "ms_rle", "ms_bwt_sort0", "ms_bwt_rounds", "ms_bwt_finish",
                                         "ms_mtf", "ms_huffman", "ms_emit", "ms_walk")] + \
               [(n, C.c_uint64) for n in ("n_input", "n_rle", "n_mtf", "n_out_bits", "n_blocks", "bwt_rounds", "bwt_active_volume",
                                          "kernel_launches")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_lib = None


def load_library() -> C.CDLL:
    """Loads libbz2b200.so; raises if it has not been built (python -m bzip2_opencl_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m bzip2_opencl_b200.build` (nvcc, sm_100a). "
                           "There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, sz, u64p = C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)
    L.bz2b200_device_count.restype = C.c_int
    L.bz2b200_create.restype = C.c_int
    L.bz2b200_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int]
    L.bz2b200_destroy.restype = None
    L.bz2b200_destroy.argtypes = [vp]
    L.bz2b200_max_output_bytes.restype = sz
    L.bz2b200_max_output_bytes.argtypes = [vp, sz]
    L.bz2b200_scan_blocks.restype = C.c_int
    L.bz2b200_scan_blocks.argtypes = [vp, vp, sz, vp, sz, u64p]
    L.bz2b200_compress_blocks.restype = C.c_int
    L.bz2b200_compress_blocks.argtypes = [vp, vp, sz, C.c_int, C.c_uint, vp, sz, u64p, C.POINTER(sz), vp, vp, sz, u64p]
    L.bz2b200_compress_stream.restype = C.c_int
    L.bz2b200_compress_stream.argtypes = [vp, vp, sz, vp, sz, C.POINTER(sz)]
    L.bz2b200_decompress_stream.restype = C.c_int
    L.bz2b200_decompress_stream.argtypes = [vp, vp, sz, vp, sz, C.POINTER(sz)]
    L.bz2b200_decompress_concat.restype = C.c_int
    L.bz2b200_decompress_concat.argtypes = [vp, vp, sz, vp, sz, C.POINTER(sz), C.POINTER(C.c_uint64)]
    L.bz2b200_get_stats.restype = C.c_int
    L.bz2b200_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.bz2b200_stage_dump.restype = C.c_int
    L.bz2b200_stage_dump.argtypes = [vp, C.c_uint64, C.c_int, vp, sz, C.POINTER(sz)]
    L.bz2b200_strerror.restype = C.c_char_p
    L.bz2b200_strerror.argtypes = [C.c_int]
    L.bz2b200_last_error.restype = C.c_char_p
    L.bz2b200_last_error.argtypes = [vp]
    L.bz2b200_set_profiling.restype = C.c_int
    L.bz2b200_set_profiling.argtypes = [vp, C.c_int]
    L.bz2b200_kernel_times.restype = C.c_int
    L.bz2b200_kernel_times.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double), u64p, u64p]
    L.bz2b200_create_multi.restype = C.c_int
    L.bz2b200_create_multi.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.c_int]
    L.bz2b200_device_count_of.restype = C.c_int
    L.bz2b200_device_count_of.argtypes = [vp]
    L.bz2b200_recommended_blocks.restype = C.c_int
    L.bz2b200_recommended_blocks.argtypes = [C.c_int, C.c_int, C.c_int]
    L.bz2b200_host_alloc.restype = vp
    L.bz2b200_host_alloc.argtypes = [sz]
    L.bz2b200_host_free.restype = None
    L.bz2b200_host_free.argtypes = [vp]
    L.bz2b200_scan_prepare.restype = C.c_int
    L.bz2b200_scan_prepare.argtypes = [vp, vp, sz]
    L.bz2b200_scan_blocks_ex.restype = C.c_int
    L.bz2b200_scan_blocks_ex.argtypes = [vp, vp, sz, C.c_int, C.c_uint64, vp, sz, u64p, C.POINTER(sz)]
    L.bz2b200_compress_scanned.restype = C.c_int
    L.bz2b200_compress_scanned.argtypes = [vp, C.c_uint, vp, sz, u64p, vp, vp, sz, u64p]
    L.bz2b200_compress_blocks_from.restype = C.c_int
    L.bz2b200_compress_blocks_from.argtypes = [vp, vp, sz, C.c_int, C.c_uint64, C.c_uint, vp, sz, u64p, C.POINTER(sz), vp, vp, sz, u64p, WALKED_FN, vp]
    L.bz2b200_shift_bits.restype = C.c_int
    L.bz2b200_shift_bits.argtypes = [vp, vp, C.c_uint64, C.c_uint, C.c_uint, vp, sz]
    _lib = L
    return L


class Bz2Error(RuntimeError):
    def __init__(self, code: int, text: str):
        super().__init__(f"bz2b200 error {code}: {text}")
        self.code = code


def _np_ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Compressor:
    """One handle = one GPU.  block size = base * multiplier; `parallel` = blocks per device batch (the reference's -p)."""

    def __init__(self, multiplier: int = 9, parallel: int = 64, base: int = 100000, device: int = 0, devices=None):
        """devices: a list of CUDA device ids -> one multi-device handle (bz2b200_create_multi); otherwise one handle on `device`."""
        self.lib = load_library()
        self.h = C.c_void_p()
        self.base, self.multiplier, self.parallel, self.device = base, multiplier, parallel, device
        self.devices = list(devices) if devices is not None else None
        if self.devices is not None:
            ids = (C.c_int * len(self.devices))(*self.devices)
            rc = self.lib.bz2b200_create_multi(C.byref(self.h), ids, len(self.devices), base, multiplier, parallel)
            if rc == E_ARG:
                raise ValueError("Invalid device list")
        else:
            rc = self.lib.bz2b200_create(C.byref(self.h), device, base, multiplier, parallel)
        if rc == E_BLOCKSIZE:
            raise ValueError("Invalid block size")  # std::invalid_argument, OutputStream.hpp:73-76
        if rc == E_PARALLEL:
            raise ValueError("Invalid parallel block count")  # OutputStream.hpp:78-81
        if rc != OK:
            raise Bz2Error(rc, self.lib.bz2b200_strerror(rc).decode() + " (no CUDA device? there is no CPU fallback)")

    def close(self):
        if self.h:
            self.lib.bz2b200_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != OK:
            raise Bz2Error(rc, self.lib.bz2b200_strerror(rc).decode() + ": " + self.lib.bz2b200_last_error(self.h).decode())

    @property
    def block_size(self) -> int:
        return self.base * self.multiplier

    def max_output_bytes(self, n: int) -> int:
        return int(self.lib.bz2b200_max_output_bytes(self.h, n))

    def max_blocks(self, n: int) -> int:
        return n // max(1, (self.block_size - 5) * 4 // 5) + 4

    # ---- host-buffer API -------------------------------------------------------------------------------------
    def compress(self, data) -> bytes:
        """Whole .bz2 stream for `data` (bytes-like, host)."""
        src = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, dtype=np.uint8)
        cap = self.max_output_bytes(len(data)) + 64
        out = np.empty(cap, dtype=np.uint8)
        n = C.c_size_t(0)
        self._check(self.lib.bz2b200_compress_stream(self.h, _np_ptr(src), len(data), _np_ptr(out), cap, C.byref(n)))
        return out[:n.value].tobytes()

    # ---- decompress side ---------------------------------------------------------------------------
    def decompress(self, stream, out_capacity: int | None = None, multi: bool = False) -> bytes:
        """Whole .bz2 stream -> original bytes (host buffers).  Raises Bz2Error(E_FORMAT, <the reference's exception text>) on a
        bad stream.  Without out_capacity the buffer grows until the output fits.  multi=True decodes every concatenated stream
        (bz2b200_decompress_concat); self.last_streams is then the number of streams decoded."""
        src = np.frombuffer(stream, dtype=np.uint8) if len(stream) else np.zeros(1, dtype=np.uint8)
        cap = out_capacity if out_capacity is not None else max(1 << 16, 6 * len(stream))
        while True:
            out = np.empty(max(cap, 1), dtype=np.uint8)
            n = C.c_size_t(0)
            if multi:
                ns = C.c_uint64(0)
                rc = self.lib.bz2b200_decompress_concat(self.h, _np_ptr(src), len(stream), _np_ptr(out), cap, C.byref(n), C.byref(ns))
                self.last_streams = int(ns.value)
            else:
                rc = self.lib.bz2b200_decompress_stream(self.h, _np_ptr(src), len(stream), _np_ptr(out), cap, C.byref(n))
            if rc == E_OUTPUT and out_capacity is None:
                cap = int(n.value) if int(n.value) > cap else 2 * cap   # n = what the stream decodes to: one retry
                continue
            self._check(rc)
            return out[:n.value].tobytes()

    def decompress_ptr(self, in_ptr: int, nbytes: int, out_ptr: int, out_cap: int) -> int:
        """Raw-pointer variant (host or device pointers); returns the number of bytes written."""
        n = C.c_size_t(0)
        self._check(self.lib.bz2b200_decompress_stream(self.h, C.c_void_p(in_ptr), nbytes, C.c_void_p(out_ptr), out_cap, C.byref(n)))
        return int(n.value)

    def scan_blocks(self, data=None, *, ptr: int = 0, nbytes: int = 0) -> np.ndarray:
        if data is not None:
            src = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, dtype=np.uint8)
            ptr, nbytes = src.ctypes.data, len(data)
        cap = self.max_blocks(nbytes)
        starts = np.zeros(cap, dtype=np.uint64)
        nb = C.c_uint64(0)
        self._check(self.lib.bz2b200_scan_blocks(self.h, C.c_void_p(ptr), nbytes, _np_ptr(starts), cap, C.byref(nb)))
        return starts[:nb.value].copy()

    # ---- slice primitives (one stream over several handles) ---------------------------------------------
    def scan_prepare(self, ptr: int, nbytes: int):
        self._check(self.lib.bz2b200_scan_prepare(self.h, C.c_void_p(ptr), nbytes))

    def scan_blocks_ex(self, ptr: int, nbytes: int, is_final: bool, first_start: int = 0, want_starts: bool = True):
        """Returns (block starts relative to ptr, consumed): consumed = where the first unfinished block starts (the next slice's first_start)."""
        cap = self.max_blocks(nbytes) if want_starts else 0
        starts = np.zeros(max(cap, 1), dtype=np.uint64)
        nb = C.c_uint64(0)
        used = C.c_size_t(0)
        self._check(self.lib.bz2b200_scan_blocks_ex(self.h, C.c_void_p(ptr), nbytes, int(is_final), first_start, _np_ptr(starts) if want_starts else None, cap,
                                                    C.byref(nb), C.byref(used)))
        return starts[:nb.value].copy() if want_starts else int(nb.value), int(used.value)

    def compress_scanned_ptr(self, out_ptr: int, out_cap: int, n_blocks: int, start_bit: int = 0):
        cap = max(1, n_blocks)
        crcs = np.zeros(cap, dtype=np.uint32)
        nbits_b = np.zeros(cap, dtype=np.uint64)
        nbits = C.c_uint64(0)
        nb = C.c_uint64(0)
        self._check(self.lib.bz2b200_compress_scanned(self.h, start_bit, C.c_void_p(out_ptr), out_cap, C.byref(nbits), _np_ptr(crcs), _np_ptr(nbits_b), cap,
                                                      C.byref(nb)))
        return dict(nbits=nbits.value, crcs=crcs[:nb.value].copy(), block_nbits=nbits_b[:nb.value].copy())

    def compress_blocks_from_ptr(self, in_ptr: int, nbytes: int, out_ptr: int, out_cap: int, is_final: bool, first_start: int, on_walked=None, start_bit: int = 0):
        """Scan + compress of a slice (bz2b200_compress_blocks_from): on_walked(next_start) is called from inside the call as soon as the walk
        has been through the slice.  Returns dict(nbits, consumed, crcs, block_nbits)."""
        cap = self.max_blocks(nbytes)
        crcs = np.zeros(cap, dtype=np.uint32)
        nbits_b = np.zeros(cap, dtype=np.uint64)
        nbits = C.c_uint64(0)
        consumed = C.c_size_t(0)
        nb = C.c_uint64(0)
        cb = WALKED_FN(lambda nxt, _u: on_walked(int(nxt))) if on_walked is not None else WALKED_FN(0)
        self._check(self.lib.bz2b200_compress_blocks_from(self.h, C.c_void_p(in_ptr), nbytes, int(is_final), first_start, start_bit, C.c_void_p(out_ptr), out_cap,
                                                          C.byref(nbits), C.byref(consumed), _np_ptr(crcs), _np_ptr(nbits_b), cap, C.byref(nb), cb, None))
        return dict(nbits=nbits.value, consumed=consumed.value, crcs=crcs[:nb.value].copy(), block_nbits=nbits_b[:nb.value].copy())

    def shift_bits_ptr(self, src_ptr: int, nbits: int, shift: int, carry: int, dst_ptr: int, dst_cap: int):
        self._check(self.lib.bz2b200_shift_bits(self.h, C.c_void_p(src_ptr), nbits, shift, carry, C.c_void_p(dst_ptr), dst_cap))

    # ---- raw-pointer API (host or device pointers) -----------------------------------------------------------
    def compress_blocks_ptr(self, in_ptr: int, nbytes: int, out_ptr: int, out_cap: int, is_final: bool = True, start_bit: int = 0):
        """Returns dict(nbits, consumed, crcs, block_nbits). in/out may be device pointers (e.g. torch tensor.data_ptr())."""
        cap = self.max_blocks(nbytes)
        crcs = np.zeros(cap, dtype=np.uint32)
        nbits_b = np.zeros(cap, dtype=np.uint64)
        nbits = C.c_uint64(0)
        consumed = C.c_size_t(0)
        nb = C.c_uint64(0)
        self._check(self.lib.bz2b200_compress_blocks(self.h, C.c_void_p(in_ptr), nbytes, int(is_final), start_bit, C.c_void_p(out_ptr), out_cap,
                                                     C.byref(nbits), C.byref(consumed), _np_ptr(crcs), _np_ptr(nbits_b), cap, C.byref(nb)))
        return dict(nbits=nbits.value, consumed=consumed.value, crcs=crcs[:nb.value].copy(), block_nbits=nbits_b[:nb.value].copy())

    def compress_blocks(self, data, is_final: bool = True, start_bit: int = 0):
        src = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, dtype=np.uint8)
        cap = self.max_output_bytes(len(data)) + 64
        out = np.zeros(cap, dtype=np.uint8)
        r = self.compress_blocks_ptr(src.ctypes.data, len(data), out.ctypes.data, cap, is_final, start_bit)
        r["bits"] = out[:(start_bit + r["nbits"] + 7) // 8].tobytes()
        return r

    def stats(self) -> dict:
        s = Stats()
        self._check(self.lib.bz2b200_get_stats(self.h, C.byref(s)))
        return s.as_dict()

    def set_profiling(self, on: bool = True):
        self._check(self.lib.bz2b200_set_profiling(self.h, int(on)))

    def kernel_times(self) -> dict:
        """{kernel name: (total ms, launches, algorithmic bytes)} since profiling was switched on."""
        out = {}
        i = 0
        while True:
            name, ms, cnt, nbytes = C.c_char_p(), C.c_double(0), C.c_uint64(0), C.c_uint64(0)
            if self.lib.bz2b200_kernel_times(self.h, i, C.byref(name), C.byref(ms), C.byref(cnt), C.byref(nbytes)) != OK:
                break
            out[name.value.decode()] = (ms.value, cnt.value, nbytes.value)
            i += 1
        return out

    def stage(self, block_index: int, stage: int) -> np.ndarray:
        dtype = {STAGE_ORIGPTR: np.uint32, STAGE_MTF: np.uint16, STAGE_FREQ: np.int32, STAGE_RANK: np.uint32}.get(stage, np.uint8)
        cap = (self.block_size + 64) * 4
        buf = np.zeros(cap, dtype=np.uint8)
        n = C.c_size_t(0)
        self._check(self.lib.bz2b200_stage_dump(self.h, block_index, stage, _np_ptr(buf), cap, C.byref(n)))
        return buf[:n.value].view(dtype).copy()


def stream_from_blocks(multiplier: int, parts) -> bytes:
    """Host-side concatenation used by multi-GPU callers: parts = [(bits: bytes, nbits: int, crcs: array), ...] in block order.
    Adds the stream header, folds the block CRCs (OutputStream.hpp:202) and appends the trailer (OutputStream.hpp:169-172)."""
    acc = 0
    nacc = 0
    chunks = [b"BZh" + bytes([ord("0") + multiplier])]
    crc = 0
    for bits, nbits, crcs in parts:
        for c in crcs:
            crc = (((crc << 1) | (crc >> 31)) ^ int(c)) & 0xffffffff
        if nbits == 0:
            continue
        a = np.frombuffer(bits, dtype=np.uint8)[:(nbits + 7) // 8]
        if nacc == 0:
            full = nbits // 8
            chunks.append(a[:full].tobytes())
            rem = nbits - 8 * full
            if rem:
                acc, nacc = int(a[full]) >> (8 - rem), rem
        else:
            # shift the whole part right by nacc bits
            wide = np.concatenate([a, np.zeros(1, dtype=np.uint8)]).astype(np.uint16)
            shifted = ((wide >> nacc) & 0xff).astype(np.uint8)
            shifted[1:] |= ((wide[:-1] << (8 - nacc)) & 0xff).astype(np.uint8)
            shifted[0] |= (acc << (8 - nacc)) & 0xff
            total = nacc + nbits
            full = total // 8
            chunks.append(shifted[:full].tobytes())
            rem = total - 8 * full
            acc, nacc = ((int(shifted[full]) >> (8 - rem)) if rem else 0), rem
    tail = (0x177245 << 56) | (0x385090 << 32) | crc
    v = (acc << 80) | tail
    total = nacc + 80
    pad = (-total) % 8
    v <<= pad
    chunks.append(int(v).to_bytes((total + pad) // 8, "big"))
    return b"".join(chunks)
