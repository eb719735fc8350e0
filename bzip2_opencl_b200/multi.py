"""Multi-GPU host logic, one PROCESS per GPU: blocks share no state (Thesis p.30), so one stream is cut into
contiguous byte slices, one per rank, and every rank compresses the blocks that START in its slice.  There is no data-path
collective; what moves between ranks is

  * a halo: the `head_bytes` bytes in front of a slice (context for a block that began in the previous slice), once, up front;
  * one number along a chain: the stream offset where the previous rank's last complete block ended (the sequential machine of
    BlockCompressor::write, BlockCompressor.hpp:69-96, restarts there);
  * metadata for the assembly: every rank's bit count, its trailing bits and its block CRCs;
  * each rank's packed bits, shifted on its own GPU to their bit offset, to wherever the stream is assembled.

The same scheme runs inside one process behind bz2b200_create_multi (csrc/bz2b200.cu, multi_compress_blocks); this module is the
process-per-GPU form on top of the slice primitives of the C ABI (bz2b200_scan_prepare / bz2b200_scan_blocks_ex /
bz2b200_compress_scanned / bz2b200_shift_bits).  The older block-range helpers (one rank scans, everyone compresses a block range,
rank 0 concatenates on the host) are kept below: tests use them to cross-check the two assemblies.
"""
from __future__ import annotations

import numpy as np


# ---------------------------------------------------------------------------------------------------------------- slices
def head_bytes(block_size: int) -> int:
    """Raw bytes an unfinished block can hold at the very most: limit/5 + 1 full 255-byte run chunks, plus slack (same rule as
    multi_head_bytes in csrc/bz2b200.cu)."""
    limit = block_size - 6
    return ((limit // 5 + 2) * 255 + 2 * block_size + 4095) // 4096 * 4096


def plan_slices(total: int, world: int, block_size: int, ratio: float = 0.95, min_slice: int | None = None):
    """[(base, lo, hi)] per rank: rank r owns the blocks that start in [lo, hi) and stages bytes [base, hi) (base = lo - head).
    Slices shrink by `ratio` along the chain (rank r starts about r walks later than rank 0); ranks that would get less than
    `min_slice` bytes get nothing (lo == hi == total)."""
    if min_slice is None:
        min_slice = 16 * block_size
    k = max(1, min(world, total // max(1, min_slice)))
    w = [ratio ** d for d in range(k)]
    tot = sum(w)
    head = head_bytes(block_size)
    out, lo, acc = [], 0, 0.0
    for d in range(world):
        if d < k:
            acc += w[d]
            hi = total if d + 1 == k else min(total, int(total * (acc / tot)) // 4096 * 4096)
            hi = max(hi, lo)
        else:
            hi = total
        base = (lo - head) // 4096 * 4096 if lo > head else 0
        out.append((base, lo, hi))
        lo = hi
    return out


def chain_walk(scan, entry, rank: int, world: int, is_final: bool, recv_start, send_start):
    """The walk of one rank.  scan(first_start_rel, final) -> (n_blocks, consumed_rel) runs bz2b200_scan_blocks_ex on the staged
    bytes [base, hi); recv_start() -> absolute offset where this rank's first block starts (rank 0: 0); send_start(v) hands the
    next rank its start.  Returns (first, n_blocks, next)."""
    base, lo, hi = entry
    s = 0 if rank == 0 else recv_start()
    n_blocks, nxt = 0, s
    if s < hi:
        if s < base:
            raise RuntimeError("an unfinished block is longer than the staged head")
        n_blocks, used = scan(s - base, bool(is_final and rank == world - 1))
        nxt = base + used
    if rank + 1 < world:
        send_start(nxt)
    return s, n_blocks, nxt


def tail_of(last_two: bytes, nbits: int) -> int:
    """16-bit window over the part's last two bytes (bytes e-1, e with e = (nbits-1)//8; one byte if the part has only one)."""
    if nbits == 0:
        return 0
    if len(last_two) == 1:
        return last_two[0]
    return (last_two[0] << 8) | last_two[1]


def place_parts(nbits_all, tails_all, start_bit: int = 0):
    """Where every rank's bits go.  Returns [(bit offset, shift, carry byte, first stream byte, bytes to copy)] per rank:
    the part is shifted right by `shift` bits with `carry` (the trailing bits of the part in front) in the top of its first byte,
    and copies `n_copy` bytes to stream byte `byte0`; the byte a later part starts in belongs to that later part.  A non-empty part
    holds at least one block, i.e. more than the 7 bits a carry can be long (a block header alone is 105 bits)."""
    assert all(int(nb) == 0 or int(nb) >= 8 for nb in nbits_all)
    k = len(nbits_all)
    off, pos = [], start_bit
    for nb in nbits_all:
        off.append(pos)
        pos += int(nb)
    last = max([d for d in range(k) if nbits_all[d]], default=-1)
    out = []
    for d in range(k):
        nb = int(nbits_all[d])
        if nb == 0:
            out.append((off[d], 0, 0, off[d] >> 3, 0))
            continue
        shift = off[d] & 7
        carry = 0
        if shift:
            p = d - 1
            while p >= 0 and not nbits_all[p]:
                p -= 1
            if p >= 0:
                q = int(nbits_all[p])
                e = (q - 1) >> 3
                have = q - 8 * (e - 1) if e else q + 8
                carry = ((int(tails_all[p]) >> (16 - have)) & ((1 << shift) - 1)) << (8 - shift)
        n_dst = (shift + nb + 7) // 8
        shares_last = d != last and ((off[d] + nb) & 7) != 0
        out.append((off[d], shift, carry, off[d] >> 3, n_dst - (1 if shares_last else 0)))
    return out


def fold_crcs(crc_lists) -> int:
    """crc = rotl(crc, 1) ^ blockCRC over all blocks in stream order (OutputStream.hpp:202)."""
    crc = 0
    for crcs in crc_lists:
        for c in crcs:
            crc = (((crc << 1) | (crc >> 31)) ^ int(c)) & 0xffffffff
    return crc


def trailer_bytes(end_bit: int, last_byte: int, crc: int) -> bytes:
    """End-of-stream magic, combined CRC and padding (OutputStream.hpp:169-172) from bit `end_bit` on; `last_byte` is the stream
    byte that bit lies in (its low bits zero).  Returns the bytes from stream byte end_bit // 8 on."""
    used = end_bit & 7
    v = ((last_byte >> (8 - used)) << 80 if used else 0) | (0x177245 << 56) | (0x385090 << 32) | (crc & 0xffffffff)
    total = used + 80
    pad = (-total) % 8
    return int(v << pad).to_bytes((total + pad) // 8, "big")


def shift_bits_host(bits: bytes, nbits: int, shift: int, carry: int) -> bytes:
    """Host restatement of k_shift_bits (tests / CPU stand-ins)."""
    n_dst = (shift + nbits + 7) // 8
    if n_dst == 0:
        return b""
    a = np.frombuffer(bits, dtype=np.uint8)[:(nbits + 7) // 8]
    wide = np.zeros(n_dst + 1, dtype=np.uint16)
    wide[1:1 + len(a)] = a
    out = ((((wide[:-1].astype(np.uint32) << 8) | wide[1:]) >> shift) & 0xff).astype(np.uint8)
    out[0] = (out[0] & (0xff >> shift)) | carry
    return out.tobytes()


def assemble_stream(multiplier: int, shifted_parts, placement, nbits_all, crc: int) -> bytes:
    """The whole .bz2 stream on the host from the ranks' shifted parts (rank order; placement = place_parts(nbits_all, tails, 32):
    the body starts behind the 4-byte header).  Every stream byte comes from exactly one part; the trailer goes behind the last bit."""
    body = bytearray(b"BZh" + bytes([ord("0") + multiplier]))
    for (off, shift, carry, byte0, n_copy), p in zip(placement, shifted_parts):
        if n_copy:
            assert byte0 == len(body), "parts are not contiguous"
            body += bytes(p[:n_copy])
    end_bit = 32 + sum(int(x) for x in nbits_all)
    last = body.pop() if end_bit & 7 else 0
    assert len(body) == end_bit >> 3
    return bytes(body) + trailer_bytes(end_bit, last, crc)


# ---------------------------------------------------------------------------------------------------------------- block ranges (round 1)
def shard_block_ranges(n_blocks: int, world: int):
    """[(b_lo, b_hi)] per rank: contiguous ranges of ceil(B / world) blocks (keeps each GPU's input slice contiguous)."""
    per = -(-n_blocks // world) if n_blocks else 0
    return [(min(r * per, n_blocks), min((r + 1) * per, n_blocks)) for r in range(world)]


def shard_byte_ranges(starts, n_bytes: int, world: int):
    """[(byte_lo, byte_hi, b_lo, b_hi)] per rank from the block start offsets of the whole stream."""
    nb = len(starts)
    out = []
    for b_lo, b_hi in shard_block_ranges(nb, world):
        lo = int(starts[b_lo]) if b_lo < nb else n_bytes
        hi = int(starts[b_hi]) if b_hi < nb else n_bytes
        out.append((lo, hi, b_lo, b_hi))
    return out


def compress_shard(compress_blocks, data, rng):
    """Runs `compress_blocks(bytes, is_final=True)` (the C-ABI call) on one rank's byte range.  A range that ends on a block
    boundary compresses to exactly the blocks the whole stream would contain there: the machine starts fresh at every
    block start and finishRLE() flushes the same pending run the refusal would have (BlockCompressor.hpp:71-74,112-118)."""
    lo, hi, b_lo, b_hi = rng
    if hi <= lo:
        return b"", 0, np.zeros(0, dtype=np.uint32)
    r = compress_blocks(data[lo:hi], True)
    assert len(r["crcs"]) == b_hi - b_lo, "shard does not reproduce the stream's block boundaries"
    return r["bits"], int(r["nbits"]), r["crcs"]


def gather_stream(multiplier: int, part, rank: int, world: int, group=None):
    """gather every rank's (bits, nbits, crcs) on rank 0 and build the .bz2 stream there; other ranks get None."""
    import torch.distributed as dist
    from .api import stream_from_blocks
    parts = [None] * world if rank == 0 else None
    dist.gather_object(part, parts, dst=0, group=group)
    if rank != 0:
        return None
    return stream_from_blocks(multiplier, parts)


# ---------------------------------------------------------------------------------------------------------------- one rank of a stream
class StreamRank:
    """One process (= one GPU) of `world` compressing ONE stream of `total` bytes by byte slices.  torch.distributed must be
    initialised; `meta_group` is a gloo group for the few numbers that travel between ranks (the chain of block starts, bit
    counts, CRCs), the packed bits travel over the default (NCCL) group.  Used by bench.py --gpus N and the tests."""

    def __init__(self, comp, total: int, rank: int, world: int, meta_group=None, ratio: float = 0.95, is_final: bool = True):
        import torch
        self.torch = torch
        self.comp, self.total, self.rank, self.world, self.meta, self.is_final = comp, total, rank, world, meta_group, is_final
        self.plan = plan_slices(total, world, comp.block_size, ratio)
        self.base, self.lo, self.hi = self.plan[rank]
        self.halo = self.lo - self.base
        self.cap = comp.max_output_bytes(self.hi - self.base) + 64
        self.out = torch.zeros(self.cap, dtype=torch.uint8, device="cuda")
        self.shifted = torch.zeros(self.cap + 16, dtype=torch.uint8, device="cuda")
        self.max_blocks = comp.max_blocks(total)

    # -- device-resident input: `buf` holds [base, hi) with the slice itself in buf[halo:]; the halo comes from the rank(s) in front
    def exchange_halo(self, buf):
        import torch.distributed as dist
        ops = []
        if self.rank + 1 < self.world:
            nb, nlo, nhi = self.plan[self.rank + 1]
            need = nlo - nb
            if need:
                assert need <= self.hi - self.base, "halo longer than the slice in front"
                ops.append(dist.P2POp(dist.isend, buf[buf.numel() - need:], self.rank + 1))
        if self.rank > 0 and self.halo:
            ops.append(dist.P2POp(dist.irecv, buf[:self.halo], self.rank - 1))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def _recv_start(self):
        import torch.distributed as dist
        t = self.torch.zeros(1, dtype=self.torch.int64)
        dist.recv(t, src=self.rank - 1, group=self.meta)
        return int(t[0])

    def _send_start(self, v):
        import torch.distributed as dist
        dist.send(self.torch.tensor([v], dtype=self.torch.int64), dst=self.rank + 1, group=self.meta)

    def compress(self, ptr: int):
        """ptr -> bytes [base, hi) of the stream (host or device memory).  Leaves this rank's bits, shifted to their bit offset in
        the stream (32 header bits in front), in self.shifted; returns dict(placement of every rank, nbits of every rank, crc,
        n_copy, byte0, end_bit)."""
        import time
        import torch.distributed as dist
        torch = self.torch
        c = self.comp
        n = self.hi - self.base
        ph = {}
        t0 = time.perf_counter()
        if n and self.rank > 0:
            c.scan_prepare(ptr, n)   # (rank 0 starts at once: its copies run under its kernels)
        ph["prepare_queue"] = time.perf_counter() - t0
        s = 0
        if self.rank > 0:
            t = time.perf_counter()
            s = self._recv_start()
            ph["wait_for_start"] = time.perf_counter() - t
        ph["start_known_at"] = time.perf_counter() - t0
        nbits, crcs, tail, nb = 0, np.zeros(0, dtype=np.uint32), 0, 0
        handed = [False]

        def on_walked(nxt):   # called from inside the compress call as soon as the walk is through the slice
            handed[0] = True
            ph["walked_at"] = time.perf_counter() - t0
            if self.rank + 1 < self.world:
                self._send_start(self.base + nxt)
        t = time.perf_counter()
        if s < self.hi:
            if s < self.base:
                raise RuntimeError("an unfinished block is longer than the staged head")
            q = c.compress_blocks_from_ptr(ptr, n, self.out.data_ptr(), self.cap, bool(self.is_final and self.rank == self.world - 1), s - self.base, on_walked)
            nbits, crcs, nb = int(q["nbits"]), q["crcs"], len(q["crcs"])
            if nbits:
                e = (nbits - 1) >> 3
                tail = tail_of(bytes(self.out[max(0, e - 1):e + 1].cpu().numpy()), nbits)
        if not handed[0] and self.rank + 1 < self.world:
            self._send_start(s)       # nothing of this slice starts a block: the chain passes through
        ph["compress"] = time.perf_counter() - t
        t = time.perf_counter()
        # metadata: bit count, trailing bits, block CRCs of every rank
        mine = torch.zeros(3 + self.max_blocks, dtype=torch.int64)
        mine[0], mine[1], mine[2] = nbits, tail, len(crcs)
        if len(crcs):
            mine[3:3 + len(crcs)] = torch.from_numpy(crcs.astype(np.int64))
        allm = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(allm, mine, group=self.meta)
        nbits_all = [int(m[0]) for m in allm]
        placement = place_parts(nbits_all, [int(m[1]) for m in allm], 32)
        off, shift, carry, byte0, n_copy = placement[self.rank]
        if nbits:
            c.shift_bits_ptr(self.out.data_ptr(), nbits, shift, carry, self.shifted.data_ptr(), self.cap + 16)
        crc = fold_crcs([m[3:3 + int(m[2])].tolist() for m in allm])
        ph["meta_and_shift"] = time.perf_counter() - t
        ph["done_at"] = time.perf_counter() - t0
        self.phases = {k: round(1e3 * v, 3) for k, v in ph.items()}  # host clock, ms, this rank
        return dict(placement=placement, nbits_all=nbits_all, crc=crc, n_copy=n_copy, byte0=byte0, end_bit=32 + sum(nbits_all), n_blocks=nb,
                    blocks_all=[int(m[2]) for m in allm])

    def gather_device(self, res, dst):
        """Every rank's shifted bytes to rank 0's device buffer `dst` (the whole stream: header, parts, trailer).  Returns the
        stream length in bytes on rank 0."""
        import torch.distributed as dist
        torch = self.torch
        ops = []
        if self.rank == 0:
            if res["n_copy"]:
                dst[res["byte0"]:res["byte0"] + res["n_copy"]].copy_(self.shifted[:res["n_copy"]])
            for r in range(1, self.world):
                off, shift, carry, byte0, n_copy = res["placement"][r]
                if n_copy:
                    ops.append(dist.P2POp(dist.irecv, dst[byte0:byte0 + n_copy], r))
        elif res["n_copy"]:
            ops.append(dist.P2POp(dist.isend, self.shifted[:res["n_copy"]], 0))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        if self.rank != 0:
            return 0
        end_bit = res["end_bit"]
        last = int(dst[end_bit >> 3]) if end_bit & 7 else 0
        tr = trailer_bytes(end_bit, last, res["crc"])
        head = torch.tensor(list(b"BZh" + bytes([ord("0") + self.comp.multiplier])), dtype=torch.uint8)
        dst[:4].copy_(head)
        dst[end_bit >> 3:(end_bit >> 3) + len(tr)].copy_(torch.tensor(list(tr), dtype=torch.uint8))
        return (end_bit >> 3) + len(tr)
