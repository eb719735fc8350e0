"""Multi-GPU host logic: blocks share no state (Thesis p.30), so a stream is sharded by CONTIGUOUS BLOCK RANGES,
one process per GPU, with no data-path collective.  The only exchange is metadata: the block start offsets (found by one
boundary scan, bz2b200_scan_blocks) going out, and each rank's packed bit stream + block CRCs coming back to rank 0, which
concatenates them at bit granularity and folds the stream CRC (OutputStream.hpp:202,225-239)."""
from __future__ import annotations

import numpy as np


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
