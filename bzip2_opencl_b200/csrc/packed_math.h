// packed_math.h -- the arithmetic of the round-0 keys over a block's own alphabet (bwt.cuh: k_bwt_count_packed, first sweep of
// k_radix_onesweep).  A block uses `alpha` of the 256 byte values (BlockCompressor.hpp:136,148: blockValuesPresent); mapped to
// 0..alpha-1 in byte order a symbol needs bits = ceil(log2 alpha) bits (at least 1), and the 40 key bits of a sort element hold the
// first floor(40 / bits) symbols of the rotation (at most 16) plus the leading bits of the next one.  Mapping and truncation are
// monotone, so key order is rotation order on the first h0 bytes, refined further by the partial symbol -- a refinement that is
// consistent with the final order, which is all prefix doubling needs.  Host + device (tests/host/test_packed_keys.cpp).
#pragma once
#include <stdint.h>
#if defined(__CUDACC__)
#define BZ_HD3 __host__ __device__ __forceinline__
#else
#define BZ_HD3 inline
#endif
namespace bz {
constexpr int PK_KEY_BITS = 40;  // key bits of a 64-bit sort element (20 more hold the rotation index)
constexpr int PK_MAX_SYMS = 16;  // symbols per key at most
struct PackedShape {
    int bits;   // bits per symbol
    int hc;     // symbols that reach into the key (the last one possibly only with its leading bits)
    int width;  // hc * bits
    int h0;     // whole symbols in the key = bytes the rotations are ordered by after round 0
};
BZ_HD3 PackedShape packed_shape(uint32_t alpha)
{
    PackedShape s;
    s.bits = 1;
    while ((1u << s.bits) < alpha) ++s.bits;
    const int reach = (PK_KEY_BITS + s.bits - 1) / s.bits, whole = PK_KEY_BITS / s.bits;
    s.hc = reach < PK_MAX_SYMS ? reach : PK_MAX_SYMS;
    s.width = s.hc * s.bits;
    s.h0 = whole < PK_MAX_SYMS ? whole : PK_MAX_SYMS;
    return s;
}
// sliding window: c holds the last hc symbols, newest in the low bits
BZ_HD3 uint64_t packed_push(uint64_t c, uint32_t sym, const PackedShape &s) { return ((c << s.bits) | sym) & ((1ull << s.width) - 1ull); }
// the 40-bit key of a full window
BZ_HD3 uint64_t packed_key_of(uint64_t c, const PackedShape &s)
{
    return s.width >= PK_KEY_BITS ? c >> (s.width - PK_KEY_BITS) : c << (PK_KEY_BITS - s.width);
}
} // namespace bz
