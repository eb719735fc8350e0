// rle.cuh -- stage A: block boundaries + RLE1 + symbol presence + block CRC32 on the device.
//
// Replaces the reference's per-byte host loop: app.cpp:108-116 -> OutputStream::write(int) (OutputStream.hpp:131-142)
// -> BlockCompressor::write / writeRun / finishRLE (BlockCompressor.hpp:69-154) and CRC32::updateCRC (CRC32.hpp:75-86).
// See rle_seg.h for the algebra.  Kernels:
//   k_rle_tile_seg      one warp summarises one 2 KiB tile of the raw input                  (parallel)
//   k_rle_reduce        64-ary tree of summaries                                             (parallel)
//   k_rle_walk          ONE warp walks the blocks through the tree -> block start offsets     (sequential over blocks)
//   k_rle_btile_seg     summaries of block-relative tiles                                    (parallel)
//   k_rle_btile_scan    one warp per block: machine state at every block-relative tile start  (parallel over blocks)
//   k_rle_emit          one warp per block-relative tile writes the RLE1 bytes + presence bits(parallel)
//   k_crc_chunks        per-block CRC32 by chunk CRCs combined with x^(8k) mod P shifts       (parallel)
#pragma once
#include "util.cuh"
#include "rle_seg.h"
#include "crc_math.h"

namespace bz {

constexpr int RLE_SM_STRIDE = 68; // bytes of shared memory per lane (64 + 4 pad: conflict-free lane-strided byte reads)
constexpr int RLE_WARPS = 8;

__device__ __forceinline__ Seg shfl_seg(const Seg &s, int src)
{
    Seg r;
    r.len = shfl_u64(s.len, src);
    r.lead = shfl_u64(s.lead, src);
    r.trail = shfl_u64(s.trail, src);
    r.mid = shfl_u64(s.mid, src);
    r.first = __shfl_sync(FULL, s.first, src);
    r.last = __shfl_sync(FULL, s.last, src);
    r.uniform = __shfl_sync(FULL, s.uniform, src);
    r.pad = 0;
    return r;
}
__device__ __forceinline__ Seg shfl_up_seg(const Seg &s, int d)
{
    Seg r;
    r.len = shfl_up_u64(s.len, d);
    r.lead = shfl_up_u64(s.lead, d);
    r.trail = shfl_up_u64(s.trail, d);
    r.mid = shfl_up_u64(s.mid, d);
    r.first = __shfl_up_sync(FULL, s.first, d);
    r.last = __shfl_up_sync(FULL, s.last, d);
    r.uniform = __shfl_up_sync(FULL, s.uniform, d);
    r.pad = 0;
    return r;
}
__device__ __forceinline__ Seg warp_seg_incl_scan(Seg s)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Seg o = shfl_up_seg(s, d);
        if (lane_id() >= d) s = seg_combine(o, s);
    }
    return s;
}
__device__ __forceinline__ RleState shfl_state(const RleState &s, int src)
{
    RleState r;
    r.bl = shfl_u64(s.bl, src);
    r.cur = __shfl_sync(FULL, s.cur, src);
    r.run = __shfl_sync(FULL, s.run, src);
    return r;
}

// Stage n (<= RLE_TILE) bytes in[lo..lo+n) into the warp's padded shared-memory tile.
__device__ __forceinline__ void warp_load_tile(const uint8_t *__restrict__ in, uint64_t lo, uint32_t n, uint8_t *sm)
{
    const int lane = lane_id();
    if (n == RLE_TILE && ((lo & 15ull) == 0) && ((reinterpret_cast<uintptr_t>(in) & 15u) == 0)) {
        const uint4 *src = reinterpret_cast<const uint4 *>(in + lo);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int idx = k * 32 + lane; // uint4 index inside the tile
            uint4 v = __ldg(src + idx);
            int byte = idx * 16;
            uint32_t *dst = reinterpret_cast<uint32_t *>(sm + (byte >> 6) * RLE_SM_STRIDE + (byte & 63));
            dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
        }
    } else {
        for (uint32_t k = lane; k < n; k += 32) sm[(k >> 6) * RLE_SM_STRIDE + (k & 63)] = __ldg(in + lo + k);
    }
    __syncwarp();
}

// ---- the same staging through the TMA engine (Blackwell / Hopper bulk copy) --------------------------------------------------
// Block-relative tiles start at arbitrary bytes of the input, so the generic path above falls back to byte loads (64 LDG.U8 + 64
// STS.U8 per lane).  Instead ONE lane asks the copy engine for the 16-byte aligned span that covers the tile
// (cp.async.bulk.shared::cluster.global, completion on an mbarrier: SASS UBLKCP + SYNCS), and the warp then moves the bytes from that
// linear landing buffer into the padded tile, 32 consecutive words per step, shifted by the tile's offset inside its first word.
//   lin   16-byte aligned landing buffer of RLE_BULK_BYTES bytes (per warp);  bar  the warp's mbarrier (8 bytes, initialised by
//   bulk_init); phase  parity of this use (0 for the first copy through the barrier, then alternating)
constexpr uint32_t RLE_BULK_BYTES = RLE_TILE + 32;
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_init(uint64_t *bar)
{
    if (lane_id() == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
}
__device__ __forceinline__ void warp_load_tile_bulk(const uint8_t *__restrict__ in, uint64_t lo, uint32_t n, uint8_t *sm, uint8_t *lin, uint64_t *bar, uint32_t phase)
{
    const int lane = lane_id();
    const uint64_t a0 = lo & ~15ull;
    const uint32_t skew = (uint32_t)(lo - a0);
    const uint32_t bytes = (skew + n + 15u) & ~15u;
    if (lane == 0) {
        const uint32_t b = smem_addr(bar);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(lin)), "l"(in + a0), "r"(bytes), "r"(b)
                     : "memory");
    }
    {
        const uint32_t b = smem_addr(bar);
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(b), "r"(phase) : "memory");
        }
    }
    // linear -> padded (RLE_SM_STRIDE bytes per 64): word k of the tile = bytes skew + 4k .. of the landing buffer
    const uint32_t *lw = reinterpret_cast<const uint32_t *>(lin);
    const uint32_t w0 = skew >> 2, sh = 8u * (skew & 3u);
    const uint32_t nw = (n + 3u) >> 2;
    for (uint32_t k0 = 0; k0 < nw; k0 += 32) {
        const uint32_t k = k0 + (uint32_t)lane;
        const uint32_t x = lw[w0 + min(k, nw)]; // (word nw is the last one the copy can have touched: skew + n <= 4 (w0 + nw) + 3)
        uint32_t y = __shfl_down_sync(FULL, x, 1);
        if (lane == 31) y = lw[w0 + min(k + 1, nw)];
        if (k < nw) *reinterpret_cast<uint32_t *>(sm + (k >> 4) * RLE_SM_STRIDE + ((k & 15u) << 2)) = __funnelshift_r(x, y, sh);
    }
    __syncwarp();
}

__device__ __forceinline__ uint32_t lane_count(uint32_t n)
{
    int c = (int)n - lane_id() * (int)RLE_LANE_BYTES;
    return c < 0 ? 0u : (c > (int)RLE_LANE_BYTES ? RLE_LANE_BYTES : (uint32_t)c);
}

// Bit i (1..63) of the result: byte i of the lane's 64 staged bytes equals byte i-1 (bit 0 is 0).  Four bytes per step:
// per-byte SIMD compare of a word with itself shifted by one byte, the four result bytes gathered with one multiply.
__device__ __forceinline__ uint64_t lane_eq_mask(const uint8_t *p)
{
    const uint32_t *w = reinterpret_cast<const uint32_t *>(p); // lane offsets are multiples of 4
    uint64_t eq = 0;
    uint32_t prev = ~w[0] << 24; // byte "-1" differs from byte 0
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const uint32_t x = w[k];
        const uint32_t d = __vcmpeq4(x, __funnelshift_l(prev, x, 8)); // 0xff where byte j == byte j-1
        eq |= (uint64_t)(((d & 0x01010101u) * 0x01020408u) >> 24) << (4 * k);
        prev = x;
    }
    return eq;
}
// no run of four or more equal bytes inside a full lane
__device__ __forceinline__ bool lane_short_runs(uint64_t eq) { return (eq & (eq >> 1) & (eq >> 2)) == 0; }

__device__ __forceinline__ Seg lane_seg(const uint8_t *sm, uint32_t cnt)
{
    const uint8_t *p = sm + lane_id() * RLE_SM_STRIDE;
    if (cnt == RLE_LANE_BYTES) return seg_from_eq_mask(p[0], p[RLE_LANE_BYTES - 1], lane_eq_mask(p)); // (rle_seg.h)
    return seg_from_bytes([p](uint32_t i) { return (uint32_t)p[i]; }, cnt);
}

// ---------------------------------------------------------------- level 0 summaries of the raw input
// ltab (may be null): the summary of the lanes IN FRONT of every lane of the tile, packed into 64 bits (pack_lane_prefix) -- the walk
// reads the machine's state at any 64-byte boundary of a tile from it instead of staging and scanning the tile again.
__device__ __forceinline__ uint64_t pack_lane_prefix(const Seg &s)
{
    return (uint64_t)s.lead | ((uint64_t)s.trail << 12) | ((uint64_t)s.mid << 24) | ((uint64_t)(s.first & 0xffu) << 36) | ((uint64_t)(s.last & 0xffu) << 44);
}
__device__ __forceinline__ TSeg unpack_lane_prefix(uint64_t v, uint32_t len)
{
    TSeg t;
    t.len = len;
    t.lead = (uint32_t)v & 0xfffu; t.trail = (uint32_t)(v >> 12) & 0xfffu; t.mid = (uint32_t)(v >> 24) & 0xfffu;
    t.first = (uint32_t)(v >> 36) & 0xffu; t.last = (uint32_t)(v >> 44) & 0xffu;
    return t;
}
__global__ void __launch_bounds__(RLE_WARPS * 32) k_rle_tile_seg(const uint8_t *__restrict__ in, uint64_t n_bytes, Seg *__restrict__ seg0, uint64_t tile0,
                                                                   uint64_t n_tiles, uint64_t *__restrict__ ltab)
{
    __shared__ __align__(16) uint8_t sm_all[RLE_WARPS][32 * RLE_SM_STRIDE];
    uint64_t tile = tile0 + (uint64_t)blockIdx.x * RLE_WARPS + warp_id();
    if (tile >= n_tiles) return;
    uint8_t *sm = sm_all[warp_id()];
    uint64_t lo = tile * RLE_TILE;
    uint32_t n = (uint32_t)min((uint64_t)RLE_TILE, n_bytes - lo);
    warp_load_tile(in, lo, n, sm);
    Seg s = warp_seg_incl_scan(lane_seg(sm, lane_count(n)));
    if (lane_id() == 31) seg0[tile] = s;
    if (ltab) {
        Seg ex = shfl_up_seg(s, 1);
        if (lane_id() == 0) ex = seg_empty();
        uint64_t v = pack_lane_prefix(ex); // (lanes behind a partial lane are empty and never looked at)
        // bit 63 of lane 0's entry: a full tile in which no run of four or more equal bytes forms (runs across lanes included; the
        // run carried in from the tile in front is the walk's to check) -- the walk then needs no scan at all (short_run_search)
        bool bad = n != RLE_TILE;
        if (!bad) {
            const uint8_t *p = sm + lane_id() * RLE_SM_STRIDE;
            const uint64_t eq = lane_eq_mask(p);
            const uint32_t prev_last = __shfl_up_sync(FULL, (uint32_t)p[RLE_LANE_BYTES - 1], 1);
            const uint32_t prev_top = __shfl_up_sync(FULL, (uint32_t)(eq >> 62), 1);
            const bool b0 = lane_id() > 0 && prev_last == (uint32_t)p[0];
            bad = !lane_short_runs(eq) || (b0 && prev_top == 3u) || (b0 && (prev_top & 2u) && (eq & 2ull)) || (b0 && (eq & 6ull) == 6ull);
        }
        const bool short_tile = !__any_sync(FULL, bad);
        if (lane_id() == 0 && short_tile) v |= 1ull << 63;
        ltab[tile * 32 + lane_id()] = v;
    }
}

// ---------------------------------------------------------------- tree level k -> k+1
__global__ void __launch_bounds__(RLE_WARPS * 32) k_rle_reduce(const Seg *__restrict__ child, uint64_t n_child, Seg *__restrict__ parent, uint64_t n_parent)
{
    uint64_t p = (uint64_t)blockIdx.x * RLE_WARPS + warp_id();
    if (p >= n_parent) return;
    uint64_t c0 = p * RLE_FANOUT + 2 * lane_id();
    Seg a = c0 < n_child ? child[c0] : seg_empty();
    Seg b = c0 + 1 < n_child ? child[c0 + 1] : seg_empty();
    Seg s = warp_seg_incl_scan(seg_combine(a, b));
    if (lane_id() == 31) parent[p] = s;
}

// ---------------------------------------------------------------- the sequential walk (one warp)
struct WalkLevels {
    const Seg *seg[RLE_MAX_LEVELS];
    uint64_t cnt[RLE_MAX_LEVELS];
    int n_levels;
};

// inclusive warp scan of 32-bit summaries: two shuffles per field pair instead of one per 64-bit half
__device__ __forceinline__ TSeg shfl_up_tseg(const TSeg &s, int d)
{
    // len/lead/trail/mid fit 13 bits, first/last 8: three words
    const uint32_t a = s.len | (s.lead << 13) | ((s.first & 0x3fu) << 26), b = s.trail | (s.mid << 13) | ((s.last & 0x3fu) << 26),
                   c = (s.first >> 6) | ((s.last >> 6) << 2);
    const uint32_t a2 = __shfl_up_sync(FULL, a, d), b2 = __shfl_up_sync(FULL, b, d), c2 = __shfl_up_sync(FULL, c, d);
    TSeg r;
    r.len = a2 & 0x1fffu; r.lead = (a2 >> 13) & 0x1fffu; r.first = (a2 >> 26) | ((c2 & 3u) << 6);
    r.trail = b2 & 0x1fffu; r.mid = (b2 >> 13) & 0x1fffu; r.last = (b2 >> 26) | (((c2 >> 2) & 3u) << 6);
    return r;
}
__device__ __forceinline__ TSeg warp_tseg_incl_scan(TSeg s)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const TSeg o = shfl_up_tseg(s, d);
        if (lane_id() >= d) s = tseg_combine(o, s);
    }
    return s;
}

// The walk's tile simulation: bytes [lo, hi) of one tile with the machine in state `st` at lo -> first refused byte (st.bl > limit
// in front of it) and the state there, or hi and the state behind the last byte.  One warp.
//   * a full tile in which no run of four or more equal bytes forms (text): the state in front of every byte is closed form
//     (short_run_search, rle_seg.h), nothing is scanned;
//   * otherwise the lanes' 64-byte summaries are scanned (32-bit fields) and the ONE lane that can hold the refusal is scanned again
//     with two bytes per lane.
struct DevTileScan {
    const uint8_t *in;
    uint8_t *sm;
    const uint64_t *ltab; // lane prefix table of k_rle_tile_seg (null: every tile is staged and scanned here)

    // the 64 bytes (cnt_c of them valid) of one lane, two per thread: the refusal inside them, if any.  v0, v1 = bytes 2 lane, 2 lane + 1.
    __device__ __forceinline__ bool inside_lane(uint32_t v0, uint32_t v1, uint32_t cnt_c, const RleState &sc, uint64_t limit, uint32_t &fi, RleState &fs,
                                                RleState *behind) const
    {
        const int lane = lane_id();
        const uint32_t i0 = 2u * (uint32_t)lane;
        TSeg piece;
        piece.len = i0 + 1 < cnt_c ? 2u : (i0 < cnt_c ? 1u : 0u);
        piece.first = v0;
        piece.last = piece.len == 2u ? v1 : v0;
        piece.lead = piece.trail = (piece.len == 2u && v0 != v1) ? 1u : piece.len;
        piece.mid = 0;
        const TSeg pin = warp_tseg_incl_scan(piece);
        TSeg pex = shfl_up_tseg(pin, 1);
        if (lane == 0) pex = tseg_of(seg_empty());
        const RleState e0 = tstate_apply(pex, sc); // in front of byte i0
        RleState e1 = e0;                          // in front of byte i0 + 1
        if (i0 < cnt_c) state_step(e1, v0);
        if (behind) *behind = shfl_state(tstate_apply(pin, sc), 31); // behind the lane's last byte
        const bool h0 = i0 >= 1 && i0 < cnt_c && e0.bl > limit, h1 = i0 + 1 < cnt_c && e1.bl > limit;
        const unsigned hb = __ballot_sync(FULL, h0 || h1);
        if (!hb) return false;
        const int w = __ffs((int)hb) - 1;
        const uint32_t my_i = h0 ? i0 : i0 + 1;
        fi = __shfl_sync(FULL, my_i, w);
        fs = shfl_state(h0 ? e0 : e1, w);
        return true;
    }

    __device__ TileScanResult operator()(uint64_t lo, uint64_t hi, RleState st, uint64_t limit) const
    {
        const int lane = lane_id();
        const uint32_t n = (uint32_t)(hi - lo);
        if (ltab && n == RLE_TILE && (lo % RLE_TILE) == 0) {
            // a whole tile: the state in front of every lane from the table, then the one lane that can hold the refusal, two bytes per thread
            TileScanResult r;
            const uint64_t raw = __ldg(ltab + (lo / RLE_TILE) * 32 + lane);
            // (loaded next to the table: the 32 bytes around the place where a tile without long runs would refuse -- the state in front
            // of byte i is then st.bl + st.run + i - pending run, short_run_search in rle_seg.h)
            const int64_t K = (int64_t)limit - (int64_t)(st.bl + st.run);
            const uint32_t wbase = K < 4 ? 0u : (K - 4 > (int64_t)(RLE_TILE - 32) ? RLE_TILE - 32 : (uint32_t)(K - 4));
            const uint32_t wbyte = __ldg(in + lo + wbase + lane);
            const TSeg excl = unpack_lane_prefix(raw, 64u * (uint32_t)lane);
            if (__shfl_sync(FULL, (uint32_t)(raw >> 63), 0) && st.run < 4) {
                const uint32_t first = __shfl_sync(FULL, excl.first, 1), lead0 = __shfl_sync(FULL, excl.lead, 1); // lane 1's prefix = lane 0's summary
                if (!(st.run > 0 && st.cur == first) || st.run + lead0 < 4) {
                    uint32_t pos;
                    r.st = short_run_search([wbyte, wbase](uint32_t j) { return __shfl_sync(FULL, wbyte, (int)(j - wbase)); }, RLE_TILE, st, limit, 0u, pos);
                    r.pos = lo + pos;
                    return r;
                }
            }
            const RleState s0 = tstate_apply(excl, st);
            const unsigned over = __ballot_sync(FULL, s0.bl > limit);
            const int f = over ? __ffs((int)over) - 1 : 32;
            const int c = f - 1;
            if (c >= 0) {
                const uint16_t two = __ldg(reinterpret_cast<const uint16_t *>(in + lo + (uint64_t)c * RLE_LANE_BYTES) + lane); // (tiles are 16-byte aligned)
                uint32_t fi;
                RleState fs, behind;
                if (inside_lane(two & 0xffu, two >> 8, RLE_LANE_BYTES, shfl_state(s0, c), limit, fi, fs, &behind)) {
                    r.pos = lo + (uint64_t)c * RLE_LANE_BYTES + fi;
                    r.st = fs;
                    return r;
                }
                if (!over) { r.pos = hi; r.st = behind; return r; }
            }
            r.pos = lo + (uint64_t)f * RLE_LANE_BYTES; // at the first byte of lane f
            r.st = shfl_state(s0, f);
            return r;
        }
        __syncwarp();
        warp_load_tile(in, lo, n, sm);
        const uint32_t cnt = lane_count(n);
        const uint8_t *p = sm + lane * RLE_SM_STRIDE;
        const uint8_t *tile = sm;
        auto get = [tile](uint32_t j) { return (uint32_t)tile[(j >> 6) * RLE_SM_STRIDE + (j & 63u)]; };
        TileScanResult r;
        uint64_t eq = 0;
        bool lane_short = false; // a full lane without a run of four (the carried run not considered)
        if (cnt == RLE_LANE_BYTES) {
            eq = lane_eq_mask(p);
            lane_short = lane_short_runs(eq);
        }
        if (n == RLE_TILE) {
            // no run of four anywhere, joins across lanes and with the carried run included?
            const uint32_t my_last = p[RLE_LANE_BYTES - 1];
            uint32_t prev_last = __shfl_up_sync(FULL, my_last, 1);
            uint32_t prev_top = __shfl_up_sync(FULL, (uint32_t)(eq >> 62), 1); // bit 1: bytes 63,62 equal; bit 0: bytes 62,61 equal
            bool prev_any = true;
            if (lane == 0) {
                prev_any = st.run > 0;
                prev_last = st.cur;
                prev_top = (st.run >= 2 ? 2u : 0u) | (st.run >= 3 ? 1u : 0u);
            }
            const bool b0 = prev_any && prev_last == (uint32_t)p[0];
            // (lane > 0: the previous lane's bit 0 is not in prev_top, but a run through a whole lane is not short anyway)
            bool bad = !lane_short || (lane == 0 && st.run >= 4);
            bad |= b0 && (prev_top == 3u);                                  // x x x | x
            bad |= b0 && (prev_top & 2u) && (eq & 2ull);                    // x x | x x
            bad |= b0 && ((eq & 6ull) == 6ull);                             // x | x x x
            if (!__any_sync(FULL, bad)) {
                uint32_t pos;
                r.st = short_run_search(get, n, st, limit, 0u, pos);
                r.pos = lo + pos;
                __syncwarp();
                return r;
            }
        }
        const TSeg mine = tseg_of(cnt == RLE_LANE_BYTES ? seg_from_eq_mask(p[0], p[RLE_LANE_BYTES - 1], eq) : lane_seg(sm, cnt));
        const TSeg incl = warp_tseg_incl_scan(mine);
        TSeg excl = shfl_up_tseg(incl, 1);
        if (lane == 0) excl = tseg_of(seg_empty());
        const RleState s0 = tstate_apply(excl, st); // in front of the lane's first byte
        const unsigned have = __ballot_sync(FULL, cnt > 0);
        const int z = 31 - __clz((int)have);       // last lane with bytes
        const unsigned over = __ballot_sync(FULL, cnt > 0 && s0.bl > limit);
        // the refusal is at the first byte of the first lane that starts over the limit, or inside the lane in front of it; with no
        // such lane, inside the last lane or nowhere
        const int f = over ? __ffs((int)over) - 1 : z + 1;
        const int c = f - 1; // the one lane whose interior (bytes 1 .. cnt-1) can hold the refusal
        bool found = false;
        uint32_t fi = 0;
        RleState fs = s0;
        const uint32_t cnt_c = c >= 0 ? __shfl_sync(FULL, cnt, c) : 0u;
        if (cnt_c > 1) { // (warp-uniform) lane c's bytes two per lane: the same scan one level down gives the state in front of each of them
            const uint8_t *q = sm + c * RLE_SM_STRIDE;
            const uint32_t i0 = 2u * (uint32_t)lane;
            found = inside_lane(i0 < cnt_c ? q[i0] : 0u, i0 + 1 < cnt_c ? q[i0 + 1] : 0u, cnt_c, shfl_state(s0, c), limit, fi, fs, nullptr);
        }
        if (found) { // inside lane c
            r.pos = lo + (uint64_t)c * RLE_LANE_BYTES + fi;
            r.st = fs;
        } else if (over) { // at the first byte of lane f
            r.pos = lo + (uint64_t)f * RLE_LANE_BYTES;
            r.st = shfl_state(s0, f);
        } else { // nowhere: the state behind the last byte
            r.pos = hi;
            r.st = shfl_state(tstate_apply(mine, s0), z);
        }
        __syncwarp();
        return r;
    }
};

// A sibling group (<= 64 nodes of one level) in one go: lane l loads nodes 2l and 2l+1, a warp scan gives the machine
// state in front of every node, a ballot finds the first node that overflows the block.
struct DevSiblingStep {
    WalkLevels lv;
    __device__ bool operator()(int level, uint64_t i0, uint64_t iend, RleState &st, uint64_t limit, uint64_t &found) const
    {
        const int lane = lane_id();
        const Seg *nodes = lv.seg[level];
        const uint64_t ia = i0 + 2 * lane, ib = ia + 1;
        Seg a = ia < iend ? nodes[ia] : seg_empty();
        Seg b = ib < iend ? nodes[ib] : seg_empty();
        Seg incl = warp_seg_incl_scan(seg_combine(a, b));
        Seg excl = shfl_up_seg(incl, 1);
        if (lane == 0) excl = seg_empty();
        RleState sa = state_apply(excl, st), sb = state_apply(a, sa), sc = state_apply(b, sb);
        const bool oa = a.len && sb.bl > limit;
        const bool ob = !oa && b.len && sc.bl > limit;
        const unsigned bal = __ballot_sync(FULL, oa || ob);
        if (bal) {
            const int f = __ffs(bal) - 1;
            const int first_is_a = __shfl_sync(FULL, (int)oa, f);
            found = i0 + 2 * (uint64_t)f + (first_is_a ? 0 : 1);
            RleState pick = oa ? sa : sb;
            st = shfl_state(pick, f);
            return true;
        }
        st = shfl_state(sc, 31);
        return false;
    }
};

// Resumable: info[0] = blocks found so far, info[1] = input position reached, info[2] = 1 once the input is exhausted (or,
// for a non-final call, the remaining bytes do not complete a block).  One launch finds at most `max_new` further blocks, so
// the host can run it on a side stream one batch ahead of the compression kernels.
// starts[b] = start offset of block b, starts[info[0]] = position reached; rle_len[b] = RLE1 length of block b.
__global__ void __launch_bounds__(32) k_rle_walk(const uint8_t *__restrict__ in, uint64_t n_bytes, uint64_t limit, int is_final, WalkLevels lv,
                                                  uint64_t *__restrict__ starts, uint32_t *__restrict__ rle_len, uint64_t cap, uint64_t *__restrict__ info,
                                                  uint64_t max_new)
{
    __shared__ __align__(16) uint8_t sm[32 * RLE_SM_STRIDE];
    DevTileScan scan{in, sm, nullptr};
    DevSiblingStep step{lv};
    uint64_t nb = info[0], s = info[1];
    const uint64_t target = nb + max_new;
    uint64_t finished = 0;
    __syncwarp();
    while (nb < target) {
        if (s >= n_bytes) { finished = 1; break; }
        RleState st;
        uint64_t e = walk_block_wide(s, n_bytes, limit, lv.n_levels, lv.cnt, step, scan, st);
        bool complete = (e < n_bytes) || is_final; // e == n_bytes without a refusal: flushed only when the stream ends
        if (!complete) { finished = 1; break; }
        if (lane_id() == 0 && nb < cap) {
            starts[nb] = s;
            rle_len[nb] = (uint32_t)state_final_length(st);
        }
        ++nb;
        s = e;
    }
    if (s >= n_bytes) finished = 1;
    __syncwarp();
    if (lane_id() == 0) {
        if (nb <= cap) starts[nb] = s;
        info[0] = nb;
        info[1] = s;
        info[2] = finished;
    }
}

// ---------------------------------------------------------------- prefix states + the walk on them
// pre[l][i] = summary of everything in front of node i of level l.  One warp per parent: exclusive scan of its <= 64 children
// on top of the parent's own prefix.  Level 0 stores the global machine's STATE at the tile start instead (16 bytes),
// pstate[n_tiles] = state at the end of the available input.
struct PState { uint64_t bl; uint32_t cur, run; };

__global__ void __launch_bounds__(RLE_WARPS * 32) k_rle_prefix(const Seg *__restrict__ child, uint64_t n_child, const Seg *__restrict__ parent_pre,
                                                                uint64_t n_parent, Seg *__restrict__ child_pre, PState *__restrict__ pstate)
{
    uint64_t p = (uint64_t)blockIdx.x * RLE_WARPS + warp_id();
    if (p >= n_parent) return;
    const int lane = lane_id();
    uint64_t c0 = p * RLE_FANOUT + 2 * lane;
    Seg a = c0 < n_child ? child[c0] : seg_empty();
    Seg b = c0 + 1 < n_child ? child[c0 + 1] : seg_empty();
    Seg incl = warp_seg_incl_scan(seg_combine(a, b));
    Seg excl = shfl_up_seg(incl, 1);
    if (lane == 0) excl = seg_empty();
    const Seg base = parent_pre ? parent_pre[p] : seg_empty();
    const Seg pa = seg_combine(base, excl), pb = seg_combine(pa, a);
    if (pstate) {
        if (c0 < n_child) { RleState st = state_apply(pa, state_fresh()); PState o{st.bl, st.cur, st.run}; pstate[c0] = o; }
        if (c0 + 1 < n_child) { RleState st = state_apply(pb, state_fresh()); PState o{st.bl, st.cur, st.run}; pstate[c0 + 1] = o; }
        // the state behind the last tile
        if (c0 + 1 == n_child - 1 || c0 == n_child - 1) {
            RleState st = state_apply(seg_combine(pb, b), state_fresh());
            PState o{st.bl, st.cur, st.run};
            pstate[n_child] = o;
        }
    } else {
        if (c0 < n_child) child_pre[c0] = pa;
        if (c0 + 1 < n_child) child_pre[c0 + 1] = pb;
    }
}

struct DevRunEnd {
    const uint8_t *in;
    uint64_t n;
    WalkLevels lv;
    // first q > s with in[q] != in[s], or n.  32 bytes per step inside the tile of s, then whole uniform nodes of the tree.
    __device__ uint64_t operator()(uint64_t s) const
    {
        const int lane = lane_id();
        const uint32_t v = __ldg(in + s);
        const uint64_t tile_end = min(n, (s / RLE_TILE + 1) * RLE_TILE);
        for (uint64_t q0 = s + 1; q0 < tile_end; q0 += 32) {
            const uint64_t q = q0 + lane;
            const bool diff = q < tile_end && __ldg(in + q) != v;
            const unsigned m = __ballot_sync(FULL, diff);
            if (m) return q0 + (uint64_t)(__ffs((int)m) - 1);
        }
        uint64_t t = tile_end / RLE_TILE; // first tile not yet looked at (tile_end is a tile boundary unless it is n)
        if (tile_end >= n) return n;
        for (;;) {
            if (t >= lv.cnt[0]) return n;
            // the largest aligned node starting at tile t that is one run of v
            int level = 0;
            uint64_t idx = t, span = 1;
            for (;;) {
                const Seg sg = lv.seg[level][idx];
                const bool ok = sg.uniform && sg.first == v;
                if (!ok) {
                    if (level == 0) { // the run ends inside (or right at the start of) this tile
                        if (sg.first != v) return t * RLE_TILE;
                        return min(n, t * RLE_TILE + sg.lead);
                    }
                    break; // fall back to the child level result found before
                }
                if (level + 1 < lv.n_levels && idx % RLE_FANOUT == 0 && idx / RLE_FANOUT < lv.cnt[level + 1]) {
                    const Seg up = lv.seg[level + 1][idx / RLE_FANOUT];
                    if (up.uniform && up.first == v) { ++level; idx /= RLE_FANOUT; span *= RLE_FANOUT; continue; }
                }
                break;
            }
            t += span;
        }
    }
};

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// v = in[s], r = end of the run of v that starts at s, vr = in[r]: one 32-byte look behind s settles nearly every block start
struct DevHead {
    const uint8_t *in;
    uint64_t n;
    DevRunEnd run_end;
    __device__ void operator()(uint64_t s, uint32_t &v, uint64_t &r, uint32_t &vr) const
    {
        const int lane = lane_id();
        const uint64_t q = s + (uint64_t)lane;
        const uint32_t mine = q < n ? (uint32_t)__ldg(in + q) : 0x100u; // (past the end: differs from every byte)
        v = __shfl_sync(FULL, mine, 0);
        const unsigned diff = __ballot_sync(FULL, mine != v);
        if (diff) {
            const int f = __ffs((int)diff) - 1;
            r = min(n, s + (uint64_t)f);
            vr = __shfl_sync(FULL, mine, f) & 0xffu;
            return;
        }
        r = run_end(s); // a run of 32 or more: tile by tile, then whole uniform nodes of the tree
        vr = r < n ? (uint32_t)__ldg(in + r) : 0u;
    }
};

struct DevPState {
    const PState *ps;
    __device__ RleState operator()(uint64_t t) const
    {
        const PState p = ps[t];
        RleState r;
        r.bl = p.bl; r.cur = p.cur; r.run = p.run;
        return r;
    }
};

// smallest t > t_lo with ps[t].bl > lim (ps[].bl is non-decreasing), n_tiles + 1 if none; before = ps[t - 1].
// Blocks of one input are about equally long, so the answer is looked for first in a window of 32 prefix states around
// "as many tiles further on as the last block took"; the window is loaded (preload) before the bytes at the block start are
// looked at, so that the two round trips overlap.  A miss falls back to 32 probes 16 tiles apart, then 16 in a row.
struct DevFirstOver {
    const PState *ps;
    uint64_t n_tiles;
    uint64_t w0;   // the preloaded window holds ps[w0 + lane]
    PState mine;
    bool have;
    __device__ void preload(uint64_t first)
    {
        w0 = first;
        have = true;
        const uint64_t t = min(n_tiles, first + (uint64_t)lane_id());
        const ulonglong2 q = __ldg(reinterpret_cast<const ulonglong2 *>(ps + t));
        mine.bl = q.x; mine.cur = (uint32_t)q.y; mine.run = (uint32_t)(q.y >> 32);
    }
    __device__ static RleState as_state(const PState &p) { RleState r; r.bl = p.bl; r.cur = p.cur; r.run = p.run; return r; }
    __device__ uint64_t operator()(uint64_t t_lo, uint64_t lim, RleState &before)
    {
        const int lane = lane_id();
        if (have && w0 + 31 <= n_tiles) {
            // lane f = first tile of the window that lies behind t_lo and is over the limit.  If f >= 1, the tile in front of it is
            // either t_lo itself or a tile behind t_lo that is not over the limit (and then, bl being monotone, no earlier one is):
            // the answer, with its predecessor's state in lane f - 1.  f == 0 says nothing about the tiles in front of the window.
            const bool over = (w0 + (uint64_t)lane > t_lo) && mine.bl > lim;
            const unsigned m = __ballot_sync(FULL, over);
            if (m) {
                const int f = __ffs((int)m) - 1;
                if (f >= 1) {
                    before.bl = shfl_u64(mine.bl, f - 1);
                    before.cur = __shfl_sync(FULL, mine.cur, f - 1);
                    before.run = __shfl_sync(FULL, mine.run, f - 1);
                    return w0 + (uint64_t)f;
                }
            } else if (w0 + 31 > t_lo) {
                t_lo = w0 + 31; // nothing up to the window's last tile is over the limit
            }
        }
        for (;;) {
            if (t_lo >= n_tiles) return n_tiles + 1;
            const uint64_t t = min(n_tiles, (uint64_t)(t_lo + 16u * (uint64_t)(lane + 1)));
            const unsigned m = __ballot_sync(FULL, ps[t].bl > lim);
            if (m) {
                const int f = __ffs((int)m) - 1;
                const uint64_t base = t_lo + 16ull * (uint64_t)f; // ps[base].bl <= lim (or base == t_lo)
                const uint64_t t2 = min(n_tiles, (uint64_t)(base + 1 + (uint64_t)(lane & 15)));
                const unsigned m2 = __ballot_sync(FULL, ps[t2].bl > lim) & 0xffffu;
                const uint64_t ans = base + 1 + (uint64_t)(__ffs((int)m2) - 1); // m2 != 0: the probe 16 further on was over the limit
                before = as_state(ps[ans - 1]);
                return ans;
            }
            t_lo = min(n_tiles, (uint64_t)(t_lo + 512ull));
        }
    }
};

// Resumable like k_rle_walk (same info[] protocol); info[12], info[13] carry the global machine's state at info[1].
__global__ void __launch_bounds__(32) k_rle_walk_prefix(const uint8_t *__restrict__ in, uint64_t n_bytes, uint64_t limit, int is_final, WalkLevels lv,
                                                         const PState *__restrict__ pstate, uint64_t *__restrict__ starts, uint32_t *__restrict__ rle_len,
                                                         uint64_t cap, uint64_t *__restrict__ info, uint64_t max_new, const uint64_t *__restrict__ ltab)
{
    __shared__ __align__(16) uint8_t sm[32 * RLE_SM_STRIDE];
    DevTileScan scan{in, sm, ltab};
    DevRunEnd run_end{in, n_bytes, lv};
    DevHead head{in, n_bytes, run_end};
    DevPState ps{pstate};
    const uint64_t n_tiles = (n_bytes + RLE_TILE - 1) / RLE_TILE;
    DevFirstOver first_over{pstate, n_tiles, 0, PState{0, 0, 0}, false};
    const int lane = lane_id();
    uint64_t nb = info[0], s = info[1];
    RleState g;
    g.bl = info[12]; g.cur = (uint32_t)(info[13] >> 32); g.run = (uint32_t)info[13];
    // info[3] != 0: the walk starts inside the buffer (a slice of a longer stream, see bz2b200_scan_blocks_ex): the global machine's
    // state at s is the prefix state of its tile pushed through the tile's bytes in front of s
    uint64_t seek = info[3];
    if (seek && s < n_bytes) {
        const uint64_t t0 = s / RLE_TILE;
        g = ps(t0);
        if (s % RLE_TILE) g = scan(t0 * RLE_TILE, s, g, ~0ull).st;
        seek = 0;
    }
    const uint64_t target = nb + max_new;
    uint64_t finished = 0;
    uint64_t dt = (limit + RLE_TILE) / RLE_TILE; // tiles the last block took (start: a block without runs)
    __syncwarp();
    while (nb < target) {
        if (s >= n_bytes) { finished = 1; break; }
        const uint64_t ts0 = s / RLE_TILE;
        // the window of prefix states where this block should end, and -- into L2 only -- the tiles around that place and around
        // the place where the NEXT block should end together with its window: every later round trip of the chain stops at L2
        {
            const uint64_t guess = ts0 + dt;
            first_over.preload(guess > 14 ? guess - 14 : 0);
            const uint64_t ta = guess > 1 ? guess - 1 : 0;                       // tiles ta .. ta+2: 48 lines of 128 B
            for (uint32_t l = lane; l < 48; l += 32) {
                const uint64_t off = ta * RLE_TILE + (uint64_t)l * 128;
                if (off < n_bytes) prefetch_l2(in + off);
            }
            const uint64_t g2 = guess + dt;
            const uint64_t tb = g2 > 2 ? g2 - 2 : 0;                             // next block: tiles tb .. tb+4, states tb-30 .. tb+34
            for (uint32_t l = lane; l < 80; l += 32) {
                const uint64_t off = tb * RLE_TILE + (uint64_t)l * 128;
                if (off < n_bytes) prefetch_l2(in + off);
            }
            if (lane < 8) {
                const uint64_t t8 = (tb > 30 ? tb - 30 : 0) + 8ull * (uint64_t)lane;
                if (t8 <= n_tiles) prefetch_l2(pstate + t8);
            } else if (ltab && lane < 16) { // the lane prefix tables (256 B each) of the tiles where this block and the next should end
                const uint64_t t2 = (lane < 12 ? ta : tb + 1) + (uint64_t)((lane - 8) & 3) / 2, half = (uint64_t)(lane & 1);
                if (t2 < n_tiles) prefetch_l2(ltab + t2 * 32 + half * 16);
            }
        }
        RleState st, g2 = g;
        const uint64_t e = walk_block_prefix(s, n_bytes, limit, g2, head, ps, first_over, scan, st);
        const bool complete = (e < n_bytes) || is_final; // e == n_bytes without a refusal: flushed only when the stream ends
        if (!complete) { finished = 1; break; }
        if (lane == 0 && nb < cap) {
            starts[nb] = s;
            rle_len[nb] = (uint32_t)state_final_length(st);
        }
        ++nb;
        dt = max((uint64_t)1, e / RLE_TILE - ts0);
        s = e;
        g = g2;
    }
    if (s >= n_bytes) finished = 1;
    __syncwarp();
    if (lane == 0) {
        if (nb <= cap) starts[nb] = s;
        info[0] = nb;
        info[1] = s;
        info[2] = finished;
        info[3] = seek;
        info[12] = g.bl;
        info[13] = ((uint64_t)g.cur << 32) | g.run;
    }
}

// ---------------------------------------------------------------- block-relative tiles
// Tile k of block b covers input [starts[b] + k*RLE_TILE, min(.., starts[b+1])).
__device__ __forceinline__ bool btile_range(const uint64_t *starts, uint32_t b, uint64_t k, uint64_t &lo, uint32_t &n)
{
    uint64_t s = starts[b], e = starts[b + 1];
    lo = s + k * RLE_TILE;
    if (lo >= e) return false;
    n = (uint32_t)min((uint64_t)RLE_TILE, e - lo);
    return true;
}

__global__ void __launch_bounds__(RLE_WARPS * 32) k_rle_btile_seg(const uint8_t *__restrict__ in, const uint64_t *__restrict__ starts, uint32_t b0,
                                                                    Seg *__restrict__ segB, uint64_t max_tiles, int bulk)
{
    __shared__ __align__(16) uint8_t sm_all[RLE_WARPS][32 * RLE_SM_STRIDE];
    __shared__ __align__(16) uint8_t lin_all[RLE_WARPS][RLE_BULK_BYTES];
    __shared__ __align__(8) uint64_t bar_all[RLE_WARPS];
    uint64_t k = (uint64_t)blockIdx.x * RLE_WARPS + warp_id();
    uint32_t bl = blockIdx.y;
    uint64_t lo; uint32_t n;
    if (k >= max_tiles || !btile_range(starts, b0 + bl, k, lo, n)) return;
    uint8_t *sm = sm_all[warp_id()];
    if (bulk) {
        bulk_init(&bar_all[warp_id()]);
        warp_load_tile_bulk(in, lo, n, sm, lin_all[warp_id()], &bar_all[warp_id()], 0u);
    } else
        warp_load_tile(in, lo, n, sm);
    Seg s = warp_seg_incl_scan(lane_seg(sm, lane_count(n)));
    if (lane_id() == 31) segB[(uint64_t)bl * max_tiles + k] = s;
}

__device__ __forceinline__ uint64_t pack_state(const RleState &s) { return (s.bl & 0xffffffffull) | ((uint64_t)(s.run & 0xff) << 32) | ((uint64_t)(s.cur & 0xff) << 40); }
__device__ __forceinline__ RleState unpack_state(uint64_t v)
{
    RleState s;
    s.bl = v & 0xffffffffull;
    s.run = (uint32_t)(v >> 32) & 0xff;
    s.cur = (uint32_t)(v >> 40) & 0xff;
    return s;
}

// one warp per block: state of the machine at the first byte of every block-relative tile
__global__ void __launch_bounds__(RLE_WARPS * 32) k_rle_btile_scan(const Seg *__restrict__ segB, const uint64_t *__restrict__ starts, uint32_t b0, uint32_t nb,
                                                                     uint64_t *__restrict__ tstate, uint64_t max_tiles)
{
    uint32_t bl = blockIdx.x * RLE_WARPS + warp_id();
    if (bl >= nb) return;
    uint64_t len = starts[b0 + bl + 1] - starts[b0 + bl];
    uint64_t nt = (len + RLE_TILE - 1) / RLE_TILE;
    Seg carry = seg_empty();
    const int lane = lane_id();
    for (uint64_t k0 = 0; k0 < nt; k0 += 32) {
        uint64_t k = k0 + lane;
        Seg mine = k < nt ? segB[(uint64_t)bl * max_tiles + k] : seg_empty();
        Seg incl = warp_seg_incl_scan(mine);
        Seg excl = shfl_up_seg(incl, 1);
        if (lane == 0) excl = seg_empty();
        if (k < nt) tstate[(uint64_t)bl * max_tiles + k] = pack_state(state_apply(seg_combine(carry, excl), state_fresh()));
        carry = seg_combine(carry, shfl_seg(incl, 31));
    }
}

__device__ __forceinline__ void bm_set(uint32_t (&bm)[8], uint32_t v)
{
#pragma unroll
    for (int w = 0; w < 8; ++w) bm[w] |= (w == (int)(v >> 5)) ? (1u << (v & 31)) : 0u;
}

__device__ __forceinline__ void emit_run(uint8_t *dst, uint32_t v, uint32_t len, uint32_t (&)[8])
{
    // writeRun, BlockCompressor.hpp:134-154
    dst[0] = (uint8_t)v;
    if (len > 1) dst[1] = (uint8_t)v;
    if (len > 2) dst[2] = (uint8_t)v;
    if (len > 3) {
        dst[3] = (uint8_t)v;
        dst[4] = (uint8_t)(len - 4);
    }
}

// one warp per block-relative tile: RLE1 bytes into rle[bl*stride ...]
constexpr int RLE_OUT_STAGE = 2640; // bytes a tile can emit (2048 * 5/4 + carried run) + 16 of alignment slack, multiple of 16

__global__ void __launch_bounds__(RLE_WARPS * 32) k_rle_emit(const uint8_t *__restrict__ in, const uint64_t *__restrict__ starts, uint32_t b0,
                                                               const uint64_t *__restrict__ tstate, uint64_t max_tiles, uint8_t *__restrict__ rle,
                                                               uint64_t stride, int bulk)
{
    __shared__ __align__(16) uint8_t sm_all[RLE_WARPS][32 * RLE_SM_STRIDE];
    __shared__ __align__(16) uint8_t so_all[RLE_WARPS][RLE_OUT_STAGE];
    __shared__ __align__(8) uint64_t bar_all[RLE_WARPS];
    static_assert(RLE_OUT_STAGE >= RLE_BULK_BYTES, "the output staging area doubles as the bulk copy's landing buffer");
    uint64_t k = (uint64_t)blockIdx.x * RLE_WARPS + warp_id();
    uint32_t bl = blockIdx.y;
    uint64_t lo; uint32_t n;
    if (k >= max_tiles || !btile_range(starts, b0 + bl, k, lo, n)) return;
    const int lane = lane_id();
    uint8_t *sm = sm_all[warp_id()];
    if (bulk) { // the landing buffer is the output staging area: it is not written before the tile has been moved out of it
        bulk_init(&bar_all[warp_id()]);
        warp_load_tile_bulk(in, lo, n, sm, so_all[warp_id()], &bar_all[warp_id()], 0u);
    } else
        warp_load_tile(in, lo, n, sm);
    uint32_t cnt = lane_count(n);
    Seg incl = warp_seg_incl_scan(lane_seg(sm, cnt));
    Seg excl = shfl_up_seg(incl, 1);
    if (lane == 0) excl = seg_empty();
    RleState s = state_apply(excl, unpack_state(tstate[(uint64_t)bl * max_tiles + k]));
    const uint8_t *p = sm + lane * RLE_SM_STRIDE;
    uint32_t blen = (uint32_t)s.bl, cur = s.cur, run = s.run;
    // The tile writes the runs that END inside it: output bytes [base, end) of the block.  They are assembled in shared memory
    // at the same offset mod 16 as their final address, then stored with aligned 16-byte stores.
    const uint32_t base = __shfl_sync(FULL, blen, 0);
    uint8_t *so = so_all[warp_id()] + (base & 15u) - base; // so[x] <-> output byte x of the block
    // finishRLE (BlockCompressor.hpp:112-118): the lane holding the block's last byte flushes the pending run
    const bool last_lane = cnt > 0 && (lo + (uint64_t)lane * RLE_LANE_BYTES + cnt == starts[b0 + bl + 1]);
    bool done = false;
    if (cnt == RLE_LANE_BYTES && !last_lane && run < 4) {
        // no run of 4+ inside the lane and none completed by the bytes carried in: the RLE1 output of the runs that end in this
        // lane is the carried run followed by the lane's bytes up to its own (still pending) trailing run -- a plain copy
        const uint64_t eq = lane_eq_mask(p);
        const uint32_t lead = (uint32_t)__ffsll((long long)~(eq >> 1));
        if (lane_short_runs(eq) && (run == 0 || cur != p[0] || run + lead < 4)) {
            const uint32_t trail = 1u + (uint32_t)__clzll((long long)~eq);
            uint8_t *o = so + blen;
            for (uint32_t i = 0; i < run; ++i) o[i] = (uint8_t)cur;
            o += run;
            const uint32_t m = RLE_LANE_BYTES - trail;
            for (uint32_t i = 0; i < m; ++i) o[i] = p[i];
            blen += run + m;
            done = true;
        }
    }
    if (!done) {
        uint32_t bm[8]; // (presence comes from the byte histogram of the RLE1 block, k_radix_bases)
        for (uint32_t i = 0; i < cnt; ++i) {
            uint32_t v = p[i];
            if (run == 0) { cur = v; run = 1; }
            else if (cur == v) {
                if (++run > 254) { emit_run(so + blen, cur, 255, bm); blen += 5; run = 0; }
            } else {
                emit_run(so + blen, cur, run, bm);
                blen += rle_enc_small(run);
                cur = v; run = 1;
            }
        }
        if (last_lane && run > 0) { emit_run(so + blen, cur, run, bm); blen += rle_enc_small(run); }
    }
    // end of the tile's output = where the last lane with bytes stopped
    const unsigned have = __ballot_sync(FULL, cnt > 0);
    const uint32_t end = __shfl_sync(FULL, blen, 31 - __clz((int)have));
    __syncwarp();
    uint8_t *out = rle + (uint64_t)bl * stride;
    const uint8_t *sb = so_all[warp_id()];
    const uint32_t a0 = base & ~15u;                     // first 16-byte chunk of the block's output this tile touches
    for (uint32_t c = a0 + lane * 16; c < end; c += 32 * 16) {
        const uint8_t *src = sb + (c - a0);
        if (c >= base && c + 16 <= end) *reinterpret_cast<uint4 *>(out + c) = *reinterpret_cast<const uint4 *>(src);
        else
            for (uint32_t x = max(c, base); x < min(c + 16, end); ++x) out[x] = src[x - c];
    }
}

// ---------------------------------------------------------------- CRC32 (CRC32.hpp:30-92): MSB-first, poly 0x04C11DB7
constexpr int CRC_THREADS = 256;
constexpr int CRC_BYTES_PER_THREAD = 64;
constexpr int CRC_CHUNK = CRC_THREADS * CRC_BYTES_PER_THREAD; // 16 KiB

__constant__ uint32_t c_crc_table[256];      // CRC32.hpp:33-65
__constant__ uint32_t c_crc_pow8[48];        // x^(8 * 2^j) mod P
__constant__ uint32_t c_crc_lane_shift[256]; // x^(8 * 64 * (255 - t)) mod P

__device__ __forceinline__ uint32_t crc_xpow8(uint64_t k)
{
    return crc_xpow8_t(k, [](int j) { return c_crc_pow8[j]; });
}

// Chunks are aligned to the END of each block so every thread's shift is a compile-time-known multiple of 64 bytes.
// crc_acc[bl] accumulates (XOR) the "pure" CRC (register started at 0); k_crc_finish folds in the 0xFFFFFFFF start value.
__global__ void __launch_bounds__(CRC_THREADS) k_crc_chunks(const uint8_t *__restrict__ in, const uint64_t *__restrict__ starts, uint32_t b0,
                                                             uint32_t *__restrict__ crc_acc)
{
    __shared__ uint32_t tab[256];
    __shared__ uint32_t red[CRC_THREADS / 32];
    tab[threadIdx.x] = c_crc_table[threadIdx.x];
    __syncthreads();
    uint32_t bl = blockIdx.y;
    uint64_t s = starts[b0 + bl], e = starts[b0 + bl + 1];
    uint64_t c = blockIdx.x;
    if (c * CRC_CHUNK >= e - s) return;
    uint64_t chunk_end = e - c * CRC_CHUNK;
    // thread t covers [chunk_end - 64*(256-t), chunk_end - 64*(255-t)) clipped to >= s
    int64_t hi = (int64_t)chunk_end - (int64_t)CRC_BYTES_PER_THREAD * (255 - (int)threadIdx.x);
    int64_t lo = hi - CRC_BYTES_PER_THREAD;
    if (lo < (int64_t)s) lo = (int64_t)s;
    uint32_t crc = 0;
    for (int64_t q = lo; q < hi; ++q) crc = (crc << 8) ^ tab[((crc >> 24) ^ __ldg(in + q)) & 0xff];
    if (crc) crc = crc_mulmod(crc, c_crc_lane_shift[threadIdx.x]);
    crc = __reduce_xor_sync(FULL, crc);
    if (lane_id() == 0) red[warp_id()] = crc;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t x = 0;
        for (int w = 0; w < CRC_THREADS / 32; ++w) x ^= red[w];
        if (x) {
            x = crc_mulmod(x, crc_xpow8(c * CRC_CHUNK));
            atomicXor(&crc_acc[bl], x);
        }
    }
}

// crc[bl] = ~(pure ^ 0xFFFFFFFF * x^(8 len))   (CRC32.hpp:67,70-73)
__global__ void k_crc_finish(const uint64_t *__restrict__ starts, uint32_t b0, uint32_t nb, uint32_t *__restrict__ crc_acc)
{
    uint32_t bl = blockIdx.x * blockDim.x + threadIdx.x;
    if (bl >= nb) return;
    uint64_t len = starts[b0 + bl + 1] - starts[b0 + bl];
    crc_acc[bl] = ~(crc_acc[bl] ^ crc_mulmod(0xffffffffu, crc_xpow8(len)));
}

} // namespace bz
