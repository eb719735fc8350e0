// bwt.cuh -- stage B: Burrows-Wheeler transform of every block's cyclic rotations (reference: DivSufSortBWT,
// /root/reference/kernel.cpp:2429-2456 and everything it calls, :61-2427 -- a sequential divsufsort per work-item).
//
// B200 formulation: all blocks of a batch are sorted together, segment by segment (a block = a segment of `stride`
// element slots), by prefix doubling:
//   round 0   key = first KEY0_BYTES bytes of the rotation; segmented LSD radix sort (8-bit digits)
//   regroup   equal keys form groups; rank[i] = first sorted position of i's group; singletons are final
//   round h   only rotations in groups of size > 1 stay "active": key = (group << nbits) | rank[(i+h) mod n],
//             segmented LSD radix sort of the active list, scatter back into the group's positions, regroup.
//   stop      when no block has active rotations, or 2h >= n (non-primitive block: ties are final).
//   finish    L[j] = T[(SA[j]-1) mod n]; origPtr = rank[0] = first position of rotation 0's tie group.
// Sort elements are single 64-bit words: (sort key << 20) | rotation index (block positions fit 20 bits), so a pass moves
// 8 bytes per element and needs no value array.  The LSD radix pass is hist -> per-block scan -> stable scatter (warp match_any
// multi-split, shared-memory staged so that the global writes of one digit are contiguous).
#pragma once
#include "util.cuh"
#include "packed_math.h"

namespace bz {

constexpr int RS_THREADS = 256;               // scatter: 256 threads x 16 elements
constexpr int RS_ITEMS = 16;
constexpr int RH_THREADS = 512;               // histogram of the same 4096-element tile: 512 threads x 8 elements
constexpr int RH_ITEMS = 8;
constexpr int RH_WARPS = RH_THREADS / 32;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS; // 4096 elements per radix tile
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_WARP_ELEMS = 32 * RS_ITEMS;

constexpr int VAL_BITS = 20;                 // low bits of a sort element: the rotation's start index inside its block
constexpr uint64_t VAL_MASK = (1ull << VAL_BITS) - 1ull;

constexpr int RG_THREADS = 256;
constexpr int RG_ITEMS = 8;
constexpr int RG_TILE = RG_THREADS * RG_ITEMS; // 2048 elements per regroup tile

// ------------------------------------------------------------------------------------------------ radix: histogram
// Each thread owns RS_ITEMS CONSECUTIVE keys and run-length encodes their digits in registers before touching the shared
// histogram: on the (already ordered) group half of a composite key that is one atomic per thread instead of sixteen, and the
// unordered half has few same-address collisions inside a warp anyway.
template <typename K>
__global__ void __launch_bounds__(RH_THREADS) k_radix_hist(const K *__restrict__ keys, const uint32_t *__restrict__ counts, const uint32_t *__restrict__ offs,
                                                            uint64_t stride, int shift, uint32_t mask, uint32_t *__restrict__ hist, int tiles_per_block)
{
    __shared__ uint32_t h[RH_WARPS][256];
    const uint32_t b = blockIdx.y, tile = blockIdx.x;
    const uint32_t cnt = counts[b];
    const uint32_t base = tile * RS_TILE;
    if (base >= cnt) return;
    for (int i = threadIdx.x; i < RH_WARPS * 256; i += RH_THREADS) (&h[0][0])[i] = 0;
    __syncthreads();
    const K *kp = keys + (uint64_t)b * stride + (offs ? offs[b] : 0u) + base;
    const uint32_t tcnt = min((uint32_t)RS_TILE, cnt - base);
    const uint32_t i0 = threadIdx.x * RH_ITEMS;
    uint32_t *wh = h[warp_id()];
    if (i0 < tcnt) {
        constexpr int VEC = 16 / sizeof(K); // keys per 128-bit load
        uint32_t prev = 0xffffffffu, run = 0;
        if (i0 + RH_ITEMS <= tcnt && (reinterpret_cast<uintptr_t>(kp) & 15u) == 0) {
            const uint4 *vp = reinterpret_cast<const uint4 *>(kp + i0);
#pragma unroll
            for (int v = 0; v < RH_ITEMS / VEC; ++v) {
                uint4 q = __ldg(vp + v);
                K kk[VEC];
                if (sizeof(K) == 8) { kk[0] = (K)(((uint64_t)q.y << 32) | q.x); kk[1 % VEC] = (K)(((uint64_t)q.w << 32) | q.z); }
                else { kk[0] = (K)q.x; kk[1 % VEC] = (K)q.y; kk[2 % VEC] = (K)q.z; kk[3 % VEC] = (K)q.w; }
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    uint32_t d = (uint32_t)(kk[j] >> shift) & mask;
                    if (d != prev) { if (run) atomicAdd(&wh[prev], run); prev = d; run = 0; }
                    ++run;
                }
            }
        } else {
            const uint32_t iend = min(i0 + (uint32_t)RH_ITEMS, tcnt);
            for (uint32_t i = i0; i < iend; ++i) {
                uint32_t d = (uint32_t)(kp[i] >> shift) & mask;
                if (d != prev) { if (run) atomicAdd(&wh[prev], run); prev = d; run = 0; }
                ++run;
            }
        }
        if (run) atomicAdd(&wh[prev], run);
    }
    __syncthreads();
    if (threadIdx.x < 256) {
        uint32_t sum = 0;
#pragma unroll
        for (int w = 0; w < RH_WARPS; ++w) sum += h[w][threadIdx.x];
        hist[((uint64_t)b * tiles_per_block + tile) * 256 + threadIdx.x] = sum;
    }
}

// ------------------------------------------------------------------------------------------------ radix: per-block scan
// hist[b][tile][d] -> exclusive prefix over tiles (per digit); digit_base[b][d] = exclusive prefix over digits of the totals.
__global__ void __launch_bounds__(256) k_radix_scan(uint32_t *__restrict__ hist, const uint32_t *__restrict__ counts, int tiles_per_block,
                                                     uint32_t *__restrict__ digit_base)
{
    __shared__ uint32_t scratch[33];
    const uint32_t b = blockIdx.x, d = threadIdx.x;
    const uint32_t cnt = counts[b];
    const uint32_t nt = (cnt + RS_TILE - 1) / RS_TILE;
    uint32_t *hp = hist + (uint64_t)b * tiles_per_block * 256 + d;
    uint32_t run = 0;
    for (uint32_t t = 0; t < nt; ++t) {
        uint32_t v = hp[(uint64_t)t * 256];
        hp[(uint64_t)t * 256] = run;
        run += v;
    }
    uint32_t ex = block_excl_sum(run, scratch, nullptr);
    digit_base[b * 256 + d] = ex;
}

// ------------------------------------------------------------------------------------------------ radix: stable scatter
template <int DB>
struct ScatterSmemT {
    uint32_t warp_hist[RS_WARPS][1 << DB];
    uint32_t tile_base[1 << DB];
    uint32_t gofs[1 << DB];
    uint32_t scratch[34];
    alignas(16) uint64_t keys[RS_TILE];
};
using ScatterSmem = ScatterSmemT<8>;

__global__ void __launch_bounds__(RS_THREADS) k_radix_scatter(const uint64_t *__restrict__ keys_in, uint64_t *__restrict__ keys_out,
                                                               const uint32_t *__restrict__ counts, const uint32_t *__restrict__ offs, uint64_t stride,
                                                               int shift, uint32_t mask,
                                                               const uint32_t *__restrict__ hist, int tiles_per_block,
                                                               const uint32_t *__restrict__ digit_base)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ScatterSmem &sm = *reinterpret_cast<ScatterSmem *>(smem_raw);
    const uint32_t b = blockIdx.y, tile = blockIdx.x;
    const uint32_t cnt = counts[b];
    const uint32_t base = tile * RS_TILE;
    if (base >= cnt) return;
    const uint32_t tcnt = min((uint32_t)RS_TILE, cnt - base);
    const int lane = lane_id(), w = warp_id();
    const unsigned lt = (1u << lane) - 1u;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&sm.warp_hist[0][0])[i] = 0;
    __syncthreads();

    const uint32_t off = offs ? offs[b] : 0u;
    const uint64_t *kp = keys_in + (uint64_t)b * stride + off + base;
    uint64_t key[RS_ITEMS];
    uint16_t rnk[RS_ITEMS];
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        uint32_t li = w * RS_WARP_ELEMS + i * 32 + lane;
        key[i] = li < tcnt ? kp[li] : 0ull;
    }
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        uint32_t li = w * RS_WARP_ELEMS + i * 32 + lane;
        bool valid = li < tcnt;
        uint32_t d = valid ? ((uint32_t)(key[i] >> shift) & mask) : 0xffffffffu;
        unsigned peers = __match_any_sync(FULL, d);
        int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (valid && lane == leader) {
            old = sm.warp_hist[w][d];
            sm.warp_hist[w][d] = old + __popc(peers);
        }
        old = __shfl_sync(FULL, old, leader);
        rnk[i] = (uint16_t)(old + __popc(peers & lt));
        __syncwarp();
    }
    __syncthreads();
    {
        const uint32_t d = threadIdx.x;
        uint32_t sum = 0;
        if (d < 256) {
#pragma unroll
            for (int ww = 0; ww < RS_WARPS; ++ww) {
                uint32_t t = sm.warp_hist[ww][d];
                sm.warp_hist[ww][d] = sum;
                sum += t;
            }
        }
        uint32_t tb = block_excl_sum(sum, sm.scratch, nullptr); // threads >= 256 add 0 behind the digits
        if (d < 256) {
            sm.tile_base[d] = tb;
            sm.gofs[d] = digit_base[b * 256 + d] + hist[((uint64_t)b * tiles_per_block + tile) * 256 + d] - tb;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        uint32_t li = w * RS_WARP_ELEMS + i * 32 + lane;
        if (li < tcnt) {
            uint32_t d = (uint32_t)(key[i] >> shift) & mask;
            uint32_t p = sm.tile_base[d] + sm.warp_hist[w][d] + rnk[i];
            sm.keys[p] = key[i];
        }
    }
    __syncthreads();
    uint64_t *ko = keys_out + (uint64_t)b * stride + off;
    for (uint32_t p = threadIdx.x; p < tcnt; p += RS_THREADS) {
        uint64_t k = sm.keys[p];
        uint32_t d = (uint32_t)(k >> shift) & mask;
        ko[sm.gofs[d] + p] = k;
    }
}

// ------------------------------------------------------------------------------------------------ radix: one sweep per digit
// The same stable LSD pass without the per-tile histogram kernel and the per-block scan: the digit totals of ALL passes come
// from one counting kernel up front (k_radix_count + k_radix_bases), and each pass resolves "how many keys with my digit sit
// in earlier tiles of my block" with a decoupled look-back over per-tile status words (tile aggregate / inclusive prefix), as
// in single-pass prefix scans.  A pass then reads and writes every key exactly once.
#ifndef BZ_OS_MINB
#define BZ_OS_MINB 3
#endif
#ifndef BZ_OS_MBATCH
#define BZ_OS_MBATCH 1 // keys whose MATCH.ANY are issued back to back in a sweep before their counters are updated; 1, 4, 8, 16 measured equal (30.4 / 30.0 / 31.1 / 29.9 ms per GiB): the ADU pipe is throughput-bound, not latency-bound
#endif
constexpr int RC_TILES = 16;                 // radix tiles per counting CTA
constexpr int RC_MAXPASS = 5;
constexpr uint32_t LB_AGG = 1u << 30, LB_PREFIX = 2u << 30, LB_VALUE = (1u << 30) - 1u;

// partial[(b * ctas_per_block + cta) * npass * 256 + p * 256 + d] = keys of the CTA's tiles whose pass-p digit is d (8-bit digits
// at VAL_BITS + 8p)
__global__ void __launch_bounds__(RH_THREADS) k_radix_count(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ counts, uint64_t stride, int npass,
                                                             uint32_t *__restrict__ partial, int ctas_per_block)
{
    __shared__ uint32_t h[RC_MAXPASS][256];
    const uint32_t b = blockIdx.y;
    const uint32_t cnt = counts[b];
    const uint32_t base0 = blockIdx.x * (RC_TILES * RS_TILE);
    if (base0 >= cnt) return;
    for (int i = threadIdx.x; i < RC_MAXPASS * 256; i += RH_THREADS) (&h[0][0])[i] = 0;
    __syncthreads();
    const uint64_t *kp = keys + (uint64_t)b * stride;
    const uint32_t end = min(cnt, base0 + RC_TILES * RS_TILE);
    for (uint32_t base = base0; base < end; base += RS_TILE) {
        const uint32_t i0 = base + threadIdx.x * RH_ITEMS;
        if (i0 >= end) continue;
        uint64_t kk[RH_ITEMS];
        if (i0 + RH_ITEMS <= end) {
            const uint4 *vp = reinterpret_cast<const uint4 *>(kp + i0);
#pragma unroll
            for (int v = 0; v < RH_ITEMS / 2; ++v) {
                const uint4 q = __ldg(vp + v);
                kk[2 * v] = ((uint64_t)q.y << 32) | q.x;
                kk[2 * v + 1] = ((uint64_t)q.w << 32) | q.z;
            }
        } else {
#pragma unroll
            for (int j = 0; j < RH_ITEMS; ++j) kk[j] = i0 + j < end ? kp[i0 + j] : ~0ull;
        }
        const int nvalid = (int)min((uint32_t)RH_ITEMS, end - i0);
        for (int p = 0; p < npass; ++p) { // consecutive keys: run-length encode the digit before touching shared memory
            const int sh = VAL_BITS + 8 * p;
            uint32_t prev = 0xffffffffu, run = 0;
#pragma unroll
            for (int j = 0; j < RH_ITEMS; ++j) {
                if (j < nvalid) {
                    const uint32_t d = (uint32_t)(kk[j] >> sh) & 0xffu;
                    if (d != prev) { if (run) atomicAdd(&h[p][prev], run); prev = d; run = 0; }
                    ++run;
                }
            }
            if (run) atomicAdd(&h[p][prev], run);
        }
    }
    __syncthreads();
    uint32_t *out = partial + ((uint64_t)b * ctas_per_block + blockIdx.x) * npass * 256;
    for (int i = threadIdx.x; i < npass * 256; i += RH_THREADS) out[i] = (&h[0][0])[i];
}

// digit_base[(p * n_blocks + b) * 256 + d] = keys of block b whose pass-p digit is < d
// (nrep > 1: the totals of pass 0 are valid for nrep passes -- round 0, where every pass's digit is a byte of the same cyclic text)
// present != nullptr: also the block's symbol-presence bitmap (value v <-> bit v & 31 of word v >> 5) from the pass-0 totals -- for
// round 0 those are the byte counts of the RLE1 block, which is exactly blockValuesPresent (BlockCompressor.hpp:136,148).
__global__ void __launch_bounds__(256) k_radix_bases(const uint32_t *__restrict__ partial, const uint32_t *__restrict__ counts, int npass, int ctas_per_block,
                                                      uint32_t n_blocks, uint32_t *__restrict__ digit_base, int nrep, uint32_t *__restrict__ present)
{
    __shared__ uint32_t scratch[33];
    const uint32_t b = blockIdx.x, d = threadIdx.x;
    const uint32_t cnt = counts[b];
    const uint32_t nc = (cnt + RC_TILES * RS_TILE - 1) / (RC_TILES * RS_TILE);
    for (int p = 0; p < npass; ++p) {
        uint32_t tot = 0;
        for (uint32_t c = 0; c < nc; ++c) tot += partial[(((uint64_t)b * ctas_per_block + c) * npass + p) * 256 + d];
        if (present && p == 0) {
            const unsigned m = __ballot_sync(FULL, tot != 0);
            if (lane_id() == 0) present[b * 8 + warp_id()] = m;
        }
        const uint32_t ex = block_excl_sum(tot, scratch, nullptr);
        for (int r = 0; r < (npass == 1 ? nrep : 1); ++r) digit_base[((uint64_t)(p + r) * n_blocks + b) * 256 + d] = ex;
    }
}
// the same for digits of DB bits (1 << DB bins per pass, (1 << DB) / 256 consecutive bins per thread)
template <int DB>
__global__ void __launch_bounds__(256) k_radix_bases_wide(const uint32_t *__restrict__ partial, const uint32_t *__restrict__ counts, int npass, int ctas_per_block,
                                                           uint32_t n_blocks, uint32_t *__restrict__ digit_base)
{
    constexpr int NB = 1 << DB, DPT = NB / 256;
    __shared__ uint32_t scratch[33];
    const uint32_t b = blockIdx.x;
    const uint32_t cnt = counts[b];
    const uint32_t nc = (cnt + RC_TILES * RS_TILE - 1) / (RC_TILES * RS_TILE);
    for (int p = 0; p < npass; ++p) {
        uint32_t tot[DPT], mine = 0;
#pragma unroll
        for (int q = 0; q < DPT; ++q) {
            const uint32_t d = threadIdx.x * DPT + q;
            uint32_t t = 0;
            for (uint32_t c = 0; c < nc; ++c) t += partial[(((uint64_t)b * ctas_per_block + c) * npass + p) * NB + d];
            tot[q] = t;
            mine += t;
        }
        uint32_t ex = block_excl_sum(mine, scratch, nullptr);
#pragma unroll
        for (int q = 0; q < DPT; ++q) {
            digit_base[((uint64_t)p * n_blocks + b) * NB + threadIdx.x * DPT + q] = ex;
            ex += tot[q];
        }
    }
}

// byte histogram of the RLE1 block (same partial[] layout as k_radix_count with npass = 1)
__global__ void __launch_bounds__(256) k_bwt_bytehist(const uint8_t *__restrict__ rle, const uint32_t *__restrict__ n_arr, uint64_t stride,
                                                       uint32_t *__restrict__ partial, int ctas_per_block)
{
    __shared__ uint32_t h[8][256];
    const uint32_t b = blockIdx.y;
    const uint32_t n = n_arr[b];
    const uint32_t base0 = blockIdx.x * (RC_TILES * RS_TILE);
    if (base0 >= n) return;
    for (int i = threadIdx.x; i < 8 * 256; i += 256) (&h[0][0])[i] = 0;
    __syncthreads();
    const uint8_t *T = rle + (uint64_t)b * stride;
    const uint32_t end = min(n, base0 + RC_TILES * RS_TILE);
    uint32_t *wh = h[warp_id()];
    for (uint32_t i = base0 + threadIdx.x * 16; i < end; i += 256 * 16) {
        const uint4 q = __ldg(reinterpret_cast<const uint4 *>(T + i)); // the slot is padded past n
        const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (i + j < end) atomicAdd(&wh[(w4[j >> 2] >> (8 * (j & 3))) & 0xffu], 1u);
    }
    __syncthreads();
    uint32_t s = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += h[w][threadIdx.x];
    partial[((uint64_t)b * ctas_per_block + blockIdx.x) * 256 + threadIdx.x] = s;
}

// Packed round-0 keys (see k_bwt_count_packed): thread v < 256 fills map[v] = rank of byte value v among the block's used values;
// bits per symbol, symbols per key and their total width come back in every thread.
__device__ __forceinline__ PackedShape packed_setup(const uint32_t *__restrict__ present, uint32_t b, uint8_t *map)
{
    const uint32_t v = threadIdx.x;
    uint32_t alpha = 0, below = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const uint32_t m = present[b * 8 + w];
        alpha += __popc(m);
        if ((uint32_t)w < (v >> 5)) below += __popc(m);
        else if ((uint32_t)w == (v >> 5)) below += __popc(m & ((1u << (v & 31u)) - 1u));
    }
    if (v < 256) map[v] = (uint8_t)below;
    return packed_shape(alpha);
}
// Mapped symbols of positions [base, base + need) of the block (wrapping) into sym[], 16 per thread, all threads of the CTA.
// Layout: symbol p lives at byte p + 4 * (p >> 4) (one pad word after every 16 symbols), so that threads walking 16 consecutive
// symbols each (a 20-byte stride between lanes) never meet in a shared-memory bank.
constexpr int PK_STAGE_BYTES = (4096 + 32) / 16 * 20;
__device__ __forceinline__ uint32_t pk_addr(uint32_t p) { return p + ((p >> 4) << 2); }
__device__ __forceinline__ void packed_stage(const uint8_t *__restrict__ T, uint32_t n, uint32_t base, uint32_t need, const uint8_t *map, uint8_t *sym)
{
    for (uint32_t j = threadIdx.x * 16; j < need; j += blockDim.x * 16) {
        if (base + j + 16 <= n) {
            const uint4 q = __ldg(reinterpret_cast<const uint4 *>(T + base + j));
            const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
            uint32_t *dst = reinterpret_cast<uint32_t *>(sym + pk_addr(j));
#pragma unroll
            for (int k = 0; k < 4; ++k)
                dst[k] = (uint32_t)map[w4[k] & 0xffu] | ((uint32_t)map[(w4[k] >> 8) & 0xffu] << 8) | ((uint32_t)map[(w4[k] >> 16) & 0xffu] << 16) |
                         ((uint32_t)map[w4[k] >> 24] << 24);
        } else {
            for (uint32_t k = j; k < min(j + 16, need); ++k) { uint32_t p = base + k; sym[pk_addr(k)] = map[T[p >= n ? (p - n) % n : p]]; }
        }
    }
}
// the packed keys of the 16 consecutive rotations starting at tile position p0 (a multiple of 16): a sliding window of hc symbols
template <typename F>
__device__ __forceinline__ void packed_keys16(const uint8_t *sym, uint32_t p0, const PackedShape &ps, F &&emit)
{
    uint64_t c = 0;
    for (int j = 0; j + 1 < ps.hc; ++j) c = (c << ps.bits) | sym[pk_addr(p0 + (uint32_t)j)];
    const uint32_t tail = p0 + (uint32_t)ps.hc - 1u;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        c = packed_push(c, sym[pk_addr(tail + (uint32_t)i)], ps);
        emit(i, packed_key_of(c, ps));
    }
}

// FROM_TEXT: the first pass of round 0 builds its elements on the fly, (first text_bytes bytes of rotation i, big-endian) << 20 | i,
// from the block `text` instead of reading them from keys_in.
template <bool FROM_TEXT, int DB = 8>
__global__ void __launch_bounds__(RS_THREADS, BZ_OS_MINB) k_radix_onesweep(const uint64_t *__restrict__ keys_in, uint64_t *__restrict__ keys_out,
                                                                const uint32_t *__restrict__ counts, uint64_t stride, int shift,
                                                                const uint32_t *__restrict__ digit_base, uint32_t *__restrict__ tile_state,
                                                                int tiles_per_block, uint32_t *__restrict__ ticket, const uint8_t *__restrict__ text,
                                                                int text_bytes, const uint32_t *__restrict__ present)
{
    constexpr uint32_t NB = 1u << DB, DMASK = NB - 1u; // bins of one pass
    constexpr int DPT = (int)NB / RS_THREADS;          // consecutive bins per thread in the scan
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ScatterSmemT<DB> &sm = *reinterpret_cast<ScatterSmemT<DB> *>(smem_raw);
    __shared__ uint8_t s_map[256]; // text_bytes == 0: keys over the block's own alphabet
    // tiles are taken in ticket order, so the tiles a CTA looks back at are always held by CTAs that are already running
    if (threadIdx.x == 0) sm.scratch[33] = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tk = sm.scratch[33];
    // tile-major over the batch: the previous tile of the same block is gridDim.y tickets back, i.e. normally finished already,
    // so the look-back below is one step deep instead of walking through every tile that is still in flight
    const uint32_t b = tk % gridDim.y, tile = tk / gridDim.y;
    const uint32_t cnt = counts[b];
    const uint32_t base = tile * RS_TILE;
    if (base >= cnt) return;
    const uint32_t tcnt = min((uint32_t)RS_TILE, cnt - base);
    const int lane = lane_id(), w = warp_id();
    const unsigned lt = (1u << lane) - 1u;
    for (int i = threadIdx.x; i < RS_WARPS * (int)NB; i += RS_THREADS) (&sm.warp_hist[0][0])[i] = 0;
    uint64_t key[RS_ITEMS];
    uint16_t rnk[RS_ITEMS];
    if (FROM_TEXT && text_bytes == 0) {
        const PackedShape ps = packed_setup(present, b, s_map);
        __syncthreads();
        uint8_t *tb = reinterpret_cast<uint8_t *>(sm.keys);
        packed_stage(text + (uint64_t)b * stride, cnt, base, tcnt + (uint32_t)ps.hc - 1u, s_map, tb);
        __syncthreads();
        // (this sweep ranks with atomics, so which thread holds which element does not matter: 16 consecutive rotations per thread)
        const uint32_t p0 = threadIdx.x * RS_ITEMS;
        packed_keys16(tb, p0, ps, [&](int i, uint64_t k) { key[i] = p0 + i < tcnt ? ((k << VAL_BITS) | (base + p0 + i)) : 0ull; });
    } else if (FROM_TEXT) {
        // stage the tile's bytes (+ text_bytes - 1 for the last key, wrapping around the block) in the output staging area
        uint8_t *tb = reinterpret_cast<uint8_t *>(sm.keys);
        const uint8_t *T = text + (uint64_t)b * stride;
        const uint32_t extra = (uint32_t)text_bytes - 1u;
        for (uint32_t j = threadIdx.x * 16; j < tcnt + extra; j += RS_THREADS * 16) {
            if (base + j + 16 <= cnt) *reinterpret_cast<uint4 *>(tb + j) = __ldg(reinterpret_cast<const uint4 *>(T + base + j));
            else
                for (uint32_t k = j; k < min(j + 16, tcnt + extra); ++k) { uint32_t p = base + k; tb[k] = T[p >= cnt ? (p - cnt) % cnt : p]; }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            const uint32_t li = w * RS_WARP_ELEMS + i * 32 + lane;
            // (text_bytes == 1: the key is the byte itself -- the decompressor's counting sort of the BWT column, dec.cuh)
            uint64_t k4 = text_bytes == 1 ? (uint64_t)tb[li]
                                          : (uint64_t)(((uint32_t)tb[li] << 24) | ((uint32_t)tb[li + 1] << 16) | ((uint32_t)tb[li + 2] << 8) | (uint32_t)tb[li + 3]);
            if (text_bytes == 5) k4 = (k4 << 8) | tb[li + 4]; // 40 key bits + 20 index bits still fit the element
            key[i] = li < tcnt ? ((k4 << VAL_BITS) | (base + li)) : 0ull;
        }
    } else {
        const uint64_t *kp = keys_in + (uint64_t)b * stride + base;
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            const uint32_t li = w * RS_WARP_ELEMS + i * 32 + lane;
            key[i] = li < tcnt ? kp[li] : 0ull;
        }
    }
    __syncthreads();
    // Round 0's first pass has no earlier order to preserve: inside the tile its keys take their slots with one shared-memory
    // atomic each instead of the stable warp multi-split (the decompressor's 1-byte counting sort must stay stable).
    // (Not used for the bag sorts: their low digits can be nearly constant, and 32 lanes on one counter serialise.)
    const bool unstable = FROM_TEXT && text_bytes != 1;
    const bool contig = FROM_TEXT && text_bytes == 0; // packed keys: thread t holds tile positions 16 t .. 16 t + 15
    if (unstable) {
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            const uint32_t li = contig ? threadIdx.x * RS_ITEMS + i : w * RS_WARP_ELEMS + i * 32 + lane;
            if (li < tcnt) rnk[i] = (uint16_t)atomicAdd(&sm.warp_hist[0][(uint32_t)(key[i] >> shift) & DMASK], 1u);
        }
    } else {
        // the matches of BZ_OS_MBATCH keys are issued back to back (their latency overlaps), then the counters are updated in key order
#pragma unroll
        for (int h = 0; h < RS_ITEMS; h += BZ_OS_MBATCH) {
            unsigned peers[BZ_OS_MBATCH];
#pragma unroll
            for (int j = 0; j < BZ_OS_MBATCH; ++j) {
                const uint32_t li = w * RS_WARP_ELEMS + (h + j) * 32 + lane;
                const uint32_t d = li < tcnt ? ((uint32_t)(key[h + j] >> shift) & DMASK) : 0xffffffffu;
                peers[j] = __match_any_sync(FULL, d);
            }
#pragma unroll
            for (int j = 0; j < BZ_OS_MBATCH; ++j) {
                const uint32_t li = w * RS_WARP_ELEMS + (h + j) * 32 + lane;
                const bool valid = li < tcnt;
                const uint32_t d = (uint32_t)(key[h + j] >> shift) & DMASK;
                const int leader = __ffs(peers[j]) - 1;
                uint32_t old = 0;
                if (valid && lane == leader) {
                    old = sm.warp_hist[w][d];
                    sm.warp_hist[w][d] = old + __popc(peers[j]);
                }
                old = __shfl_sync(FULL, old, leader);
                rnk[h + j] = (uint16_t)(old + __popc(peers[j] & lt));
                __syncwarp();
            }
        }
    }
    __syncthreads();
    {
        // every thread owns DPT consecutive digits (RS_THREADS == 256; DPT = 1 for 8-bit digits)
        uint32_t sums[DPT], mine = 0;
        volatile uint32_t *st = tile_state + ((uint64_t)b * tiles_per_block) * NB;
#pragma unroll
        for (int q = 0; q < DPT; ++q) {
            const uint32_t d = threadIdx.x * DPT + q;
            uint32_t sum = 0;
#pragma unroll
            for (int ww = 0; ww < RS_WARPS; ++ww) {
                const uint32_t t = sm.warp_hist[ww][d];
                sm.warp_hist[ww][d] = sum;
                sum += t;
            }
            sums[q] = sum;
            mine += sum;
            // publish this tile's count, then add up the tiles in front of it
            st[(uint64_t)tile * NB + d] = (tile == 0 ? LB_PREFIX : LB_AGG) | sum;
        }
        uint32_t tb = block_excl_sum(mine, sm.scratch, nullptr);
#pragma unroll
        for (int q = 0; q < DPT; ++q) {
            const uint32_t d = threadIdx.x * DPT + q;
            uint32_t before = 0;
            if (tile > 0) {
                for (int t = (int)tile - 1;; --t) {
                    uint32_t v;
                    do { v = st[(uint64_t)t * NB + d]; } while ((v & ~LB_VALUE) == 0);
                    before += v & LB_VALUE;
                    if (v & LB_PREFIX) break;
                }
                st[(uint64_t)tile * NB + d] = LB_PREFIX | (before + sums[q]);
            }
            sm.tile_base[d] = tb;
            sm.gofs[d] = digit_base[b * NB + d] + before - tb;
            tb += sums[q];
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const uint32_t li = contig ? threadIdx.x * RS_ITEMS + i : w * RS_WARP_ELEMS + i * 32 + lane;
        if (li < tcnt) {
            const uint32_t d = (uint32_t)(key[i] >> shift) & DMASK;
            sm.keys[sm.tile_base[d] + (unstable ? 0u : sm.warp_hist[w][d]) + rnk[i]] = key[i];
        }
    }
    __syncthreads();
    uint64_t *ko = keys_out + (uint64_t)b * stride;
    for (uint32_t p = threadIdx.x; p < tcnt; p += RS_THREADS) {
        const uint64_t k = sm.keys[p];
        ko[sm.gofs[(uint32_t)(k >> shift) & DMASK] + p] = k;
    }
}

// ------------------------------------------------------------------------------------------------ round 0 keys
// element = (first `nbytes` (<= 4) bytes of rotation i, big-endian) << 20 | i
__global__ void __launch_bounds__(256) k_bwt_keys0(const uint8_t *__restrict__ rle, const uint32_t *__restrict__ n_arr, uint64_t stride, int nbytes,
                                                    uint64_t *__restrict__ keys)
{
    const uint32_t b = blockIdx.y;
    const uint32_t n = n_arr[b];
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t *T = rle + (uint64_t)b * stride;
    uint64_t key = 0;
    uint32_t p = i;
    for (int k = 0; k < nbytes; ++k) { // nbytes <= 5: 40 key bits above the 20 index bits
        key = (key << 8) | T[p];
        if (++p == n) p = 0;
    }
    keys[(uint64_t)b * stride + i] = (key << VAL_BITS) | i;
}

// Round-0 keys over the block's own alphabet.  The block uses `alpha` of the 256 byte values (the presence bitmap); mapped to
// 0..alpha-1 in byte order a symbol needs bits = ceil(log2 alpha) bits, and the 40 key bits of an element hold the first
// floor(40 / bits) symbols of the rotation (capped at 16) plus the leading bits of the next one -- 8 symbols for lower-case text,
// 6 for program text, 5 for binary data.  Mapping and truncation are monotone, so the order of the keys is the order of the
// rotations on h0 = floor(40 / bits) bytes, refined further by the partial symbol (a refinement consistent with the final order
// is all prefix doubling needs).  h0 goes to h0_out[b]: every block of a batch doubles from its own depth.
constexpr int KP_TILE = 4096;
// digit histograms of the packed keys for all five passes, straight from the text (the keys themselves are built again by the first
// sweep): partial[] in k_radix_count's layout, one CTA per RC_TILES radix tiles
template <int DB, int NPASS>
__global__ void __launch_bounds__(256) k_bwt_count_packed(const uint8_t *__restrict__ rle, const uint32_t *__restrict__ n_arr, uint64_t stride,
                                                           const uint32_t *__restrict__ present, uint32_t *__restrict__ partial, int ctas_per_block,
                                                           uint32_t *__restrict__ h0_out)
{
    constexpr int NB = 1 << DB;
    __shared__ uint8_t map[256];
    __shared__ __align__(16) uint8_t sym[PK_STAGE_BYTES];
    __shared__ uint32_t h[NPASS][NB];
    const uint32_t b = blockIdx.y, n = n_arr[b];
    const uint32_t base0 = blockIdx.x * (RC_TILES * RS_TILE);
    const PackedShape ps = packed_setup(present, b, map);
    if (blockIdx.x == 0 && threadIdx.x == 0) h0_out[b] = (uint32_t)ps.h0;
    if (base0 >= n) return;
    for (int i = threadIdx.x; i < NPASS * NB; i += 256) (&h[0][0])[i] = 0;
    const uint8_t *T = rle + (uint64_t)b * stride;
    const uint32_t end = min(n, base0 + RC_TILES * RS_TILE);
    for (uint32_t base = base0; base < end; base += KP_TILE) {
        const uint32_t tcnt = min((uint32_t)KP_TILE, end - base);
        __syncthreads();
        packed_stage(T, n, base, tcnt + (uint32_t)ps.hc - 1u, map, sym);
        __syncthreads();
        const uint32_t p0 = threadIdx.x * 16;
        if (p0 < tcnt)
            packed_keys16(sym, p0, ps, [&](int i, uint64_t key) {
                if (p0 + i < tcnt) {
#pragma unroll
                    for (int p = 0; p < NPASS; ++p) atomicAdd(&h[p][(uint32_t)(key >> (DB * p)) & (uint32_t)(NB - 1)], 1u);
                }
            });
    }
    __syncthreads();
    uint32_t *out = partial + ((uint64_t)b * ctas_per_block + blockIdx.x) * NPASS * NB;
    for (int i = threadIdx.x; i < NPASS * NB; i += 256) out[i] = (&h[0][0])[i];
}

__global__ void __launch_bounds__(256) k_fill_u32(uint32_t *__restrict__ p, uint32_t n, uint32_t v)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ------------------------------------------------------------------------------------------------ regroup
// After a sort, equal keys form the new groups.  For every element: rank[rotation] = SA position of its group's first element;
// a rotation alone in its group (or any rotation once 2h >= n) is final and its BWT byte goes straight to L[position].
// Rotations that stay ambiguous are kept for the next round in one of two places:
//   * groups that lie inside one SA tile (RG_TILE consecutive SA positions) move to that tile's LOCAL region: all later
//     refinements of such a group stay inside the tile, so k_bwt_local sorts and regroups them in shared memory;
//   * the rest (groups larger than a tile or straddling a tile edge) go to the block's BAG, an unordered list that the next
//     round sorts with the global radix sort.
// Elements carry the SA position of their group's first element (bag: absolute, local region: relative to the tile), so a
// sorted list needs no position array: position = group position + index inside the (old) group.
struct RoundArgs {
    const uint32_t *n_arr;      // [G] block lengths
    uint64_t stride;
    const uint8_t *T;           // [G*stride] RLE1 blocks
    uint8_t *L;                 // [G*stride] BWT last column
    uint32_t *rank;             // [G*stride]
    uint64_t *loc;              // [G*stride] local regions: tile t of block b owns loc[b*stride + t*RG_TILE ...]
    uint32_t *loc_cnt;          // [G*tiles] elements in each local region
    int tiles_per_block;
    unsigned long long *totals; // [0] rotations kept for the next round (all blocks)
    const uint32_t *h0;         // [G] bytes the round-0 key of each block orders its rotations by (differs per block with packed keys)
    uint32_t round;             // rotations of block b are now ordered by their first h0[b] << round bytes
    int split;                  // 0: keep every group in the bag (debug)
    // Local regions that hold elements, as a list of region ids (block * tiles_per_block + tile): the kernels of a round are launched
    // over the list instead of over every (tile, block) pair -- after the first rounds most regions are empty, and an empty CTA still
    // costs its launch.  work_in: the regions of this round; work_out / work_out_cnt: whoever makes a region non-empty for the next
    // round (its own local sort keeping elements, or the first regroup CTA that appends to an empty one) adds it.
    const uint32_t *work_in;
    uint32_t *work_out;
    uint32_t *work_out_cnt;
};
__device__ __forceinline__ uint32_t sorted_len_of(const RoundArgs &A, uint32_t b) { return A.h0[b] << A.round; } // h0 <= 16, round <= 20

constexpr int GPOS_SHIFT = 2 * VAL_BITS;  // element = group position << 40 | rank of the second half << 20 | rotation index
constexpr int SA_TILE_BITS = 11;
static_assert(RS_THREADS == 256, "one thread per digit in k_radix_onesweep");
static_assert((1 << SA_TILE_BITS) == RG_TILE, "SA tile = regroup tile");
constexpr int INT_INF = 0x7fffffff;

// kp[] is sorted non-decreasing by (e >> shift) and kp[pos] >> shift == key: first / last slot of the run of `key`.
// Called by a whole warp: 32 neighbouring slots in one round trip (a short run ends there), then a 32-ary search.
__device__ __forceinline__ uint32_t warp_run_first(const uint64_t *kp, uint32_t pos, uint64_t key, int shift)
{
    const uint32_t lane = (uint32_t)lane_id();
    {
        const bool eq = pos >= lane + 1 && (kp[pos - 1 - lane] >> shift) == key;
        const unsigned ne = ~__ballot_sync(FULL, eq);
        if (ne) return pos - (uint32_t)(__ffs((int)ne) - 1);
    }
    uint32_t lo = 0, hi = pos - 32; // slot hi holds the key; the answer is in [lo, hi]
    while (lo < hi) {
        const uint32_t chunk = (hi - lo + 31) / 32;
        const uint32_t x = lo + lane * chunk;
        const bool t = x >= hi || (kp[x] >> shift) == key;
        const unsigned tb = __ballot_sync(FULL, t);
        if (tb == 0) { lo += 31 * chunk + 1; continue; }
        const uint32_t f = (uint32_t)(__ffs((int)tb) - 1);
        if (f == 0) return lo;
        hi = min(hi, lo + f * chunk);
        lo += (f - 1) * chunk + 1;
    }
    return hi;
}
__device__ __forceinline__ uint32_t warp_run_last(const uint64_t *kp, uint32_t pos, uint32_t cnt, uint64_t key, int shift)
{
    const uint32_t lane = (uint32_t)lane_id();
    {
        const bool eq = pos + 1 + lane < cnt && (kp[pos + 1 + lane] >> shift) == key;
        const unsigned ne = ~__ballot_sync(FULL, eq);
        if (ne) return pos + (uint32_t)(__ffs((int)ne) - 1);
    }
    uint32_t lo = pos + 32, hi = cnt - 1; // slot lo holds the key; the answer is in [lo, hi]
    while (lo < hi) {
        const uint32_t chunk = (hi - lo + 31) / 32;
        const uint32_t x = hi - lane * chunk; // lo >= 32 keeps this from wrapping
        const bool t = x <= lo || (kp[x] >> shift) == key;
        const unsigned tb = __ballot_sync(FULL, t);
        if (tb == 0) { hi -= 31 * chunk + 1; continue; }
        const uint32_t f = (uint32_t)(__ffs((int)tb) - 1);
        if (f == 0) return hi;
        lo = max(lo, hi - f * chunk);
        hi -= (f - 1) * chunk + 1;
    }
    return lo;
}

// What a thread knows about its ITEMS consecutive slots a0 .. a0+ITEMS-1 of a sorted list after the block-wide scans.
template <int ITEMS>
struct GroupView {
    uint32_t P[ITEMS];    // SA position of the new group's first element (relative to the list's position base)
    uint32_t p[ITEMS];    // SA position of the element
    uint32_t size[ITEMS]; // elements in the new group
};

// e[i] = slot a0+i (0 past the end); prev_e = slot a0-1 (unused when a0 == 0); next_e = slot a0+ITEMS (unused past the end).
// edge_s / edge_o: first slot of the new / old group that contains the tile's first slot; edge_e: last slot of the new group
// that contains the tile's last slot.  ROUND0: the list is the whole block sorted by its first key (one old group at 0).
template <bool ROUND0, int ITEMS>
__device__ __forceinline__ void group_view(const uint64_t (&e)[ITEMS], uint64_t prev_e, uint64_t next_e, uint32_t a0, uint32_t cnt, uint32_t edge_s,
                                           uint32_t edge_o, uint32_t edge_e, int *s_i, GroupView<ITEMS> &v)
{
    bool nf[ITEMS], of[ITEMS], nx[ITEMS];
    int last_n = -1, last_o = -1, first_end = INT_INF;
    if (a0 < cnt) {
        uint64_t prev = prev_e;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t a = a0 + i;
            nf[i] = (a == 0) || ((e[i] >> VAL_BITS) != (prev >> VAL_BITS));
            of[i] = ROUND0 ? (a == 0) : ((a == 0) || ((e[i] >> GPOS_SHIFT) != (prev >> GPOS_SHIFT)));
            prev = e[i];
        }
#pragma unroll
        for (int i = 0; i < ITEMS - 1; ++i) nx[i] = (a0 + i + 1 >= cnt) || nf[i + 1];
        nx[ITEMS - 1] = (a0 + ITEMS >= cnt) || ((next_e >> VAL_BITS) != (e[ITEMS - 1] >> VAL_BITS));
#pragma unroll
        for (int i = 0; i < ITEMS; ++i)
            if (a0 + i < cnt) {
                if (nf[i]) last_n = (int)(a0 + i);
                if (of[i]) last_o = (int)(a0 + i);
            }
#pragma unroll
        for (int i = ITEMS - 1; i >= 0; --i)
            if (a0 + i < cnt && nx[i]) first_end = (int)(a0 + i);
    }
    int prev_n, prev_o, next_end;
    block_scan_max2_minsuffix(last_n, last_o, first_end, s_i, prev_n, prev_o, next_end);
    if (a0 >= cnt) return;
    int e_run = next_end != INT_INF ? next_end : (int)edge_e;
    int en[ITEMS];
#pragma unroll
    for (int i = ITEMS - 1; i >= 0; --i) {
        if (a0 + i < cnt && nx[i]) e_run = (int)(a0 + i);
        en[i] = e_run;
    }
    int s_run = prev_n >= 0 ? prev_n : (int)edge_s;
    int o_run = ROUND0 ? 0 : (prev_o >= 0 ? prev_o : (int)edge_o);
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const uint32_t a = a0 + i;
        if (a < cnt) {
            if (nf[i]) s_run = (int)a;
            if (!ROUND0 && of[i]) o_run = (int)a;
            const uint32_t gpos = ROUND0 ? 0u : (uint32_t)(e[i] >> GPOS_SHIFT);
            v.P[i] = gpos + (uint32_t)(s_run - o_run);
            v.p[i] = gpos + (a - (uint32_t)o_run);
            v.size[i] = (uint32_t)(en[i] - s_run + 1);
        }
    }
}

// One CTA per RG_TILE slots of a block's sorted list (round 0: all rotations by their first key; later: the sorted bag).
template <bool ROUND0>
__global__ void __launch_bounds__(RG_THREADS) k_bwt_regroup(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ counts, RoundArgs A,
                                                             uint64_t *__restrict__ bag, uint32_t *__restrict__ bag_cnt)
{
    __shared__ int s_i[96];
    __shared__ uint32_t s_u[34];
    __shared__ uint32_t s_edge[3];
    __shared__ uint32_t s_base, s_tmain, s_lbase, s_lother;
    const uint32_t b = blockIdx.y, tile = blockIdx.x;
    const uint32_t cnt = counts[b];
    const uint32_t base = tile * RG_TILE;
    if (base >= cnt) return;
    const uint64_t *kp = keys + (uint64_t)b * A.stride;
    const uint32_t tile_last = min(base + (uint32_t)RG_TILE, cnt) - 1;
    if (threadIdx.x == 0) { s_tmain = 0xffffffffu; s_lother = 0; }
    if (warp_id() < 3) { // where do the groups that reach over the tile's edges start / end?
        uint32_t r;
        if (warp_id() == 0) r = warp_run_first(kp, base, kp[base] >> VAL_BITS, VAL_BITS);
        else if (warp_id() == 1) r = ROUND0 ? 0u : warp_run_first(kp, base, kp[base] >> GPOS_SHIFT, GPOS_SHIFT);
        else r = warp_run_last(kp, tile_last, cnt, kp[tile_last] >> VAL_BITS, VAL_BITS);
        if (lane_id() == 0) s_edge[warp_id()] = r;
    }
    const uint32_t a0 = base + threadIdx.x * RG_ITEMS;
    uint64_t e[RG_ITEMS];
    uint64_t prev_e = 0, next_e = 0;
    if (a0 < cnt) {
        if (a0 + RG_ITEMS <= cnt) {
            const uint4 *vp = reinterpret_cast<const uint4 *>(kp + a0);
#pragma unroll
            for (int q = 0; q < RG_ITEMS / 2; ++q) {
                const uint4 x = __ldg(vp + q);
                e[2 * q] = ((uint64_t)x.y << 32) | x.x;
                e[2 * q + 1] = ((uint64_t)x.w << 32) | x.z;
            }
        } else {
#pragma unroll
            for (int i = 0; i < RG_ITEMS; ++i) e[i] = a0 + i < cnt ? kp[a0 + i] : 0ull;
        }
        if (a0 > 0) prev_e = kp[a0 - 1];
        if (a0 + RG_ITEMS < cnt) next_e = kp[a0 + RG_ITEMS];
    }
    __syncthreads(); // s_edge
    GroupView<RG_ITEMS> v;
    group_view<ROUND0, RG_ITEMS>(e, prev_e, next_e, a0, cnt, s_edge[0], s_edge[1], s_edge[2], s_i, v);
    const uint32_t n = A.n_arr[b];
    const bool done = sorted_len_of(A, b) >= n;
    const uint64_t o = (uint64_t)b * A.stride;
    uint32_t kg = 0; // kept in the bag
    unsigned gmask = 0, lmask = 0, fmask = 0;
    uint32_t tmin = 0xffffffffu;
    if (a0 < cnt) {
        uint32_t *rank = A.rank + o;
#pragma unroll
        for (int i = 0; i < RG_ITEMS; ++i) {
            if (a0 + i < cnt) {
                rank[(uint32_t)(e[i] & VAL_MASK)] = v.P[i];
                if (v.size[i] == 1 || done) fmask |= 1u << i;
                else if (A.split && (v.P[i] >> SA_TILE_BITS) == ((v.P[i] + v.size[i] - 1) >> SA_TILE_BITS)) { lmask |= 1u << i; tmin = min(tmin, v.P[i] >> SA_TILE_BITS); }
                else { gmask |= 1u << i; ++kg; }
            }
        }
        // final rotations: L[position] = T[rotation - 1]; loads first, then stores
        const uint8_t *T = A.T + o;
        uint8_t *L = A.L + o;
        uint8_t tb[RG_ITEMS];
#pragma unroll
        for (int i = 0; i < RG_ITEMS; ++i) {
            const uint32_t val = (uint32_t)(e[i] & VAL_MASK);
            tb[i] = (fmask & (1u << i)) ? T[val ? val - 1 : n - 1] : (uint8_t)0;
        }
#pragma unroll
        for (int i = 0; i < RG_ITEMS; ++i)
            if (fmask & (1u << i)) L[v.p[i]] = tb[i];
    }
    if (tmin != 0xffffffffu) atomicMin(&s_tmain, tmin);
    uint32_t total_g, total_m;
    uint32_t off = block_excl_sum(kg, s_u, &total_g); // its barriers also publish s_tmain
    // local kept elements: those of the CTA's main SA tile get one reservation for the whole CTA, stragglers one per run
    const uint32_t tmain = s_tmain;
    unsigned mmask = 0;
#pragma unroll
    for (int i = 0; i < RG_ITEMS; ++i)
        if ((lmask & (1u << i)) && (v.P[i] >> SA_TILE_BITS) == tmain) mmask |= 1u << i;
    uint32_t offm = block_excl_sum((uint32_t)__popc(mmask), s_u, &total_m);
    uint32_t *lc = A.loc_cnt + (uint64_t)b * A.tiles_per_block;
    if (threadIdx.x == 0) {
        s_base = total_g ? atomicAdd(&bag_cnt[b], total_g) : 0u;
        s_lbase = total_m ? atomicAdd(&lc[tmain], total_m) : 0u;
        if (total_m && s_lbase == 0) A.work_out[atomicAdd(A.work_out_cnt, 1u)] = b * (uint32_t)A.tiles_per_block + tmain;
    }
    const unsigned omask = lmask & ~mmask;
    if (omask) {
        uint32_t tl[RG_ITEMS], len[RG_ITEMS];
#pragma unroll
        for (int i = 0; i < RG_ITEMS; ++i) tl[i] = (omask & (1u << i)) ? (v.P[i] >> SA_TILE_BITS) : 0xffffffffu;
#pragma unroll
        for (int i = RG_ITEMS - 1; i >= 0; --i)
            len[i] = tl[i] == 0xffffffffu ? 0u : 1u + ((i + 1 < RG_ITEMS && tl[(i + 1) % RG_ITEMS] == tl[i]) ? len[(i + 1) % RG_ITEMS] : 0u);
        uint64_t *reg = nullptr;
#pragma unroll
        for (int i = 0; i < RG_ITEMS; ++i) {
            if (tl[i] != 0xffffffffu) {
                if (i == 0 || tl[(i + RG_ITEMS - 1) % RG_ITEMS] != tl[i]) {
                    const uint32_t at = atomicAdd(&lc[tl[i]], len[i]);
                    if (at == 0) A.work_out[atomicAdd(A.work_out_cnt, 1u)] = b * (uint32_t)A.tiles_per_block + tl[i];
                    reg = A.loc + o + (uint64_t)tl[i] * RG_TILE + at;
                }
                *reg++ = ((uint64_t)(v.P[i] & (RG_TILE - 1)) << GPOS_SHIFT) | (e[i] & VAL_MASK);
            }
        }
        atomicAdd(&s_lother, (uint32_t)__popc(omask));
    }
    __syncthreads();
    if (gmask) {
        uint64_t *bg = bag + o + s_base;
#pragma unroll
        for (int i = 0; i < RG_ITEMS; ++i)
            if (gmask & (1u << i)) bg[off++] = ((uint64_t)v.P[i] << GPOS_SHIFT) | (e[i] & VAL_MASK);
    }
    if (mmask) {
        uint64_t *reg = A.loc + o + (uint64_t)tmain * RG_TILE + s_lbase;
#pragma unroll
        for (int i = 0; i < RG_ITEMS; ++i)
            if (mmask & (1u << i)) reg[offm++] = ((uint64_t)(v.P[i] & (RG_TILE - 1)) << GPOS_SHIFT) | (e[i] & VAL_MASK);
    }
    if (threadIdx.x == 0) {
        const uint32_t kept = total_g + total_m + s_lother;
        if (kept) atomicAdd(A.totals, (unsigned long long)kept);
    }
}

// max / sum over blocks of the new bag sizes (grid sizing and loop control on the host)
__global__ void __launch_bounds__(256) k_bwt_roundstats(const uint32_t *__restrict__ bag_cnt, uint32_t n_blocks, unsigned long long *__restrict__ totals)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t c = b < n_blocks ? bag_cnt[b] : 0u;
    const uint32_t mx = __reduce_max_sync(FULL, c), sm = __reduce_add_sync(FULL, c);
    if (lane_id() == 0 && mx) {
        atomicMax(totals + 1, (unsigned long long)mx);
        atomicAdd(totals + 2, (unsigned long long)sm);
    }
}

// ------------------------------------------------------------------------------------------------ local regions
// One CTA per (SA tile, block): sort the region's elements by (group position, rank of the second half) in shared memory
// (LSD radix, warp match_any multi-split, digits that are constant over the region skipped), regroup them exactly like
// k_bwt_regroup does, and write the still-ambiguous ones back to the front of the same region.
#ifndef BZ_LS_THREADS
#define BZ_LS_THREADS 256
#endif
#ifndef BZ_LS_MINB
#define BZ_LS_MINB 4
#endif
constexpr int LS_THREADS = BZ_LS_THREADS;
constexpr int LS_ITEMS = RG_TILE / LS_THREADS; // 4
constexpr int LS_WARPS = LS_THREADS / 32;
#ifndef BZ_LS_SMALL_FROM
#define BZ_LS_SMALL_FROM 32
#endif
#ifndef BZ_LS_WARP_MIN
#define BZ_LS_WARP_MIN 64
#endif
constexpr uint32_t LS_SMALL_FROM = BZ_LS_SMALL_FROM;  // try the counting path once rotations are ordered by at least this many bytes
constexpr int LS_SMALL_MAX = 32;        // ... and no group of the region is larger than this

static_assert(LS_THREADS == 256, "one thread per digit in the scan of k_bwt_local");
struct LocalSmem {
    uint64_t e[RG_TILE];
    uint32_t warp_hist[2][LS_WARPS][256]; // double buffered: the next pass's counters are cleared while this pass runs
    uint32_t tile_base[256];
    uint32_t wtot[LS_WARPS];
    uint32_t scratch[34];
    int scr_i[96];
    uint32_t red[2 * LS_WARPS];
};

__global__ void __launch_bounds__(LS_THREADS, BZ_LS_MINB) k_bwt_local(RoundArgs A)
{
    __shared__ __align__(16) LocalSmem sm;
    const uint32_t rid = A.work_in[blockIdx.x];
    const uint32_t b = rid / (uint32_t)A.tiles_per_block, t = rid % (uint32_t)A.tiles_per_block;
    uint32_t *cnt_p = A.loc_cnt + (uint64_t)b * A.tiles_per_block + t;
    const uint32_t c = *cnt_p;
    if (c == 0) return;
    const uint64_t o = (uint64_t)b * A.stride;
    uint64_t *reg = A.loc + o + (uint64_t)t * RG_TILE;
    const int lane = lane_id(), w = warp_id();
    // ---- load; which key bits differ inside the region?
    uint32_t or_lo = 0, and_lo = 0xffffffffu;
    for (uint32_t j = threadIdx.x; j < c; j += LS_THREADS) {
        const uint64_t x = reg[j];
        sm.e[j] = x;
        const uint32_t k = (uint32_t)(x >> VAL_BITS); // 31 key bits
        or_lo |= k; and_lo &= k;
    }
    or_lo = __reduce_or_sync(FULL, or_lo);
    and_lo = __reduce_and_sync(FULL, and_lo);
    if (lane == 0) { sm.red[w] = or_lo; sm.red[LS_WARPS + w] = and_lo; }
    __syncthreads();
    uint32_t diff;
    {
        uint32_t x = 0, y = 0xffffffffu;
#pragma unroll
        for (int ww = 0; ww < LS_WARPS; ++ww) { x |= sm.red[ww]; y &= sm.red[LS_WARPS + ww]; }
        diff = x ^ y;
    }
    // ---- deep rounds: groups are mostly a handful of rotations (long repeats keep many groups of 2..9 alive for ~log n rounds).
    // If the region is still in group order (kept elements are; appended ones may not be) and no group is larger than
    // LS_SMALL_MAX, every element finds its place by counting the smaller keys inside its own group: no radix passes at all.
    bool sorted_small = false;
    if (sorted_len_of(A, b) >= LS_SMALL_FROM) {
        const uint32_t a0 = threadIdx.x * LS_ITEMS;
        uint64_t e[LS_ITEMS];
        int last_s = -1, first_s = INT_INF;
        bool bad = false;
        unsigned gf = 0;
        if (a0 < c) {
            uint32_t pg = a0 > 0 ? (uint32_t)(sm.e[a0 - 1] >> GPOS_SHIFT) : 0u;
#pragma unroll
            for (int i = 0; i < LS_ITEMS; ++i) {
                const uint32_t a = a0 + i;
                e[i] = a < c ? sm.e[a] : 0ull;
                if (a < c) {
                    const uint32_t g = (uint32_t)(e[i] >> GPOS_SHIFT);
                    if (a == 0 || g != pg) { gf |= 1u << i; last_s = (int)a; if (first_s == INT_INF) first_s = (int)a; }
                    bad |= a > 0 && g < pg;
                    pg = g;
                }
            }
        }
        int prev_s, unused, next_s;
        block_scan_max2_minsuffix(last_s, -1, first_s, sm.scr_i, prev_s, unused, next_s);
        int gs[LS_ITEMS], ge[LS_ITEMS];
        int mx = 0;
        if (a0 < c) {
            int nxt = next_s == INT_INF ? (int)c : next_s;
#pragma unroll
            for (int i = LS_ITEMS - 1; i >= 0; --i) {
                ge[i] = nxt - 1;
                if (gf & (1u << i)) nxt = (int)(a0 + i);
            }
            int run = prev_s;
#pragma unroll
            for (int i = 0; i < LS_ITEMS; ++i) {
                if (gf & (1u << i)) run = (int)(a0 + i);
                gs[i] = run;
                if (a0 + i < c) mx = max(mx, ge[i] - gs[i] + 1);
            }
        }
        if (!__syncthreads_or((bad || mx > LS_SMALL_MAX) ? 1 : 0)) {
            uint32_t np[LS_ITEMS];
            if (a0 < c) {
#pragma unroll
                for (int i = 0; i < LS_ITEMS; ++i) {
                    const int j = (int)(a0 + i);
                    if (j < (int)c) {
                        const uint32_t kj = (uint32_t)(e[i] >> VAL_BITS);
                        uint32_t r = 0;
                        for (int q = gs[i]; q <= ge[i]; ++q) {
                            const uint32_t kq = (uint32_t)(sm.e[q] >> VAL_BITS);
                            r += (kq < kj || (kq == kj && q < j)) ? 1u : 0u;
                        }
                        np[i] = (uint32_t)gs[i] + r;
                    }
                }
            }
            __syncthreads();
            if (a0 < c) {
#pragma unroll
                for (int i = 0; i < LS_ITEMS; ++i)
                    if (a0 + i < c) sm.e[np[i]] = e[i];
            }
            __syncthreads();
            sorted_small = true;
        }
    }
    const unsigned lt = (1u << lane) - 1u;
    // the sort uses as many warps as have at least two rows of 32 elements to rank; warp w owns elements [w*chunk, (w+1)*chunk)
    const uint32_t nwu = min((uint32_t)LS_WARPS, (c + BZ_LS_WARP_MIN - 1) / BZ_LS_WARP_MIN);
    const uint32_t chunk = ((c + nwu - 1) / nwu + 31) & ~31u;
    const uint32_t wbase = (uint32_t)w * chunk;
    const uint32_t wcnt = ((uint32_t)w < nwu && wbase < c) ? min(chunk, c - wbase) : 0u;
    for (uint32_t i = threadIdx.x; i < nwu * 256; i += LS_THREADS) (&sm.warp_hist[0][0][0])[i] = 0;
    __syncthreads();
    int hb = 0;
    // The first digit that differs has no earlier order to preserve: its pass need not be stable, so elements take their slots
    // with one shared-memory atomic each (no warp match, no per-warp counters).
    int shift_first = sorted_small ? 32 : 0;
    while (shift_first < 32 && ((diff >> shift_first) & 0xffu) == 0) shift_first += 8;
    if (shift_first < 32) {
        const int sh = shift_first + VAL_BITS;
        uint32_t *h0 = sm.warp_hist[1][0];
        h0[threadIdx.x] = 0;
        uint64_t key[LS_ITEMS];
        uint16_t rnk[LS_ITEMS];
#pragma unroll
        for (int i = 0; i < LS_ITEMS; ++i) {
            const uint32_t q = i * 32 + lane;
            key[i] = q < wcnt ? sm.e[wbase + q] : 0ull;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < LS_ITEMS; ++i)
            if ((uint32_t)i * 32 + lane < wcnt) rnk[i] = (uint16_t)atomicAdd(&h0[(uint32_t)(key[i] >> sh) & 0xffu], 1u);
        __syncthreads();
        const uint32_t sum = h0[threadIdx.x];
        const uint32_t incl = warp_incl_sum(sum);
        if (lane == 31) sm.wtot[w] = incl;
        __syncthreads();
        {
            uint32_t before = 0;
#pragma unroll
            for (int ww = 0; ww < LS_WARPS; ++ww)
                if (ww < w) before += sm.wtot[ww];
            sm.tile_base[threadIdx.x] = before + incl - sum;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < LS_ITEMS; ++i)
            if ((uint32_t)i * 32 + lane < wcnt) sm.e[sm.tile_base[(uint32_t)(key[i] >> sh) & 0xffu] + rnk[i]] = key[i];
        __syncthreads();
    }
    for (int shift = shift_first + 8; shift < 32; shift += 8) {
        if (((diff >> shift) & 0xffu) == 0) continue; // uniform over the CTA
        const int sh = shift + VAL_BITS;
        uint32_t(*hist)[256] = sm.warp_hist[hb];
        uint64_t key[LS_ITEMS];
        uint32_t dg[LS_ITEMS];
        unsigned peers[LS_ITEMS];
        uint16_t rnk[LS_ITEMS];
#pragma unroll
        for (int i = 0; i < LS_ITEMS; ++i) {
            const uint32_t q = i * 32 + lane;
            key[i] = q < wcnt ? sm.e[wbase + q] : 0ull;
            dg[i] = q < wcnt ? ((uint32_t)(key[i] >> sh) & 0xffu) : 0xffffffffu;
        }
#pragma unroll
        for (int i = 0; i < LS_ITEMS; ++i)
            if ((uint32_t)i * 32 < wcnt) peers[i] = __match_any_sync(FULL, dg[i]);
        for (uint32_t i = threadIdx.x; i < nwu * 256; i += LS_THREADS) (&sm.warp_hist[hb ^ 1][0][0])[i] = 0;
#pragma unroll
        for (int i = 0; i < LS_ITEMS; ++i) {
            if ((uint32_t)i * 32 < wcnt) { // warp-uniform
                const uint32_t d = dg[i];
                const int leader = __ffs(peers[i]) - 1;
                uint32_t old = 0;
                if (d != 0xffffffffu && lane == leader) {
                    old = hist[w][d];
                    hist[w][d] = old + __popc(peers[i]);
                }
                old = __shfl_sync(FULL, old, leader);
                rnk[i] = (uint16_t)(old + __popc(peers[i] & lt));
                __syncwarp();
            }
        }
        __syncthreads();
        uint32_t sum = 0, incl;
        {
            const uint32_t d = threadIdx.x;
#pragma unroll
            for (int ww = 0; ww < LS_WARPS; ++ww) {
                if ((uint32_t)ww < nwu) {
                    const uint32_t x = hist[ww][d];
                    hist[ww][d] = sum;
                    sum += x;
                }
            }
            incl = warp_incl_sum(sum);
            if (lane == 31) sm.wtot[w] = incl;
        }
        __syncthreads();
        {
            uint32_t before = 0;
#pragma unroll
            for (int ww = 0; ww < LS_WARPS; ++ww)
                if (ww < w) before += sm.wtot[ww];
            sm.tile_base[threadIdx.x] = before + incl - sum;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < LS_ITEMS; ++i) {
            const uint32_t d = dg[i];
            if ((uint32_t)i * 32 < wcnt && d != 0xffffffffu) sm.e[sm.tile_base[d] + hist[w][d] + rnk[i]] = key[i];
        }
        __syncthreads();
        hb ^= 1;
    }
    // ---- regroup inside the region (positions relative to the tile)
    const uint32_t a0 = threadIdx.x * LS_ITEMS;
    uint64_t e[LS_ITEMS];
#pragma unroll
    for (int i = 0; i < LS_ITEMS; ++i) e[i] = a0 + i < c ? sm.e[a0 + i] : 0ull;
    const uint64_t prev_e = (a0 > 0 && a0 < c) ? sm.e[a0 - 1] : 0ull;
    const uint64_t next_e = a0 + LS_ITEMS < c ? sm.e[a0 + LS_ITEMS] : 0ull;
    GroupView<LS_ITEMS> v;
    group_view<false, LS_ITEMS>(e, prev_e, next_e, a0, c, 0u, 0u, c - 1, sm.scr_i, v);
    const uint32_t n = A.n_arr[b];
    const bool done = sorted_len_of(A, b) >= n;
    const uint32_t tbase = t * RG_TILE;
    unsigned kmask = 0, fmask = 0;
    if (a0 < c) {
        const uint8_t *T = A.T + o;
        uint8_t *L = A.L + o;
        uint32_t *rank = A.rank + o;
#pragma unroll
        for (int i = 0; i < LS_ITEMS; ++i) {
            if (a0 + i < c) {
                rank[(uint32_t)(e[i] & VAL_MASK)] = tbase + v.P[i];
                if (v.size[i] == 1 || done) fmask |= 1u << i;
                else kmask |= 1u << i;
            }
        }
        uint8_t tb[LS_ITEMS];
#pragma unroll
        for (int i = 0; i < LS_ITEMS; ++i) {
            const uint32_t val = (uint32_t)(e[i] & VAL_MASK);
            tb[i] = (fmask & (1u << i)) ? T[val ? val - 1 : n - 1] : (uint8_t)0;
        }
#pragma unroll
        for (int i = 0; i < LS_ITEMS; ++i)
            if (fmask & (1u << i)) L[tbase + v.p[i]] = tb[i];
    }
    uint32_t total;
    uint32_t off = block_excl_sum((uint32_t)__popc(kmask), sm.scratch, &total);
#pragma unroll
    for (int i = 0; i < LS_ITEMS; ++i)
        if (kmask & (1u << i)) reg[off++] = ((uint64_t)v.P[i] << GPOS_SHIFT) | (e[i] & VAL_MASK);
    if (threadIdx.x == 0) {
        *cnt_p = total;
        if (total) {
            atomicAdd(A.totals, (unsigned long long)total);
            A.work_out[atomicAdd(A.work_out_cnt, 1u)] = rid;
        }
    }
}

// local regions: element |= rank[(i + h) mod n] << 20
__global__ void __launch_bounds__(256) k_bwt_gather_local(RoundArgs A)
{
    const uint32_t rid = A.work_in[blockIdx.x];
    const uint32_t b = rid / (uint32_t)A.tiles_per_block, t = rid % (uint32_t)A.tiles_per_block;
    const uint32_t c = A.loc_cnt[(uint64_t)b * A.tiles_per_block + t];
    if (c == 0) return;
    const uint32_t n = A.n_arr[b], h = sorted_len_of(A, b);
    const uint64_t o = (uint64_t)b * A.stride;
    uint64_t *reg = A.loc + o + (uint64_t)t * RG_TILE;
    const uint32_t *rank = A.rank + o;
    for (uint32_t j = threadIdx.x; j < c; j += 256) {
        const uint64_t e = reg[j];
        uint32_t p = (uint32_t)(e & VAL_MASK) + h; // h < n for every block that is still active
        if (p >= n) p -= n;
        reg[j] = e | ((uint64_t)rank[p] << VAL_BITS);
    }
}

// bag: element |= rank[(i + h) mod n] << 20
__global__ void __launch_bounds__(256) k_bwt_gather(uint64_t *__restrict__ keys, const uint32_t *__restrict__ counts, const uint32_t *__restrict__ n_arr,
                                                     uint64_t stride, const uint32_t *__restrict__ rank, const uint32_t *__restrict__ h0, uint32_t round)
{
    const uint32_t b = blockIdx.y;
    const uint32_t cnt = counts[b];
    uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= cnt) return;
    const uint32_t n = n_arr[b], h = h0[b] << round;
    const uint64_t o = (uint64_t)b * stride;
    const uint64_t e = keys[o + a];
    uint32_t p = (uint32_t)(e & VAL_MASK) + h; // h < n for every block that is still active
    if (p >= n) p -= n;
    keys[o + a] = e | ((uint64_t)rank[o + p] << VAL_BITS);
}

// origPtr = rank[0] = first position of rotation 0's tie group (kernel.cpp:3114-3116).  A block that was sorted through its root
// (k_bwt_period: n = m * p) has m equal rotations in every group: its groups start m times further on.
__global__ void __launch_bounds__(128) k_bwt_finish(const uint32_t *__restrict__ rank, const uint32_t *__restrict__ n_arr, const uint32_t *__restrict__ root_arr,
                                                     uint64_t stride, uint32_t n_blocks, uint32_t *__restrict__ orig)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const uint32_t n = n_arr[b], p = root_arr[b];
    orig[b] = n ? rank[(uint64_t)b * stride] * (n / p) : 0u;
}

// ------------------------------------------------------------------------------------------------ non-primitive blocks
// A block that is a power of a shorter word, T = w^m (all zeros after RLE1: (0,0,0,0,251)^m; short-period patterns), never resolves
// under prefix doubling: every rotation has m - 1 equal twins, so all ceil(log2 n) rounds run over the whole block (SURVEY 8d:
// rho = 19 for zeros).  Its BWT follows from the root's: the sorted rotations of T are the sorted rotations of w, each m times, so
// L_T[k * m + j] = L_w[k] and the first position of rotation 0's group is m times its position among w's rotations -- the origPtr
// convention of the oracle for such blocks.  k_bwt_period finds the smallest p | n, p <= BP_MAX, with T[i] = T[i + p] for
// all i (root_len[b] = p, else n); the sort then runs on the first p bytes only and k_bwt_expand writes the block's last column.
constexpr uint32_t BP_MAX = 4096; // longest root looked for (longer exact periods are not worth a search: such blocks sort normally)
constexpr int BP_THREADS = 256;

__global__ void __launch_bounds__(BP_THREADS) k_bwt_period(const uint8_t *__restrict__ rle, const uint32_t *__restrict__ n_arr, uint64_t stride,
                                                            uint32_t *__restrict__ root_len)
{
    __shared__ uint32_t s_cand[128];
    __shared__ uint32_t s_ncand, s_best, s_fail;
    const uint32_t b = blockIdx.x, n = n_arr[b];
    const uint8_t *T = rle + (uint64_t)b * stride;
    if (threadIdx.x == 0) { s_ncand = 0; s_best = n; }
    __syncthreads();
    // candidates: divisors of n whose first 32 positions repeat
    for (uint32_t p = threadIdx.x + 1; p <= BP_MAX && 2 * p <= n; p += BP_THREADS) {
        if (n % p) continue;
        bool ok = true;
        for (uint32_t j = 0; j < 32 && p + j < n && ok; ++j) ok = T[j] == T[p + j];
        if (ok) {
            const uint32_t k = atomicAdd(&s_ncand, 1u);
            if (k < 128) s_cand[k] = p;
        }
    }
    __syncthreads();
    uint32_t nc = min(s_ncand, 128u);
    // smallest candidate first; the first one that holds for the whole block is the root
    while (nc) {
        uint32_t mn = 0xffffffffu, at = 0;
        for (uint32_t k = 0; k < nc; ++k)
            if (s_cand[k] < mn) { mn = s_cand[k]; at = k; }
        __syncthreads();
        if (threadIdx.x == 0) { s_cand[at] = s_cand[nc - 1]; s_fail = 0; }
        __syncthreads();
        bool bad = false;
        for (uint32_t i = threadIdx.x; i + mn < n && !bad; i += BP_THREADS) bad = T[i] != T[i + mn];
        if (bad) s_fail = 1;
        __syncthreads();
        if (!s_fail) {
            if (threadIdx.x == 0) s_best = mn;
            break;
        }
        --nc;
        __syncthreads();
    }
    __syncthreads();
    if (threadIdx.x == 0) root_len[b] = s_best;
}

// the root's last column (first p bytes of L) aside, then every byte of it m = n / p times
__global__ void __launch_bounds__(256) k_bwt_expand_copy(const uint8_t *__restrict__ L, const uint32_t *__restrict__ n_arr, const uint32_t *__restrict__ root_arr,
                                                          uint64_t stride, uint8_t *__restrict__ scratch)
{
    const uint32_t b = blockIdx.y, p = root_arr[b];
    if (p == n_arr[b]) return;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < p) scratch[(uint64_t)b * stride + i] = L[(uint64_t)b * stride + i];
}
__global__ void __launch_bounds__(256) k_bwt_expand(uint8_t *__restrict__ L, const uint32_t *__restrict__ n_arr, const uint32_t *__restrict__ root_arr, uint64_t stride,
                                                     const uint8_t *__restrict__ scratch)
{
    const uint32_t b = blockIdx.y, n = n_arr[b], p = root_arr[b];
    if (p == n) return;
    const uint32_t m = n / p;
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) L[(uint64_t)b * stride + q] = scratch[(uint64_t)b * stride + q / m];
}

} // namespace bz
