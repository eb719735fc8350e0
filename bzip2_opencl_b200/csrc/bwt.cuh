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
struct ScatterSmem {
    uint32_t warp_hist[RS_WARPS][256];
    uint32_t tile_base[256];
    uint32_t gofs[256];
    uint32_t scratch[34];
    uint64_t keys[RS_TILE];
};

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
    uint32_t key = 0, p = i;
    for (int k = 0; k < nbytes; ++k) {
        key = (key << 8) | T[p];
        if (++p == n) p = 0;
    }
    keys[(uint64_t)b * stride + i] = ((uint64_t)key << VAL_BITS) | i;
}

// ------------------------------------------------------------------------------------------------ regroup
// Sorted elements of the active list of each block -> scatter SA, find new groups, update ranks, compact the
// still-ambiguous rotations into the next round's sort input.
struct RegroupArgs {
    const uint32_t *counts;   // [G] active elements sorted this round (small-group part first, then the rest)
    const uint32_t *n_arr;    // [G] block lengths
    uint64_t stride;
    const uint32_t *pos_old;  // position in SA of active element a (nullptr: identity, round 0)
    uint32_t *sa;             // [G*stride]
    uint32_t *rank;           // [G*stride]
    int32_t *tile_last;       // [G*tiles] last flagged index per tile (-1: none)
    uint32_t *tile_keep_s;    // [G*tiles] kept elements of tile-local groups per tile, later exclusive prefix
    uint32_t *tile_keep_b;    // [G*tiles] kept elements of the other groups per tile, later exclusive prefix
    int32_t *tile_carry;      // [G*tiles] last flagged index before the tile
    int tiles_per_block;      // regroup tiles per block
    uint32_t *cnt_t_new;      // [G] all still-active elements
    uint32_t *cnt_s_new;      // [G] ... of which in tile-local groups (sorted in shared memory next round)
    uint32_t *cnt_b_new;      // [G] ... the rest (global radix sort next round)
    uint32_t *pos_new;        // compacted positions
    uint64_t *keys_new;       // (list slot of the group start << (20 + nbits)) | rotation index; k_bwt_gather ORs in rank of the second half << 20
    int nbits;
    int split;                // 0: put every group in the global-sort part
    uint32_t sorted_len;      // rotations are now ordered by their first sorted_len bytes
    unsigned long long *totals; // [0] sum of cnt_t_new, [3] max cnt_t_new, [8] max cnt_s_new, [9] max cnt_b_new, [10] sum cnt_b_new
};

__device__ __forceinline__ void load_flags(const uint64_t *kp, uint32_t a0, uint32_t cnt, bool (&flag)[RG_ITEMS], bool (&nxt)[RG_ITEMS],
                                           uint32_t (&val)[RG_ITEMS])
{
    // flag[i]: element a0+i starts a group; nxt[i]: element a0+i ends a group; val[i]: its rotation index
    uint64_t prev = a0 > 0 && a0 - 1 < cnt ? (kp[a0 - 1] >> VAL_BITS) : 0ull;
#pragma unroll
    for (int i = 0; i < RG_ITEMS; ++i) {
        uint32_t a = a0 + i;
        uint64_t e = a < cnt ? kp[a] : 0ull;
        uint64_t cur = e >> VAL_BITS;
        val[i] = (uint32_t)(e & VAL_MASK);
        flag[i] = (a == 0) || (cur != prev);
        prev = cur;
    }
#pragma unroll
    for (int i = 0; i < RG_ITEMS - 1; ++i) nxt[i] = (a0 + i + 1 >= cnt) || flag[i + 1];
    {
        uint32_t a = a0 + RG_ITEMS;
        nxt[RG_ITEMS - 1] = (a >= cnt) || ((kp[a] >> VAL_BITS) != prev);
    }
}

// Per element: does its (new) group start AND end inside this regroup tile?  Such "tile-local" groups (<= RG_TILE elements)
// are sorted in shared memory next round; everything else goes through the global radix sort.
struct TileView {
    bool flag[RG_ITEMS], nxt[RG_ITEMS], local[RG_ITEMS];
    uint32_t val[RG_ITEMS];
    int last;          // last flagged index owned by this thread (-1)
    int prev_last;     // last flagged index of the threads before it, inside the tile (-1)
};

__device__ __forceinline__ void tile_view(const uint64_t *kp, uint32_t a0, uint32_t cnt, int split, int *s_i, TileView &v)
{
    v.last = -1;
    int first_end = 0x7fffffff;
    if (a0 < cnt) {
        load_flags(kp, a0, cnt, v.flag, v.nxt, v.val);
#pragma unroll
        for (int i = RG_ITEMS - 1; i >= 0; --i)
            if (a0 + i < cnt && v.nxt[i]) first_end = (int)(a0 + i);
#pragma unroll
        for (int i = 0; i < RG_ITEMS; ++i)
            if (a0 + i < cnt && v.flag[i]) v.last = (int)(a0 + i);
    }
    v.prev_last = block_excl_max(v.last, s_i, nullptr);
    const int next_end = block_excl_min_suffix(first_end, s_i);
    if (a0 < cnt) {
        bool started = v.prev_last >= 0;
#pragma unroll
        for (int i = 0; i < RG_ITEMS; ++i) { started = started || v.flag[i]; v.local[i] = started; }
        bool ended = next_end != 0x7fffffff;
#pragma unroll
        for (int i = RG_ITEMS - 1; i >= 0; --i) {
            if (a0 + i < cnt) ended = ended || v.nxt[i];
            v.local[i] = v.local[i] && ended && split != 0;
        }
    }
}

__global__ void __launch_bounds__(RG_THREADS) k_bwt_flags(const uint64_t *__restrict__ keys, RegroupArgs A)
{
    __shared__ int s_i[34];
    __shared__ int s_last[RG_THREADS / 32];
    __shared__ uint32_t s_ks[RG_THREADS / 32], s_kb[RG_THREADS / 32];
    const uint32_t b = blockIdx.y, tile = blockIdx.x;
    const uint32_t cnt = A.counts[b];
    const uint32_t base = tile * RG_TILE;
    if (base >= cnt) return;
    const uint64_t *kp = keys + (uint64_t)b * A.stride;
    const uint32_t a0 = base + threadIdx.x * RG_ITEMS;
    TileView v;
    tile_view(kp, a0, cnt, A.split, s_i, v);
    uint32_t ks = 0, kb = 0;
    if (a0 < cnt) {
        const uint32_t *pp = A.pos_old ? A.pos_old + (uint64_t)b * A.stride : nullptr;
        uint32_t *sa = A.sa + (uint64_t)b * A.stride;
        const bool done = A.sorted_len >= A.n_arr[b];
#pragma unroll
        for (int i = 0; i < RG_ITEMS; ++i) {
            uint32_t a = a0 + i;
            if (a < cnt) {
                const bool single = v.flag[i] && v.nxt[i];
                if (!single) { if (v.local[i]) ++ks; else ++kb; }
                // SA is only read by k_bwt_finish: write a rotation's slot when it becomes final (alone in its group, or the
                // block stops here); rotations that stay active are written by a later round
                if (single || done) sa[pp ? pp[a] : a] = v.val[i];
            }
        }
    }
    int last = __reduce_max_sync(FULL, v.last);
    ks = __reduce_add_sync(FULL, ks);
    kb = __reduce_add_sync(FULL, kb);
    if (lane_id() == 0) { s_last[warp_id()] = last; s_ks[warp_id()] = ks; s_kb[warp_id()] = kb; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int l = -1; uint32_t s = 0, k = 0;
        for (int w = 0; w < RG_THREADS / 32; ++w) { l = max(l, s_last[w]); s += s_ks[w]; k += s_kb[w]; }
        const uint64_t idx = (uint64_t)b * A.tiles_per_block + tile;
        A.tile_last[idx] = l;
        A.tile_keep_s[idx] = s;
        A.tile_keep_b[idx] = k;
    }
}

// one warp per block
__global__ void __launch_bounds__(256) k_bwt_tilescan(RegroupArgs A, uint32_t n_blocks)
{
    const uint32_t b = blockIdx.x * (blockDim.x >> 5) + warp_id();
    if (b >= n_blocks) return;
    const int lane = lane_id();
    const uint32_t cnt = A.counts[b];
    const uint32_t nt = (cnt + RG_TILE - 1) / RG_TILE;
    int carry_last = -1;
    uint32_t carry_s = 0, carry_b = 0;
    for (uint32_t t0 = 0; t0 < nt; t0 += 32) {
        uint32_t t = t0 + lane;
        uint64_t idx = (uint64_t)b * A.tiles_per_block + t;
        int l = t < nt ? A.tile_last[idx] : -1;
        uint32_t ks = t < nt ? A.tile_keep_s[idx] : 0, kb = t < nt ? A.tile_keep_b[idx] : 0;
        int li = warp_incl_max(l);
        uint32_t si = warp_incl_sum(ks), bi = warp_incl_sum(kb);
        int lprev = __shfl_up_sync(FULL, li, 1);
        if (lane == 0) lprev = -1;
        if (t < nt) {
            A.tile_carry[idx] = max(carry_last, lprev);
            A.tile_keep_s[idx] = carry_s + si - ks;
            A.tile_keep_b[idx] = carry_b + bi - kb;
        }
        carry_last = max(carry_last, __shfl_sync(FULL, li, 31));
        carry_s += __shfl_sync(FULL, si, 31);
        carry_b += __shfl_sync(FULL, bi, 31);
    }
    if (lane == 0) {
        const bool done = A.sorted_len >= A.n_arr[b]; // k_bwt_regroup writes no lists for a finished block
        A.cnt_s_new[b] = done ? 0u : carry_s;
        A.cnt_b_new[b] = done ? 0u : carry_b;
        const uint32_t nc = done ? 0u : carry_s + carry_b;
        A.cnt_t_new[b] = nc;
        if (nc) {
            atomicAdd(A.totals, (unsigned long long)nc);
            atomicMax(A.totals + 3, (unsigned long long)nc);
            atomicMax(A.totals + 8, (unsigned long long)carry_s);
            atomicMax(A.totals + 9, (unsigned long long)carry_b);
            atomicAdd(A.totals + 10, (unsigned long long)carry_b);
        }
    }
}

__global__ void __launch_bounds__(RG_THREADS) k_bwt_regroup(const uint64_t *__restrict__ keys, RegroupArgs A)
{
    __shared__ int s_i[34];
    __shared__ uint32_t s_u[34];
    const uint32_t b = blockIdx.y, tile = blockIdx.x;
    const uint32_t cnt = A.counts[b];
    const uint32_t base = tile * RG_TILE;
    if (base >= cnt) return;
    const uint64_t *kp = keys + (uint64_t)b * A.stride;
    const uint32_t a0 = base + threadIdx.x * RG_ITEMS;
    TileView v;
    tile_view(kp, a0, cnt, A.split, s_i, v);
    uint32_t ks = 0, kb = 0;
    if (a0 < cnt) {
#pragma unroll
        for (int i = 0; i < RG_ITEMS; ++i)
            if (a0 + i < cnt && !(v.flag[i] && v.nxt[i])) { if (v.local[i]) ++ks; else ++kb; }
    }
    const uint64_t tidx = (uint64_t)b * A.tiles_per_block + tile;
    const int prev_last = max(v.prev_last, A.tile_carry[tidx]);
    uint32_t off_s = A.tile_keep_s[tidx] + block_excl_sum(ks, s_u, nullptr);
    uint32_t off_b = A.cnt_s_new[b] + A.tile_keep_b[tidx] + block_excl_sum(kb, s_u, nullptr);
    if (a0 >= cnt) return;
    const uint32_t *pp = A.pos_old ? A.pos_old + (uint64_t)b * A.stride : nullptr;
    uint32_t *rank = A.rank + (uint64_t)b * A.stride;
    uint32_t *pn = A.pos_new + (uint64_t)b * A.stride;
    uint64_t *kn = A.keys_new + (uint64_t)b * A.stride;
    const int gshift = VAL_BITS + A.nbits;
    const bool done = A.sorted_len >= A.n_arr[b];
    int g = prev_last;
    uint32_t gpos = 0;
    bool gpos_valid = false;
#pragma unroll
    for (int i = 0; i < RG_ITEMS; ++i) {
        uint32_t a = a0 + i;
        if (a < cnt) {
            uint32_t mypos = pp ? pp[a] : a;
            if (v.flag[i]) { g = (int)a; gpos = mypos; gpos_valid = true; }
            else if (!gpos_valid) { gpos = pp ? pp[g] : (uint32_t)g; gpos_valid = true; }
            const uint32_t val = v.val[i];
            rank[val] = gpos;
            if (!done && !(v.flag[i] && v.nxt[i])) {
                // key high part = list slot of the group's first element: the list is [small groups | other groups], each in
                // SA order but not sorted as a whole, and a sort must leave every group in its own slots (pos_new is per slot).
                // All elements of a group are kept and share a class, so their slots are consecutive.
                const uint32_t o = v.local[i] ? off_s++ : off_b++;
                pn[o] = mypos;
                kn[o] = ((uint64_t)(o - (a - (uint32_t)g)) << gshift) | val;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ shared-memory sort of small groups
// The small-group part of an active list is a run of whole groups, each at most RG_TILE elements.  CTA k owns the groups that
// START in [k*LS_STRIDE, (k+1)*LS_STRIDE); they end inside the 2*LS_STRIDE window it loads.  Inside the window the group
// position is replaced by a dense group index, so (index, rank of the second half) fits 32 bits and is LSD-sorted in shared
// memory (same warp match_any multi-split as the global scatter, no global traffic between passes).
constexpr int LS_THREADS = 512;
constexpr int LS_ITEMS = 8;
constexpr int LS_WIN = LS_THREADS * LS_ITEMS; // 4096
constexpr int LS_STRIDE = LS_WIN / 2;         // == RG_TILE
constexpr int LS_WARPS = LS_THREADS / 32;
static_assert(LS_STRIDE == RG_TILE, "small groups are classified per regroup tile");

struct LocalSmem {
    uint64_t e[LS_WIN];
    uint32_t warp_hist[LS_WARPS][256];
    uint32_t tile_base[256];
    uint32_t gpos_of[LS_STRIDE + 2];
    uint32_t scratch[34];
    int first_start, end_start;
};

__global__ void __launch_bounds__(LS_THREADS) k_bwt_localsort(const uint64_t *__restrict__ src, uint64_t *__restrict__ dst, const uint32_t *__restrict__ cnt_s,
                                                               uint64_t stride, int nbits)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LocalSmem &sm = *reinterpret_cast<LocalSmem *>(smem_raw);
    const uint32_t b = blockIdx.y;
    const uint32_t cnt = cnt_s[b];
    const uint32_t base = blockIdx.x * LS_STRIDE;
    if (base >= cnt) return;
    const uint32_t wcnt = min((uint32_t)LS_WIN, cnt - base);
    const uint64_t *sp = src + (uint64_t)b * stride + base;
    const int gshift = VAL_BITS + nbits;
    const int lane = lane_id(), w = warp_id();
    for (uint32_t j = threadIdx.x; j < wcnt; j += LS_THREADS) sm.e[j] = sp[j];
    if (threadIdx.x == 0) { sm.first_start = 0x7fffffff; sm.end_start = (int)wcnt; }
    const uint64_t prev_g = base > 0 ? (sp[-1] >> gshift) : ~0ull;
    __syncthreads();
    // group starts among my 16 consecutive window elements
    const uint32_t j0 = threadIdx.x * LS_ITEMS;
    uint64_t el[LS_ITEMS];
    bool st[LS_ITEMS];
    {
        uint64_t pg = j0 == 0 ? prev_g : (j0 - 1 < wcnt ? (sm.e[j0 - 1] >> gshift) : 0ull);
#pragma unroll
        for (int i = 0; i < LS_ITEMS; ++i) {
            uint32_t j = j0 + i;
            el[i] = j < wcnt ? sm.e[j] : 0ull;
            uint64_t g = el[i] >> gshift;
            st[i] = j < wcnt && (g != pg || (base == 0 && j == 0));
            pg = g;
            if (st[i]) {
                if (j < (uint32_t)LS_STRIDE) atomicMin(&sm.first_start, (int)j);
                else atomicMin(&sm.end_start, (int)j);
            }
        }
    }
    __syncthreads();
    const int f = sm.first_start, eend = sm.end_start;
    if (f == 0x7fffffff) return; // no group starts in my stride: the groups here belong to the CTA on the left
    const uint32_t m = (uint32_t)(eend - f);
    // dense group index = (number of starts in [f, j]) - 1
    uint32_t mystarts = 0;
#pragma unroll
    for (int i = 0; i < LS_ITEMS; ++i) {
        int j = (int)(j0 + i);
        if (st[i] && j >= f && j < eend) ++mystarts;
    }
    uint32_t ngroups;
    uint32_t gid = block_excl_sum(mystarts, sm.scratch, &ngroups); // starts before my elements
    const uint64_t lowmask = (1ull << gshift) - 1ull;
#pragma unroll
    for (int i = 0; i < LS_ITEMS; ++i) {
        int j = (int)(j0 + i);
        if (j >= f && j < eend) {
            if (st[i]) { sm.gpos_of[gid] = (uint32_t)(el[i] >> gshift); ++gid; }
            sm.e[j - f] = ((uint64_t)(gid - 1) << gshift) | (el[i] & lowmask); // all reads of e[] happened before the last barrier
        }
    }
    __syncthreads();
    // LSD passes over bits [VAL_BITS, VAL_BITS + nbits + gbits)
    const int gbits = ngroups > 1 ? 32 - __clz((int)ngroups - 1) : 0;
    const int total_bits = nbits + gbits;
    const unsigned lt = (1u << lane) - 1u;
    for (int shift = VAL_BITS; shift < VAL_BITS + total_bits; shift += 8) {
        uint64_t key[LS_ITEMS];
        uint16_t rnk[LS_ITEMS];
#pragma unroll
        for (int i = 0; i < LS_ITEMS; ++i) {
            uint32_t li = w * (32 * LS_ITEMS) + i * 32 + lane;
            key[i] = li < m ? sm.e[li] : 0ull;
        }
        for (int i = threadIdx.x; i < LS_WARPS * 256; i += LS_THREADS) (&sm.warp_hist[0][0])[i] = 0;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < LS_ITEMS; ++i) {
            uint32_t li = w * (32 * LS_ITEMS) + i * 32 + lane;
            bool valid = li < m;
            uint32_t d = valid ? ((uint32_t)(key[i] >> shift) & 0xffu) : 0xffffffffu;
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
                for (int ww = 0; ww < LS_WARPS; ++ww) {
                    uint32_t t = sm.warp_hist[ww][d];
                    sm.warp_hist[ww][d] = sum;
                    sum += t;
                }
            }
            const uint32_t tb = block_excl_sum(sum, sm.scratch, nullptr); // threads >= 256 add 0 behind the digits
            if (d < 256) sm.tile_base[d] = tb;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < LS_ITEMS; ++i) {
            uint32_t li = w * (32 * LS_ITEMS) + i * 32 + lane;
            if (li < m) {
                uint32_t d = (uint32_t)(key[i] >> shift) & 0xffu;
                sm.e[sm.tile_base[d] + sm.warp_hist[w][d] + rnk[i]] = key[i];
            }
        }
        __syncthreads();
    }
    uint64_t *dp = dst + (uint64_t)b * stride + base + f;
    for (uint32_t j = threadIdx.x; j < m; j += LS_THREADS) {
        uint64_t x = sm.e[j];
        dp[j] = ((uint64_t)sm.gpos_of[(uint32_t)(x >> gshift)] << gshift) | (x & lowmask);
    }
}

// element |= rank[(i + h) mod n] << 20
__global__ void __launch_bounds__(256) k_bwt_gather(uint64_t *__restrict__ keys, const uint32_t *__restrict__ counts, const uint32_t *__restrict__ n_arr,
                                                     uint64_t stride, const uint32_t *__restrict__ rank, uint32_t h)
{
    const uint32_t b = blockIdx.y;
    const uint32_t cnt = counts[b];
    uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= cnt) return;
    const uint32_t n = n_arr[b];
    const uint64_t o = (uint64_t)b * stride;
    const uint64_t e = keys[o + a];
    uint32_t p = (uint32_t)(e & VAL_MASK) + h; // h < n for every block that is still active
    if (p >= n) p -= n;
    keys[o + a] = e | ((uint64_t)rank[o + p] << VAL_BITS);
}

// L[j] = T[(SA[j] - 1) mod n]; origPtr = rank[0]
__global__ void __launch_bounds__(256) k_bwt_finish(const uint8_t *__restrict__ rle, const uint32_t *__restrict__ sa, const uint32_t *__restrict__ rank,
                                                     const uint32_t *__restrict__ n_arr, uint64_t stride, uint8_t *__restrict__ L, uint32_t *__restrict__ orig)
{
    const uint32_t b = blockIdx.y;
    const uint32_t n = n_arr[b];
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t o = (uint64_t)b * stride;
    if (j == 0 && n > 0) orig[b] = rank[o];
    if (j >= n) return;
    uint32_t i = sa[o + j];
    L[o + j] = rle[o + (i ? i - 1 : n - 1)];
}

} // namespace bz
