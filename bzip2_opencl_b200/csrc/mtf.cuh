// mtf.cuh -- stage C: move-to-front + zero-run (RUNA/RUNB) coding of the BWT last column.
// Reference: MTFAndRLE2StageEncoder + valueToFront, /root/reference/kernel.cpp:2513-2649 (one sequential loop per block).
//
// MTF is a serial dependency chain; the device formulation cuts every block into tiles of MTF_TILE symbols:
//   k_mtf_recency     one warp per tile: the tile's distinct symbols by recency (a partial permutation), computed in parallel from
//                     last-occurrence positions -- what running MTF over the tile would leave at the front of the list
//   k_mtf_compose     one warp per block composes those in tile order -> the exact list at every tile start
//   k_mtf_ranks       one THREAD per tile re-runs MTF from its true start list and writes the ranks.  The first 16 list
//                     entries sit in four registers (typical BWT output resolves there); the other 240 live in shared memory
//                     and are walked 16 entries per step with per-byte SIMD compares (__vcmpeq4) and __byte_perm shifts.
//                     Symbols and ranks are staged through shared memory so global traffic stays in 64-byte pieces.
//   k_rle2_tiles / k_rle2_scan / k_rle2_emit  turn ranks into the RUNA/RUNB/rank+1/EOB symbol stream with an
//                     exclusive scan of emission counts, and histogram the symbols (mtfSymbolFrequencies).
#pragma once
#include "util.cuh"

namespace bz {

constexpr int MTF_TILE = 2048;
constexpr int MTF_WARPS = 8;

// dense symbol map of the block: map[v] = number of present values below v (kernel.cpp:2563-2571)
__device__ __forceinline__ void build_symbol_map(const uint32_t *present8, uint8_t *map /*256, shared*/)
{
    for (int v = threadIdx.x; v < 256; v += blockDim.x) {
        int c = 0;
        for (int w = 0; w < (v >> 5); ++w) c += __popc(present8[w]);
        c += __popc(present8[v >> 5] & ((1u << (v & 31)) - 1u));
        map[v] = (uint8_t)c;
    }
}

// ---- per-thread MTF over one tile.  The 256-entry list of every thread lives in shared memory as 32 little-endian
// 64-bit words (word k = entries 8k..8k+7), word 0 is kept in a register.  A lookup walks the list 8 entries at a time
// with a SWAR zero-byte test; the words it passes are shifted up by one entry on the way (that IS the move-to-front).
constexpr int MS_THREADS = 128;
constexpr int MS_CHUNK = 64;          // symbols staged per thread per step
constexpr int MS_STAGE_STRIDE = 68;   // bytes (17 words: conflict-free lane-strided access)
constexpr int MS_LIST_WORDS = 33;     // 32 words + 1 pad (conflict-free 64-bit lane-strided access)
constexpr uint64_t ONES8 = 0x0101010101010101ull;

__device__ __forceinline__ uint32_t zero_byte_index(uint64_t x) // index of the lowest zero byte of x, 8 if none
{
    uint64_t t = (x - ONES8) & ~x & 0x8080808080808080ull;
    return t ? (uint32_t)(__ffsll((long long)t) - 1) >> 3 : 8u;
}

__device__ __forceinline__ uint64_t low_bytes_mask(uint32_t bi) // bytes 0..bi
{
    return bi >= 7 ? ~0ull : ((1ull << (8 * (bi + 1))) - 1ull);
}

// ---- pass 1: the tile's distinct symbols ordered by recency (most recent first).  Fully parallel: last occurrence of every
// symbol with shared-memory atomicMax (only run ends write), then a counting rank over the <= 256 present symbols.
// lists[tile] receives the order in its first d = dcount[tile] entries; that is all k_mtf_compose reads.
constexpr int MR_WARPS = 8;
__global__ void __launch_bounds__(MR_WARPS * 32) k_mtf_recency(const uint8_t *__restrict__ L, const uint32_t *__restrict__ n_arr, uint64_t stride,
                                                                const uint32_t *__restrict__ present, uint64_t *__restrict__ lists,
                                                                uint32_t *__restrict__ dcount, int tiles_per_block)
{
    __shared__ uint8_t map[256];
    __shared__ int s_last[MR_WARPS][256];
    __shared__ int s_cpos[MR_WARPS][256];
    __shared__ __align__(8) uint8_t s_order[MR_WARPS][256];
    const uint32_t b = blockIdx.y;
    const uint32_t n = n_arr[b];
    if ((uint64_t)blockIdx.x * MR_WARPS * MTF_TILE >= n) return;
    build_symbol_map(present + b * 8, map);
    __syncthreads();
    const int lane = lane_id(), w = warp_id();
    const uint32_t tile = blockIdx.x * MR_WARPS + w;
    const uint32_t base = tile * MTF_TILE;
    if (base >= n) return;
    const uint32_t cnt = min((uint32_t)MTF_TILE, n - base);
    int *last = s_last[w], *cpos = s_cpos[w];
    uint8_t *order = s_order[w];
#pragma unroll
    for (int j = 0; j < 8; ++j) { last[lane * 8 + j] = -1; order[lane * 8 + j] = 0; }
    __syncwarp();
    const uint8_t *src = L + (uint64_t)b * stride + base;
    for (uint32_t k = 0; k < cnt; k += 32) {
        const uint32_t p = k + lane;
        const uint32_t sym = p < cnt ? map[src[p]] : 0xffffu;
        const uint32_t nxt = __shfl_down_sync(FULL, sym, 1);
        if (p < cnt && (lane == 31 || sym != nxt)) atomicMax(&last[sym], (int)p);
    }
    __syncwarp();
    // compact the present symbols: (position of last occurrence) per symbol
    int mine[8];
    uint32_t have = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { mine[j] = last[lane * 8 + j]; have += mine[j] >= 0; }
    uint32_t off = warp_incl_sum(have) - have;
    const uint32_t d = __shfl_sync(FULL, off + have, 31);
#pragma unroll
    for (int j = 0; j < 8; ++j) if (mine[j] >= 0) cpos[off++] = mine[j];
    __syncwarp();
    // rank = number of present symbols seen later
    uint32_t rk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t i = 0; i < d; ++i) {
        const int v = cpos[i];
#pragma unroll
        for (int j = 0; j < 8; ++j) rk[j] += v > mine[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) if (mine[j] >= 0) order[rk[j]] = (uint8_t)(lane * 8 + j);
    __syncwarp();
    const uint64_t lidx = ((uint64_t)b * tiles_per_block + tile) * 32;
    lists[lidx + lane] = *reinterpret_cast<const uint64_t *>(order + lane * 8);
    if (lane == 0) dcount[(uint64_t)b * tiles_per_block + tile] = d;
}

// ---- pass 2: one THREAD per tile re-runs MTF from its true start list (k_mtf_compose) and writes the ranks.
// The list is 16 x uint4 (16 entries each).  Entries 0..15 live in four registers; the rest in shared memory.  A lookup
// compares 16 entries per step with four per-byte SIMD compares (__vcmpeq4) and shifts the passed words by one entry with
// __byte_perm -- that shift IS the move-to-front.
constexpr int ML_STRIDE4 = 17; // uint4 per thread (16 + 1 pad: conflict-free 128-bit lane-strided access)

struct Quad { uint32_t x, y, z, w; };

// q shifted up by one entry with `in` entering at entry 0
__device__ __forceinline__ Quad quad_shift(const Quad &q, uint32_t in)
{
    Quad s;
    s.x = (q.x << 8) | in;
    s.y = __byte_perm(q.x, q.y, 0x6543);
    s.z = __byte_perm(q.y, q.z, 0x6543);
    s.w = __byte_perm(q.z, q.w, 0x6543);
    return s;
}
// entry index (0..15) of byte value `pat`/0x01010101 in q, or 16
__device__ __forceinline__ uint32_t quad_find(const Quad &q, uint32_t pat)
{
    uint32_t m0 = __vcmpeq4(q.x, pat), m1 = __vcmpeq4(q.y, pat), m2 = __vcmpeq4(q.z, pat), m3 = __vcmpeq4(q.w, pat);
    if ((m0 | m1 | m2 | m3) == 0) return 16;
    if (m0) return (uint32_t)(__ffs((int)m0) - 1) >> 3;
    if (m1) return 4 + ((uint32_t)(__ffs((int)m1) - 1) >> 3);
    if (m2) return 8 + ((uint32_t)(__ffs((int)m2) - 1) >> 3);
    return 12 + ((uint32_t)(__ffs((int)m3) - 1) >> 3);
}
// entries 0..p come from the shifted quad s, entries above p stay
__device__ __forceinline__ Quad quad_blend(const Quad &old, const Quad &s, uint32_t p)
{
    const uint32_t wi = p >> 2, bi = p & 3;
    const uint32_t m = bi == 3 ? 0xffffffffu : ((1u << (8 * (bi + 1))) - 1u);
    Quad r;
    r.x = wi > 0 ? s.x : ((old.x & ~m) | (s.x & m));
    r.y = wi > 1 ? s.y : (wi == 1 ? ((old.y & ~m) | (s.y & m)) : old.y);
    r.z = wi > 2 ? s.z : (wi == 2 ? ((old.z & ~m) | (s.z & m)) : old.z);
    r.w = wi == 3 ? ((old.w & ~m) | (s.w & m)) : old.w;
    return r;
}

__global__ void __launch_bounds__(MS_THREADS) k_mtf_ranks(const uint8_t *__restrict__ L, const uint32_t *__restrict__ n_arr, uint64_t stride,
                                                           const uint32_t *__restrict__ present, const uint64_t *__restrict__ lists,
                                                           int tiles_per_block, uint8_t *__restrict__ ranks)
{
    __shared__ uint4 s_list[MS_THREADS * ML_STRIDE4];
    __shared__ __align__(16) uint8_t s_stage[MS_THREADS * MS_STAGE_STRIDE];
    __shared__ uint8_t map[256];
    const uint32_t b = blockIdx.y;
    const uint32_t n = n_arr[b];
    const uint32_t tile0 = blockIdx.x * MS_THREADS;
    if ((uint64_t)tile0 * MTF_TILE >= n) return;
    build_symbol_map(present + b * 8, map);
    __syncthreads();
    const int lane = lane_id(), w = warp_id();
    const uint32_t wtile0 = tile0 + w * 32; // first tile of this warp
    if ((uint64_t)wtile0 * MTF_TILE >= n) return;
    const uint32_t tile = wtile0 + lane;
    const uint32_t base = tile * MTF_TILE;
    const uint32_t cnt = base < n ? min((uint32_t)MTF_TILE, n - base) : 0u;
    const uint4 *gl = reinterpret_cast<const uint4 *>(lists + ((uint64_t)b * tiles_per_block + tile) * 32);
    uint4 *list = s_list + threadIdx.x * ML_STRIDE4;
    Quad r = {0, 0, 0, 0};
    if (cnt) {
        uint4 v = gl[0];
        r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
        for (int k = 1; k < 16; ++k) list[k] = gl[k];
    }
    const uint8_t *src = L + (uint64_t)b * stride;
    uint8_t *dst = ranks + (uint64_t)b * stride;
    const uint32_t wcnt = min((uint32_t)MTF_TILE, n - wtile0 * MTF_TILE); // lane 0's tile is the fullest
    const uint32_t nchunks = (wcnt + MS_CHUNK - 1) / MS_CHUNK;
    uint8_t *wstage = s_stage + (w * 32) * MS_STAGE_STRIDE;
    uint8_t *mine = s_stage + threadIdx.x * MS_STAGE_STRIDE;
    for (uint32_t c = 0; c < nchunks; ++c) {
        // cooperative load: 4 lanes x 16 bytes cover one tile's 64-byte piece, 8 tiles per instruction; bytes are mapped to
        // dense symbol indices on the way in (keeps the table lookup out of the serial chain)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t tl = q * 8 + (lane >> 2), piece = lane & 3;
            uint32_t off = (wtile0 + tl) * MTF_TILE + c * MS_CHUNK + piece * 16;
            if (off < n) {
                uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + off));
                uint32_t in[4] = {v.x, v.y, v.z, v.w};
                uint32_t *sp = reinterpret_cast<uint32_t *>(wstage + tl * MS_STAGE_STRIDE + piece * 16);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t x = in[j];
                    sp[j] = (uint32_t)map[x & 0xff] | ((uint32_t)map[(x >> 8) & 0xff] << 8) | ((uint32_t)map[(x >> 16) & 0xff] << 16) |
                            ((uint32_t)map[x >> 24] << 24);
                }
            }
        }
        __syncwarp();
        int valid = (int)cnt - (int)(c * MS_CHUNK);
        valid = valid < 0 ? 0 : (valid > MS_CHUNK ? MS_CHUNK : valid);
        for (int i = 0; i < valid; ++i) {
            const uint32_t sym = mine[i];
            uint32_t rank = 0;
            if ((r.x & 0xff) != sym) {
                const uint32_t pat = sym * 0x01010101u;
                uint32_t p = quad_find(r, pat);
                Quad s = quad_shift(r, sym);
                if (p < 16) {
                    r = quad_blend(r, s, p);
                    rank = p;
                } else {
                    uint32_t carry = r.w >> 24;
                    r = s;
                    for (uint32_t k = 1;; ++k) { // the symbol is present: ends at k <= 15
                        uint4 v = list[k];
                        Quad q = {v.x, v.y, v.z, v.w};
                        p = quad_find(q, pat);
                        s = quad_shift(q, carry);
                        if (p < 16) {
                            q = quad_blend(q, s, p);
                            list[k] = make_uint4(q.x, q.y, q.z, q.w);
                            rank = 16 * k + p;
                            break;
                        }
                        list[k] = make_uint4(s.x, s.y, s.z, s.w);
                        carry = q.w >> 24;
                    }
                }
            }
            mine[i] = (uint8_t)rank;
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t tl = q * 8 + (lane >> 2), piece = lane & 3;
            uint32_t off = (wtile0 + tl) * MTF_TILE + c * MS_CHUNK + piece * 16;
            if (off < n) {
                const uint32_t *sp = reinterpret_cast<const uint32_t *>(wstage + tl * MS_STAGE_STRIDE + piece * 16);
                uint4 v = make_uint4(sp[0], sp[1], sp[2], sp[3]);
                *reinterpret_cast<uint4 *>(dst + off) = v;
            }
        }
        __syncwarp();
    }
}

// One warp per block: lists[t] (final list from identity, first dcount[t] entries = recency order) -> lists[t] = list at tile start.
__global__ void __launch_bounds__(MTF_WARPS * 32) k_mtf_compose(const uint32_t *__restrict__ n_arr, uint64_t *__restrict__ lists,
                                                                 const uint32_t *__restrict__ dcount, int tiles_per_block, uint32_t n_blocks)
{
    __shared__ uint8_t s_cur[MTF_WARPS][256], s_new[MTF_WARPS][256], s_mark[MTF_WARPS][256], s_f[MTF_WARPS][256];
    const uint32_t b = blockIdx.x * MTF_WARPS + warp_id();
    if (b >= n_blocks) return;
    const int lane = lane_id(), w = warp_id();
    const uint32_t n = n_arr[b];
    const uint32_t nt = (n + MTF_TILE - 1) / MTF_TILE;
    uint8_t *cur = s_cur[w], *nw = s_new[w], *mark = s_mark[w], *F = s_f[w];
    for (int i = lane; i < 256; i += 32) { cur[i] = (uint8_t)i; mark[i] = 0; }
    __syncwarp();
    for (uint32_t t = 0; t < nt; ++t) {
        const uint64_t lidx = ((uint64_t)b * tiles_per_block + t) * 32 + lane;
        uint64_t f = lists[lidx];
        const uint32_t d = dcount[(uint64_t)b * tiles_per_block + t];
        // publish the start list of tile t
        uint64_t c8 = *reinterpret_cast<const uint64_t *>(cur + lane * 8);
        lists[lidx] = c8;
        *reinterpret_cast<uint64_t *>(F + lane * 8) = f;
        __syncwarp();
        for (uint32_t i = lane; i < d; i += 32) { uint8_t s = F[i]; mark[s] = 1; nw[i] = s; }
        __syncwarp();
        // survivors of the old list keep their order behind the d recent symbols
        uint32_t keep = 0;
        uint8_t mine[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { mine[k] = (uint8_t)(c8 >> (8 * k)); keep += mark[mine[k]] ? 0u : 1u; }
        uint32_t off = d + warp_incl_sum(keep) - keep;
#pragma unroll
        for (int k = 0; k < 8; ++k) if (!mark[mine[k]]) nw[off++] = mine[k];
        __syncwarp();
        for (uint32_t i = lane; i < d; i += 32) mark[F[i]] = 0;
        uint8_t *tmp = cur; cur = nw; nw = tmp;
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------ RLE2
constexpr int R2_THREADS = 256;
constexpr int R2_ITEMS = 8;
constexpr int R2_TILE = R2_THREADS * R2_ITEMS; // 2048

__device__ __forceinline__ uint32_t run_digits(uint32_t r) { return r ? 31u - __clz(r + 1u) : 0u; } // bijective base-2 digits of r

struct R2Tile {
    uint32_t lead;   // zeros before the first non-zero rank (== len when all zero)
    uint32_t trail;  // zeros after the last non-zero rank
    uint32_t inner;  // symbols emitted by the non-zero ranks after the first one (run digits + the rank symbol)
    uint32_t allzero;
};

// per-element view of a tile: for a non-zero rank at a, zeros immediately before it inside the tile
__device__ __forceinline__ void r2_load(const uint8_t *rp, uint32_t a0, uint32_t cnt, uint32_t (&r)[R2_ITEMS], int &lastnz)
{
    lastnz = -1;
#pragma unroll
    for (int i = 0; i < R2_ITEMS; ++i) {
        uint32_t a = a0 + i;
        r[i] = a < cnt ? rp[a] : 0u;
        if (a < cnt && r[i]) lastnz = (int)a;
    }
}

__global__ void __launch_bounds__(R2_THREADS) k_rle2_tiles(const uint8_t *__restrict__ ranks, const uint32_t *__restrict__ n_arr, uint64_t stride,
                                                            R2Tile *__restrict__ tiles, int tiles_per_block)
{
    __shared__ int s_i[34];
    __shared__ uint32_t s_red[R2_THREADS / 32];
    __shared__ int s_first[R2_THREADS / 32];
    const uint32_t b = blockIdx.y, tile = blockIdx.x;
    const uint32_t n = n_arr[b];
    const uint32_t base = tile * R2_TILE;
    if (base >= n) return;
    const uint32_t cnt = min((uint32_t)R2_TILE, n - base); // tile-local indexing below
    const uint8_t *rp = ranks + (uint64_t)b * stride + base;
    const uint32_t a0 = threadIdx.x * R2_ITEMS;
    uint32_t r[R2_ITEMS];
    int lastnz;
    r2_load(rp, a0, cnt, r, lastnz);
    int total_last;
    int prev = block_excl_max(lastnz, s_i, &total_last);
    // first non-zero of the tile
    int firstnz = 0x7fffffff;
#pragma unroll
    for (int i = R2_ITEMS - 1; i >= 0; --i) if (a0 + i < cnt && r[i]) firstnz = (int)(a0 + i);
    firstnz = __reduce_min_sync(FULL, firstnz);
    uint32_t inner = 0;
    int p = prev;
#pragma unroll
    for (int i = 0; i < R2_ITEMS; ++i) {
        uint32_t a = a0 + i;
        if (a < cnt && r[i]) {
            if (p >= 0) inner += run_digits(a - 1 - (uint32_t)p) + 1; // not the first non-zero of the tile
            p = (int)a;
        }
    }
    inner = __reduce_add_sync(FULL, inner);
    if (lane_id() == 0) { s_red[warp_id()] = inner; s_first[warp_id()] = firstnz; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0; int f = 0x7fffffff;
        for (int w = 0; w < R2_THREADS / 32; ++w) { s += s_red[w]; f = min(f, s_first[w]); }
        R2Tile t;
        t.allzero = total_last < 0;
        t.lead = t.allzero ? cnt : (uint32_t)f;
        t.trail = t.allzero ? cnt : cnt - 1 - (uint32_t)total_last;
        t.inner = s;
        tiles[(uint64_t)b * tiles_per_block + tile] = t;
    }
}

// one thread per block: zero-run carry and output offset at every tile; final run flush + EOB; m = symbol count.
__global__ void __launch_bounds__(128) k_rle2_scan(const R2Tile *__restrict__ tiles, const uint32_t *__restrict__ n_arr, int tiles_per_block, uint32_t n_blocks,
                                                    uint32_t *__restrict__ tile_zin, uint32_t *__restrict__ tile_obase, const uint32_t *__restrict__ present,
                                                    uint16_t *__restrict__ mtf, uint64_t mtf_stride, uint32_t *__restrict__ freq, uint32_t *__restrict__ m_arr,
                                                    uint32_t *__restrict__ alpha_arr)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const uint32_t n = n_arr[b];
    const uint32_t nt = (n + R2_TILE - 1) / R2_TILE;
    uint32_t z = 0, o = 0;
    for (uint32_t t = 0; t < nt; ++t) {
        const uint64_t idx = (uint64_t)b * tiles_per_block + t;
        R2Tile T = tiles[idx];
        tile_zin[idx] = z;
        tile_obase[idx] = o;
        if (T.allzero) z += T.lead;
        else { o += run_digits(z + T.lead) + 1 + T.inner; z = T.trail; }
    }
    uint16_t *out = mtf + (uint64_t)b * mtf_stride;
    uint32_t *fq = freq + b * 258;
    if (z) { // trailing run (kernel.cpp:2616-2638)
        uint32_t x = z + 1, D = run_digits(z), na = 0, nbn = 0;
        for (uint32_t k = 0; k < D; ++k) { uint32_t s = (x >> k) & 1u; out[o++] = (uint16_t)s; if (s) ++nbn; else ++na; }
        if (na) atomicAdd(&fq[0], na);
        if (nbn) atomicAdd(&fq[1], nbn);
    }
    uint32_t uniq = 0;
    for (int w = 0; w < 8; ++w) uniq += __popc(present[b * 8 + w]);
    out[o++] = (uint16_t)(uniq + 1); // end-of-block symbol (kernel.cpp:2640-2641)
    atomicAdd(&fq[uniq + 1], 1u);
    m_arr[b] = o;
    alpha_arr[b] = uniq + 2;
}

__global__ void __launch_bounds__(R2_THREADS) k_rle2_emit(const uint8_t *__restrict__ ranks, const uint32_t *__restrict__ n_arr, uint64_t stride,
                                                           const uint32_t *__restrict__ tile_zin, const uint32_t *__restrict__ tile_obase, int tiles_per_block,
                                                           uint16_t *__restrict__ mtf, uint64_t mtf_stride, uint32_t *__restrict__ freq)
{
    __shared__ int s_i[34];
    __shared__ uint32_t s_u[34];
    __shared__ uint32_t s_hist[258];
    const uint32_t b = blockIdx.y, tile = blockIdx.x;
    const uint32_t n = n_arr[b];
    const uint32_t base = tile * R2_TILE;
    if (base >= n) return;
    for (int i = threadIdx.x; i < 258; i += R2_THREADS) s_hist[i] = 0;
    const uint32_t cnt = min((uint32_t)R2_TILE, n - base);
    const uint8_t *rp = ranks + (uint64_t)b * stride + base;
    const uint32_t a0 = threadIdx.x * R2_ITEMS;
    uint32_t r[R2_ITEMS];
    int lastnz;
    r2_load(rp, a0, cnt, r, lastnz);
    int prev = block_excl_max(lastnz, s_i, nullptr);
    const uint64_t tidx = (uint64_t)b * tiles_per_block + tile;
    const uint32_t zin = tile_zin[tidx];
    uint32_t emit = 0;
    int p = prev;
#pragma unroll
    for (int i = 0; i < R2_ITEMS; ++i) {
        uint32_t a = a0 + i;
        if (a < cnt && r[i]) {
            uint32_t run = p >= 0 ? a - 1 - (uint32_t)p : a + zin;
            emit += run_digits(run) + 1;
            p = (int)a;
        }
    }
    uint32_t off = tile_obase[tidx] + block_excl_sum(emit, s_u, nullptr); // also orders the s_hist zeroing
    uint16_t *out = mtf + (uint64_t)b * mtf_stride;
    uint32_t na = 0, nbn = 0;
    p = prev;
#pragma unroll
    for (int i = 0; i < R2_ITEMS; ++i) {
        uint32_t a = a0 + i;
        if (a < cnt && r[i]) {
            uint32_t run = p >= 0 ? a - 1 - (uint32_t)p : a + zin;
            uint32_t D = run_digits(run), x = run + 1;
            for (uint32_t k = 0; k < D; ++k) { uint32_t s = (x >> k) & 1u; out[off++] = (uint16_t)s; if (s) ++nbn; else ++na; }
            out[off++] = (uint16_t)(r[i] + 1);
            atomicAdd(&s_hist[r[i] + 1], 1u);
            p = (int)a;
        }
    }
    na = __reduce_add_sync(FULL, na);
    nbn = __reduce_add_sync(FULL, nbn);
    if (lane_id() == 0) { if (na) atomicAdd(&s_hist[0], na); if (nbn) atomicAdd(&s_hist[1], nbn); }
    __syncthreads();
    for (int i = threadIdx.x; i < 258; i += R2_THREADS) {
        uint32_t c = s_hist[i];
        if (c) atomicAdd(&freq[b * 258 + i], c);
    }
}

} // namespace bz
