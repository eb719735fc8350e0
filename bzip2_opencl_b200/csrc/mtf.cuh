// mtf.cuh -- stage C: move-to-front + zero-run (RUNA/RUNB) coding of the BWT last column.
// Reference: MTFAndRLE2StageEncoder + valueToFront, /root/reference/kernel.cpp:2513-2649 (one sequential loop per block).
//
// MTF is a serial dependency chain; the device formulation cuts every block into tiles of MTF_TILE symbols:
//   k_mtf_recency     one warp per tile: the tile's distinct symbols by recency (a partial permutation), computed in parallel from
//                     last-occurrence positions -- what running MTF over the tile would leave at the front of the list
//   k_mtf_compose     one warp per block composes those in tile order -> the exact list at every tile start
//   k_mtf_ranks       one WARP per tile re-runs MTF from its true start list, 32 symbols per step: ranks come from a
//                     branch-free dominance count over the step's previous-occurrence times, then the list is rebuilt in
//                     shared memory (recent symbols first, stable compaction of the rest).
//   k_rle2_scan / k_rle2_emit  (tile summaries come from k_mtf_ranks) turn ranks into the RUNA/RUNB/rank+1/EOB symbol stream with an
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

__device__ __forceinline__ uint32_t run_digits(uint32_t r) { return r ? 31u - __clz(r + 1u) : 0u; } // bijective base-2 digits of r

struct R2Tile {
    uint32_t lead;   // zeros before the first non-zero rank (== len when all zero)
    uint32_t trail;  // zeros after the last non-zero rank
    uint32_t inner;  // symbols emitted by the non-zero ranks after the first one (run digits + the rank symbol)
    uint32_t allzero;
};

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

// ---- pass 2: one WARP per tile re-runs MTF from its true start list (k_mtf_compose) and writes the ranks, 32 symbols per
// step, branch-free.  With tau_j = lane of the previous occurrence of lane j's symbol inside the step (or the virtual time
// -1 - v_j, v_j = position in the list at the start of the step, when there is none):
//     rank_j = [no previous occurrence] * v_j + #{ q : tau_j < q < j, tau_q < tau_j }
// (symbols that were behind s_j and occur before j jump in front of it; for a repeated symbol the count is the number of
// distinct symbols in between).  The list is then rebuilt in place: the step's distinct symbols by recency, followed by the
// untouched part of the old list in its old order (a stable compaction of the first max(v)+1 entries only).
constexpr int MW_WARPS = 8; // tiles per CTA

struct MtfWarpSmem {
    uint8_t stage[MTF_TILE]; // dense symbols in, ranks out
    uint8_t list[256];       // list[pos] = symbol
    uint8_t inv[256];        // inv[symbol] = pos
    uint8_t mark[256];       // symbol occurs in the current step
    alignas(16) uint16_t tau16[32]; // the step's previous-occurrence times + 256, lane p at halfword (p < 16 ? 2 (15 - p) : 2 (31 - p) + 1)
};

__global__ void __launch_bounds__(MW_WARPS * 32) k_mtf_ranks(const uint8_t *__restrict__ L, const uint32_t *__restrict__ n_arr, uint64_t stride,
                                                              const uint32_t *__restrict__ present, const uint64_t *__restrict__ lists,
                                                              int tiles_per_block, uint8_t *__restrict__ ranks, R2Tile *__restrict__ r2tiles)
{
    __shared__ __align__(16) MtfWarpSmem sm_all[MW_WARPS];
    __shared__ uint8_t map[256];
    const uint32_t b = blockIdx.y;
    const uint32_t n = n_arr[b];
    if ((uint64_t)blockIdx.x * MW_WARPS * MTF_TILE >= n) return;
    build_symbol_map(present + b * 8, map);
    __syncthreads();
    const int lane = lane_id(), w = warp_id();
    const uint32_t tile = blockIdx.x * MW_WARPS + w;
    const uint32_t base = tile * MTF_TILE;
    if (base >= n) return;
    const uint32_t cnt = min((uint32_t)MTF_TILE, n - base);
    MtfWarpSmem &S = sm_all[w];
    {
        const uint64_t l8 = lists[((uint64_t)b * tiles_per_block + tile) * 32 + lane];
        *reinterpret_cast<uint64_t *>(S.list + lane * 8) = l8;
        *reinterpret_cast<uint64_t *>(S.mark + lane * 8) = 0ull;
#pragma unroll
        for (int k = 0; k < 8; ++k) S.inv[(uint32_t)(l8 >> (8 * k)) & 0xffu] = (uint8_t)(lane * 8 + k);
    }
    const uint8_t *src = L + (uint64_t)b * stride + base;
#pragma unroll
    for (int q = 0; q < MTF_TILE / 512; ++q) {
        const uint32_t off = (q * 32 + lane) * 16;
        if (off < cnt) { // bytes past n inside the last 16 are in-bounds garbage and never used
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + off));
            const uint32_t in[4] = {v.x, v.y, v.z, v.w};
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t x = in[j];
                o[j] = (uint32_t)map[x & 0xff] | ((uint32_t)map[(x >> 8) & 0xff] << 8) | ((uint32_t)map[(x >> 16) & 0xff] << 16) | ((uint32_t)map[x >> 24] << 24);
            }
            *reinterpret_cast<uint4 *>(S.stage + off) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
    __syncwarp();
    const unsigned lt = (1u << lane) - 1u, gt = 0xfffffffeu << lane;
    // zero-run summary of the tile for the RLE2 stage (what k_rle2_scan composes): first / last non-zero rank, symbols
    // emitted by the non-zero ranks after the first one
    int firstnz = -1, lastnz = -1;
    uint32_t inner = 0;
    for (uint32_t c0 = 0; c0 < cnt; c0 += 32) {
        const bool valid = c0 + lane < cnt;
        const uint32_t sym = S.stage[c0 + lane];
        const unsigned peers = __match_any_sync(FULL, valid ? sym : 256u + lane);
        const unsigned lower = peers & lt;
        const bool first = lower == 0;
        const uint32_t v = S.inv[valid ? sym : 0u];
        const int tau = first ? -1 - (int)v : 31 - __clz((int)lower);
        const unsigned W = tau >= 0 ? (lt & ~((2u << tau) - 1u)) : lt;
        // acc bit q = (tau_q < tau_j).  Every lane reads all 32 times with four broadcast loads; two 16-bit times per word are
        // compared at once (x | 0x8000 - tau_j keeps bit 15 iff x >= tau_j, no borrow between the halves), and the halfword
        // order in shared memory makes word k's two result bits land on bits 15 - k and 31 - k, i.e. on their lanes.
        S.tau16[lane < 16 ? 2 * (15 - lane) : 2 * (31 - lane) + 1] = (uint16_t)(tau + 256);
        __syncwarp();
        unsigned acc = 0;
        {
            const uint32_t tj2 = (uint32_t)(tau + 256) * 0x00010001u;
            const uint4 *tp = reinterpret_cast<const uint4 *>(S.tau16);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const uint4 t4 = tp[g];
                const uint32_t ws[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) acc |= (~((ws[k] | 0x80008000u) - tj2) & 0x80008000u) >> (4 * g + k);
            }
        }
        const uint32_t rank = __popc(acc & W) + (first ? v : 0u);
        if (valid) S.stage[c0 + lane] = (uint8_t)rank;
        {
            const unsigned nz = __ballot_sync(FULL, valid && rank != 0);
            if (valid && rank != 0) {
                const unsigned below = nz & lt;
                const int p = below ? (int)c0 + 31 - __clz((int)below) : lastnz;
                if (p >= 0) inner += run_digits(c0 + lane - 1 - (uint32_t)p) + 1;
            }
            if (nz) {
                if (firstnz < 0) firstnz = (int)c0 + __ffs((int)nz) - 1;
                lastnz = (int)c0 + 31 - __clz((int)nz);
            }
        }
        // ---- rebuild the list
        const bool is_last = valid && (peers & gt) == 0;
        const unsigned Bl = __ballot_sync(FULL, is_last);
        const int maxv = __reduce_max_sync(FULL, (valid && first) ? (int)v : -1);
        if (is_last) S.mark[sym] = 1;
        __syncwarp();
        uint32_t suffix = 0; // marked entries at higher positions than the current 32-entry slice
        for (int k = maxv >> 5; k >= 0; --k) {
            const int x = k * 32 + lane;
            const uint32_t e = S.list[x];
            const bool in = x <= maxv;
            const bool mk = in && S.mark[e];
            const unsigned Bm = __ballot_sync(FULL, mk);
            const uint32_t np = (uint32_t)x + suffix + __popc(Bm & gt); // old position + marked entries behind it
            if (in && !mk) { S.list[np] = (uint8_t)e; S.inv[e] = (uint8_t)np; }
            suffix += __popc(Bm);
            __syncwarp();
        }
        if (is_last) {
            const uint32_t r = __popc(Bl & gt);
            S.list[r] = (uint8_t)sym;
            S.inv[sym] = (uint8_t)r;
            S.mark[sym] = 0;
        }
        __syncwarp();
    }
    uint8_t *dst = ranks + (uint64_t)b * stride + base;
#pragma unroll
    for (int q = 0; q < MTF_TILE / 512; ++q) {
        const uint32_t off = (q * 32 + lane) * 16;
        if (off < cnt) *reinterpret_cast<uint4 *>(dst + off) = *reinterpret_cast<const uint4 *>(S.stage + off);
    }
    inner = __reduce_add_sync(FULL, inner);
    if (lane == 0) {
        R2Tile t;
        t.allzero = lastnz < 0;
        t.lead = t.allzero ? cnt : (uint32_t)firstnz;
        t.trail = t.allzero ? cnt : cnt - 1 - (uint32_t)lastnz;
        t.inner = inner;
        r2tiles[(uint64_t)b * tiles_per_block + tile] = t;
    }
}

// One CTA per block: lists[t] (first dcount[t] entries = the tile's distinct symbols by recency) -> lists[t] = list at tile start.
// "Move the tile's symbols to the front in recency order" composes associatively -- a range of tiles acts like one tile whose
// recency order is its tiles' orders merged, newest first -- so the chain over the block's tiles is cut into MC_WARPS chunks:
// every warp (1) merges its chunk's orders into one partial list, (2) warp 0 pushes the identity list through the chunk
// summaries, (3) every warp walks its chunk again from its true start list and publishes the list in front of every tile.
// The serial chain is 2 * tiles / MC_WARPS + MC_WARPS steps instead of `tiles`.
constexpr int MC_WARPS = 8;

// cur[0 .. len) <- F[0 .. d) followed by the entries of cur that are not in F, order kept; returns the new length.  One warp;
// mark[] is all zero on entry and on exit.
__device__ __forceinline__ uint32_t mtf_compose_step(uint8_t *&cur, uint8_t *&nw, uint8_t *mark, const uint8_t *F, uint32_t d, uint32_t len)
{
    const int lane = lane_id();
    for (uint32_t i = lane; i < d; i += 32) { const uint8_t s = F[i]; mark[s] = 1; nw[i] = s; }
    __syncwarp();
    // survivors of the old list keep their order behind the d recent symbols
    uint32_t keep = 0;
    uint8_t mine[8];
    const uint64_t c8 = *reinterpret_cast<const uint64_t *>(cur + lane * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) { mine[k] = (uint8_t)(c8 >> (8 * k)); keep += ((uint32_t)(lane * 8 + k) < len && !mark[mine[k]]) ? 1u : 0u; }
    const uint32_t inc = warp_incl_sum(keep);
    uint32_t off = d + inc - keep;
#pragma unroll
    for (int k = 0; k < 8; ++k) if ((uint32_t)(lane * 8 + k) < len && !mark[mine[k]]) nw[off++] = mine[k];
    const uint32_t newlen = d + __shfl_sync(FULL, inc, 31);
    __syncwarp();
    for (uint32_t i = lane; i < d; i += 32) mark[F[i]] = 0;
    uint8_t *tmp = cur; cur = nw; nw = tmp;
    __syncwarp();
    return newlen;
}

__global__ void __launch_bounds__(MC_WARPS * 32) k_mtf_compose(const uint32_t *__restrict__ n_arr, uint64_t *__restrict__ lists,
                                                                const uint32_t *__restrict__ dcount, int tiles_per_block, uint32_t n_blocks)
{
    __shared__ __align__(8) uint8_t s_a[MC_WARPS][256], s_b[MC_WARPS][256], s_mark[MC_WARPS][256], s_f[MC_WARPS][256];
    __shared__ __align__(8) uint8_t s_sum[MC_WARPS][256];  // chunk summaries, then the list in front of every chunk
    __shared__ uint32_t s_len[MC_WARPS];
    const uint32_t b = blockIdx.x;
    if (b >= n_blocks) return;
    const int lane = lane_id(), w = warp_id();
    const uint32_t n = n_arr[b];
    const uint32_t nt = (n + MTF_TILE - 1) / MTF_TILE;
    const uint32_t per = (nt + MC_WARPS - 1) / MC_WARPS;
    const uint32_t t0 = min(nt, (uint32_t)w * per), t1 = min(nt, t0 + per);
    uint8_t *cur = s_a[w], *nw = s_b[w], *mark = s_mark[w], *F = s_f[w];
    for (int i = lane; i < 256; i += 32) mark[i] = 0;
    __syncwarp();
    const uint64_t base = (uint64_t)b * tiles_per_block;
    // (1) the chunk's summary: its tiles' recency orders merged, starting from an empty list
    {
        uint32_t len = 0;
        uint64_t f_next = t0 < t1 ? lists[(base + t0) * 32 + lane] : 0ull;
        uint32_t d_next = t0 < t1 ? dcount[base + t0] : 0u;
        for (uint32_t t = t0; t < t1; ++t) {
            const uint64_t f = f_next;
            const uint32_t d = d_next;
            if (t + 1 < t1) { f_next = lists[(base + t + 1) * 32 + lane]; d_next = dcount[base + t + 1]; }
            *reinterpret_cast<uint64_t *>(F + lane * 8) = f;
            __syncwarp();
            len = mtf_compose_step(cur, nw, mark, F, d, len);
        }
        *reinterpret_cast<uint64_t *>(s_sum[w] + lane * 8) = *reinterpret_cast<const uint64_t *>(cur + lane * 8);
        if (lane == 0) s_len[w] = len;
    }
    __syncthreads();
    // (2) warp 0: the identity list pushed through the chunk summaries; s_sum[c] becomes the list in front of chunk c
    if (w == 0) {
        for (int i = lane; i < 256; i += 32) cur[i] = (uint8_t)i;
        __syncwarp();
        for (int c = 0; c < MC_WARPS; ++c) {
            const uint64_t sum8 = *reinterpret_cast<const uint64_t *>(s_sum[c] + lane * 8);
            const uint32_t d = s_len[c];
            *reinterpret_cast<uint64_t *>(s_sum[c] + lane * 8) = *reinterpret_cast<const uint64_t *>(cur + lane * 8);
            *reinterpret_cast<uint64_t *>(F + lane * 8) = sum8;
            __syncwarp();
            (void)mtf_compose_step(cur, nw, mark, F, d, 256u);
        }
    }
    __syncthreads();
    // (3) the chunk again from its true start list: publish the list in front of every tile
    {
        *reinterpret_cast<uint64_t *>(cur + lane * 8) = *reinterpret_cast<const uint64_t *>(s_sum[w] + lane * 8);
        __syncwarp();
        uint64_t f_next = t0 < t1 ? lists[(base + t0) * 32 + lane] : 0ull;
        uint32_t d_next = t0 < t1 ? dcount[base + t0] : 0u;
        for (uint32_t t = t0; t < t1; ++t) {
            const uint64_t lidx = (base + t) * 32 + lane;
            const uint64_t f = f_next;
            const uint32_t d = d_next;
            if (t + 1 < t1) { f_next = lists[lidx + 32]; d_next = dcount[base + t + 1]; }
            lists[lidx] = *reinterpret_cast<const uint64_t *>(cur + lane * 8); // the start list of tile t
            *reinterpret_cast<uint64_t *>(F + lane * 8) = f;
            __syncwarp();
            (void)mtf_compose_step(cur, nw, mark, F, d, 256u);
        }
    }
}

// ------------------------------------------------------------------------------------------------ RLE2
constexpr int R2_THREADS = 256;
constexpr int R2_ITEMS = 8;
constexpr int R2_TILE = R2_THREADS * R2_ITEMS; // 2048
static_assert(R2_TILE == MTF_TILE, "k_mtf_ranks writes the RLE2 tile summaries");



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
    __shared__ uint16_t s_out[R2_TILE + 64]; // the tile's symbols (a run of z zeros gives fewer than z digits; the carried-in run at most 20)
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
    uint32_t total;
    uint32_t off = block_excl_sum(emit, s_u, &total); // also orders the s_hist zeroing
    uint16_t *out = s_out; // assembled in shared memory, stored to the block's symbol stream in one coalesced sweep
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
    {
        uint16_t *dst = mtf + (uint64_t)b * mtf_stride + tile_obase[tidx];
        for (uint32_t i = threadIdx.x; i < total; i += R2_THREADS) dst[i] = s_out[i];
    }
    for (int i = threadIdx.x; i < 258; i += R2_THREADS) {
        uint32_t c = s_hist[i];
        if (c) atomicAdd(&freq[b * 258 + i], c);
    }
}

} // namespace bz
