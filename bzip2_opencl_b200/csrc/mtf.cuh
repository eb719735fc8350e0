// mtf.cuh -- stage C: move-to-front + zero-run (RUNA/RUNB) coding of the BWT last column.
// Reference: MTFAndRLE2StageEncoder + valueToFront, /root/reference/kernel.cpp:2513-2649 (one sequential loop per block).
//
// MTF is a serial dependency chain; the device formulation cuts every block into tiles of MTF_TILE symbols:
//   k_mtf_sim<false>  one THREAD per tile runs MTF from the identity list and keeps only the final list; its first d
//                     entries (d = distinct symbols seen) are the tile's symbols by recency -- a partial permutation
//   k_mtf_compose     one warp per block composes those in tile order -> the exact list at every tile start
//   k_mtf_sim<true>   one thread per tile re-runs MTF from its true start list and writes the ranks
// Each thread's 256-entry list is 32 64-bit words in shared memory (word 0 in a register); a lookup walks 8 entries per
// step with a SWAR zero-byte test and shifts the words it passes -- typical BWT output resolves in word 0.  Symbols and
// ranks are staged through shared memory so global traffic stays in 64-byte pieces.
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

// rank of c in (w0, list[1..31]); moves c to the front
__device__ __forceinline__ uint32_t mtf_lookup(uint64_t &w0, uint64_t *list, uint32_t c)
{
    if ((uint32_t)(w0 & 0xff) == c) return 0;
    const uint64_t pat = ONES8 * c;
    uint32_t bi = zero_byte_index(w0 ^ pat);
    if (bi < 8) {
        uint64_t m = bi == 7 ? ~0ull : ((1ull << (8 * (bi + 1))) - 1ull);
        w0 = (w0 & ~m) | (((w0 << 8) | c) & m);
        return bi;
    }
    uint32_t carry = (uint32_t)(w0 >> 56);
    w0 = (w0 << 8) | c;
    for (uint32_t k = 1;; ++k) { // c is always present: the loop ends at k <= 31
        uint64_t w = list[k];
        bi = zero_byte_index(w ^ pat);
        if (bi < 8) {
            uint64_t m = bi == 7 ? ~0ull : ((1ull << (8 * (bi + 1))) - 1ull);
            list[k] = (w & ~m) | (((w << 8) | carry) & m);
            return 8 * k + bi;
        }
        list[k] = (w << 8) | carry;
        carry = (uint32_t)(w >> 56);
    }
}

// One THREAD per tile.  WRITE=false: lists[tile] = final list from the identity start, dcount[tile] = distinct symbols seen.
// WRITE=true : lists[tile] is the true start list (k_mtf_compose); the ranks are written.
template <bool WRITE>
__global__ void __launch_bounds__(MS_THREADS) k_mtf_sim(const uint8_t *__restrict__ L, const uint32_t *__restrict__ n_arr, uint64_t stride,
                                                         const uint32_t *__restrict__ present, uint64_t *__restrict__ lists, uint32_t *__restrict__ dcount,
                                                         int tiles_per_block, uint8_t *__restrict__ ranks)
{
    __shared__ uint64_t s_list[MS_THREADS * MS_LIST_WORDS];
    __shared__ __align__(16) uint8_t s_stage[MS_THREADS * MS_STAGE_STRIDE];
    __shared__ uint8_t map[256];
    const uint32_t b = blockIdx.y;
    const uint32_t n = n_arr[b];
    const uint32_t tile0 = blockIdx.x * MS_THREADS;
    if ((uint64_t)tile0 * MTF_TILE >= n) return;
    build_symbol_map(present + b * 8, map);
    __syncthreads();
    const int lane = lane_id(), w = warp_id();
    const uint32_t wtile0 = tile0 + w * 32;          // first tile of this warp
    if ((uint64_t)wtile0 * MTF_TILE >= n) return;
    const uint32_t tile = wtile0 + lane;
    const uint32_t base = tile * MTF_TILE;
    const uint32_t cnt = base < n ? min((uint32_t)MTF_TILE, n - base) : 0u;
    const uint64_t lidx = ((uint64_t)b * tiles_per_block + tile) * 32;
    uint64_t *list = s_list + threadIdx.x * MS_LIST_WORDS;
    uint64_t w0 = 0;
    if (cnt) {
        if (WRITE) {
            w0 = lists[lidx];
            for (int k = 1; k < 32; ++k) list[k] = lists[lidx + k];
        } else {
            w0 = 0x0706050403020100ull;
            for (uint64_t k = 1; k < 32; ++k) list[k] = 0x0706050403020100ull + ONES8 * (8 * k);
        }
    }
    const uint8_t *src = L + (uint64_t)b * stride;
    uint8_t *dst = WRITE ? ranks + (uint64_t)b * stride : nullptr;
    const uint32_t wcnt = min((uint32_t)MTF_TILE, n - wtile0 * MTF_TILE); // symbols in the warp's fullest tile
    const uint32_t nchunks = (wcnt + MS_CHUNK - 1) / MS_CHUNK;
    uint8_t *wstage = s_stage + (w * 32) * MS_STAGE_STRIDE;
    uint8_t *mine = s_stage + threadIdx.x * MS_STAGE_STRIDE;
    uint32_t d = 0;
    for (uint32_t c = 0; c < nchunks; ++c) {
        // cooperative load: 4 lanes x 16 bytes cover one tile's 64-byte piece, 8 tiles per instruction
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t tl = q * 8 + (lane >> 2), piece = lane & 3;
            uint32_t off = (wtile0 + tl) * MTF_TILE + c * MS_CHUNK + piece * 16;
            if (off < n) {
                uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + off));
                uint32_t *sp = reinterpret_cast<uint32_t *>(wstage + tl * MS_STAGE_STRIDE + piece * 16);
                sp[0] = v.x; sp[1] = v.y; sp[2] = v.z; sp[3] = v.w;
            }
        }
        __syncwarp();
        int valid = (int)cnt - (int)(c * MS_CHUNK);
        valid = valid < 0 ? 0 : (valid > MS_CHUNK ? MS_CHUNK : valid);
        for (int i = 0; i < valid; ++i) {
            uint32_t r = mtf_lookup(w0, list, map[mine[i]]);
            if (WRITE) mine[i] = (uint8_t)r;
            else if (r >= d) ++d;
        }
        __syncwarp();
        if (WRITE) {
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
    if (!WRITE && cnt) {
        lists[lidx] = w0;
        for (int k = 1; k < 32; ++k) lists[lidx + k] = list[k];
        dcount[(uint64_t)b * tiles_per_block + tile] = d;
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
