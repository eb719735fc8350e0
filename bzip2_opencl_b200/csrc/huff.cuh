// huff.cuh -- stage D: the reference's multi-table Huffman stage (HuffmanStageEncoder, /root/reference/kernel.cpp:3064-3096)
// and stage E: bit emission at arbitrary bit offsets (writeSelectorsAndHuffmanTables :2991-3041, writeBlockData :3043-3062,
// block header bits OutputStream.hpp:208-211, origPtr :3116, symbol map :2483-2511).
//
//   k_huff_seeds     per block: table count (selectTableCount :2808-2818) and seed lengths (:2859-2893)
//   k_huff_pass      one thread per 50-symbol group: cost under each table (one packed 64-bit add per symbol), argmin with
//                    lowest index winning, selector, histogram into the winner's table (shared-memory privatised)   (:2895-2944)
//   k_huff_lengths   one warp per (block, table): bitonic sort of (freq<<9|sym) + the in-place Moffat-Katajainen allocator with
//                    the 20-bit limit, restated from :2652-2857 (strictly serial part on lane 0)
//   k_huff_codes     canonical codes (:2953-2989) + size of the delta-coded table
//   k_huff_selmtf    selector move-to-front positions (:3004-3012)
//   k_block_bits / k_bit_offsets   exact bit length of every block, exclusive scan -> global bit offset of every block
//   k_emit_header    per block: header, origPtr, symbol map, selectors (unary), tables (delta coded)
//   k_data_tilebits / k_data_scan / k_emit_data   Huffman-coded symbols: per-tile bit totals, scan, shared-memory bit assembly
#pragma once
#include "util.cuh"

namespace bz {

constexpr int HUFF_GROUP = 50;      // Config.hpp:34
constexpr int HUFF_MAXALPHA = 258;  // Config.hpp:35
constexpr int HUFF_HIGHCOST = 15;   // Config.hpp:36
constexpr int HUFF_MAXLEN = 20;     // Config.hpp:37
constexpr int HUFF_MAXTABLES = 6;   // Config.hpp:40
constexpr int HUFF_LEN_STRIDE = 264; // bytes per (block, table) row of code lengths
constexpr int HP_THREADS = 128;     // groups per CTA in k_huff_pass / k_emit_data
constexpr int HP_SYMS = HP_THREADS * HUFF_GROUP; // 6400 symbols per CTA

struct HuffBlock { // per block scalars
    uint32_t m;        // MTF/RLE2 symbols incl. EOB
    uint32_t alpha;    // alphabet size = unique + 2
    uint32_t ntab;     // 2..6
    uint32_t nsel;     // ceil(m / 50)
};

__global__ void __launch_bounds__(128) k_huff_seeds(const uint32_t *__restrict__ m_arr, const uint32_t *__restrict__ alpha_arr, const uint32_t *__restrict__ freq,
                                                     uint32_t n_blocks, HuffBlock *__restrict__ hb, uint8_t *__restrict__ lens)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const int m = (int)m_arr[b], alpha = (int)alpha_arr[b];
    const int T = m >= 2400 ? 6 : m >= 1200 ? 5 : m >= 600 ? 4 : m >= 200 ? 3 : 2;
    HuffBlock h;
    h.m = m; h.alpha = alpha; h.ntab = T; h.nsel = (m + HUFF_GROUP - 1) / HUFF_GROUP;
    hb[b] = h;
    const uint32_t *f = freq + b * 258;
    int remaining = m, hi = -1;
    for (int t = 0; t < T; ++t) {
        int target = remaining / (T - t);
        int lo = hi + 1, acc = 0;
        while (acc < target && hi < alpha - 1) acc += (int)f[++hi];
        if (hi > lo && t != 0 && t != T - 1 && ((T - t) % 2) == 0) acc -= (int)f[hi--];
        uint8_t *L = lens + ((uint64_t)b * HUFF_MAXTABLES + t) * HUFF_LEN_STRIDE;
        for (int j = 0; j < alpha; ++j) L[j] = (j < lo || j > hi) ? HUFF_HIGHCOST : 0;
        remaining -= acc;
    }
}

__global__ void __launch_bounds__(HP_THREADS) k_huff_pass(const uint16_t *__restrict__ mtf, uint64_t mtf_stride, const HuffBlock *__restrict__ hb,
                                                           const uint8_t *__restrict__ lens, uint32_t *__restrict__ tf, uint8_t *__restrict__ sel,
                                                           uint64_t sel_stride)
{
    __shared__ unsigned long long s_cost[HUFF_MAXALPHA];
    __shared__ uint32_t s_hist[HUFF_MAXTABLES][HUFF_MAXALPHA];
    __shared__ uint16_t s_sym[HP_SYMS];
    const uint32_t b = blockIdx.y;
    const HuffBlock h = hb[b];
    const uint32_t sym0 = blockIdx.x * HP_SYMS;
    if (sym0 >= h.m) return;
    const uint32_t nsym = min((uint32_t)HP_SYMS, h.m - sym0);
    const uint8_t *L = lens + (uint64_t)b * HUFF_MAXTABLES * HUFF_LEN_STRIDE;
    for (uint32_t j = threadIdx.x; j < h.alpha; j += HP_THREADS) {
        unsigned long long c = 0;
        for (uint32_t t = 0; t < h.ntab; ++t) c |= (unsigned long long)L[t * HUFF_LEN_STRIDE + j] << (10 * t);
        s_cost[j] = c;
    }
    for (uint32_t i = threadIdx.x; i < HUFF_MAXTABLES * HUFF_MAXALPHA; i += HP_THREADS) (&s_hist[0][0])[i] = 0;
    const uint16_t *src = mtf + (uint64_t)b * mtf_stride + sym0;
    for (uint32_t i = threadIdx.x; i < nsym; i += HP_THREADS) s_sym[i] = src[i];
    __syncthreads();
    const uint32_t g0 = threadIdx.x * HUFF_GROUP;
    if (g0 < nsym) {
        const uint32_t gc = min((uint32_t)HUFF_GROUP, nsym - g0);
        unsigned long long acc = 0;
        for (uint32_t k = 0; k < gc; ++k) acc += s_cost[s_sym[g0 + k]];
        uint32_t best = 0, bc = (uint32_t)(acc & 1023ull);
        for (uint32_t t = 1; t < h.ntab; ++t) {
            uint32_t c = (uint32_t)(acc >> (10 * t)) & 1023u;
            if (c < bc) { bc = c; best = t; }
        }
        for (uint32_t k = 0; k < gc; ++k) atomicAdd(&s_hist[best][s_sym[g0 + k]], 1u);
        sel[(uint64_t)b * sel_stride + blockIdx.x * HP_THREADS + threadIdx.x] = (uint8_t)best;
    }
    __syncthreads();
    uint32_t *gt = tf + (uint64_t)b * HUFF_MAXTABLES * HUFF_MAXALPHA;
    for (uint32_t i = threadIdx.x; i < h.ntab * HUFF_MAXALPHA; i += HP_THREADS) {
        uint32_t c = (&s_hist[0][0])[i];
        if (c) atomicAdd(&gt[i], c);
    }
}

// ---- in-place Moffat-Katajainen code-length allocator with length limit (restates kernel.cpp:2660-2805) ----
__device__ __forceinline__ int mk_first(const int *a, int len, int i, int nodesToMove)
{
    const int limit = i;
    int k = len - 2;
    while (i >= nodesToMove && (a[i] % len) > limit) { k = i; i -= (limit - i + 1); }
    if (i < nodesToMove - 1) i = nodesToMove - 1;
    while (k > i + 1) {
        int mid = (i + k) >> 1;
        if ((a[mid] % len) > limit) k = mid; else i = mid;
    }
    return k;
}
__device__ __forceinline__ void mk_allocate(int *a, int len, int maxLen)
{
    if (len == 2) { a[1] = 1; a[0] = 1; return; }
    if (len == 1) { a[0] = 1; return; }
    // setExtendedParentPointers (:2688-2718)
    a[0] += a[1];
    {
        int head = 0, top = 2;
        for (int tail = 1; tail < len - 1; ++tail) {
            int t;
            if (top >= len || a[head] < a[top]) { t = a[head]; a[head++] = tail; }
            else t = a[top++];
            if (top >= len || (head < tail && a[head] < a[top])) { t += a[head]; a[head++] = tail + len; }
            else t += a[top++];
            a[tail] = t;
        }
    }
    // findNodesToRelocate (:2720-2728)
    int reloc = len - 2;
    for (int depth = 1; depth < maxLen - 1 && reloc > 1; ++depth) reloc = mk_first(a, len, reloc - 1, 0);
    if ((a[0] % len) >= reloc) {
        // allocateNodeLengths (:2730-2746)
        int firstNode = len - 2, next = len - 1;
        for (int depth = 1, avail = 2; avail > 0; ++depth) {
            int last = firstNode;
            firstNode = mk_first(a, len, last - 1, 0);
            for (int i = avail - (last - firstNode); i > 0; --i) a[next--] = depth;
            avail = (last - firstNode) << 1;
        }
    } else {
        // allocateNodeLengthsWithRelocation (:2748-2782)
        int sig = 0;
        for (int x = reloc - 1; x > 0; x >>= 1) ++sig;
        const int insertDepth = maxLen - sig;
        int firstNode = len - 2, next = len - 1;
        int depth = insertDepth == 1 ? 2 : 1;
        int left = insertDepth == 1 ? reloc - 2 : reloc;
        for (int avail = depth << 1; avail > 0; ++depth) {
            int last = firstNode;
            firstNode = firstNode <= reloc ? firstNode : mk_first(a, len, last - 1, reloc);
            int off = 0;
            if (depth >= insertDepth) {
                int room = 1 << (depth - insertDepth);
                off = left < room ? left : room;
            } else if (depth == insertDepth - 1) {
                off = 1;
                if (a[firstNode] == last) firstNode++;
            }
            for (int i = avail - (last - firstNode + off); i > 0; --i) a[next--] = depth;
            left -= off;
            avail = (last - firstNode + off) << 1;
        }
    }
}

// one CTA of four warps per (block, table): the sort is shared by 128 threads, the allocator itself is serial (thread 0)
constexpr int HL_THREADS = 128;
__global__ void __launch_bounds__(HL_THREADS) k_huff_lengths(const HuffBlock *__restrict__ hb, uint32_t *__restrict__ tf, uint8_t *__restrict__ lens,
                                                              uint32_t *__restrict__ data_bits, int last_pass)
{
    __shared__ uint32_t key[512];
    __shared__ int arr[HUFF_MAXALPHA + 2];
    const uint32_t b = blockIdx.y, t = blockIdx.x;
    const HuffBlock h = hb[b];
    if (t >= h.ntab) return;
    const int tid = threadIdx.x;
    const int alpha = (int)h.alpha;
    // (sorting the next power of two >= alpha keys is enough: the padding keys are larger than every real one)
    int npad = 2;
    while (npad < alpha) npad <<= 1;
    uint32_t *f = tf + ((uint64_t)b * HUFF_MAXTABLES + t) * HUFF_MAXALPHA;
    for (int i = tid; i < npad; i += HL_THREADS) key[i] = i < alpha ? ((f[i] << 9) | (uint32_t)i) : 0xffffffffu;
    __syncthreads();
    // bitonic sort, ascending (keys are distinct: same order as sortArr, kernel.cpp:2820-2833)
    for (int k = 2; k <= npad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < npad; i += HL_THREADS) {
                int p = i ^ j;
                if (p > i) {
                    uint32_t x = key[i], y = key[p];
                    bool up = (i & k) == 0;
                    if ((x > y) == up) { key[i] = y; key[p] = x; }
                }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < alpha; i += HL_THREADS) arr[i] = (int)(key[i] >> 9);
    __syncthreads();
    if (tid == 0) mk_allocate(arr, alpha, HUFF_MAXLEN);
    __syncthreads();
    uint8_t *L = lens + ((uint64_t)b * HUFF_MAXTABLES + t) * HUFF_LEN_STRIDE;
    uint32_t bits = 0;
    for (int i = tid; i < alpha; i += HL_THREADS) {
        uint32_t s = key[i] & 0x1ffu;
        L[s] = (uint8_t)arr[i];
        bits += (key[i] >> 9) * (uint32_t)arr[i];
    }
    if (last_pass) {
        bits = __reduce_add_sync(FULL, bits);
        if (lane_id() == 0 && bits) atomicAdd(&data_bits[b], bits);
    } else {
        for (int i = tid; i < HUFF_MAXALPHA; i += HL_THREADS) f[i] = 0; // fresh frequencies for the next pass (patch P1)
    }
}

// one thread per (block, table): canonical codes, merged as (len << 24) | code; bits of the delta-coded table
__global__ void __launch_bounds__(64) k_huff_codes(const HuffBlock *__restrict__ hb, const uint8_t *__restrict__ lens, uint32_t *__restrict__ codes,
                                                    uint32_t *__restrict__ table_bits, uint32_t *__restrict__ per_table_bits, uint32_t n_blocks)
{
    const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t b = id / HUFF_MAXTABLES, t = id % HUFF_MAXTABLES;
    if (b >= n_blocks) return;
    const HuffBlock h = hb[b];
    if (t >= h.ntab) return;
    const uint8_t *L = lens + ((uint64_t)b * HUFF_MAXTABLES + t) * HUFF_LEN_STRIDE;
    uint32_t *C = codes + ((uint64_t)b * HUFF_MAXTABLES + t) * HUFF_MAXALPHA;
    const int alpha = (int)h.alpha;
    uint32_t count[HUFF_MAXLEN + 2];
    for (int l = 0; l <= HUFF_MAXLEN + 1; ++l) count[l] = 0;
    int lo = 32, hi = 0;
    uint32_t tb = 5;
    int cur = L[0];
    for (int j = 0; j < alpha; ++j) {
        int l = L[j];
        count[l]++;
        lo = min(lo, l); hi = max(hi, l);
        int d = l - cur;
        tb += 2u * (uint32_t)(d < 0 ? -d : d) + 1u;
        cur = l;
    }
    uint32_t next[HUFF_MAXLEN + 2];
    uint32_t code = 0;
    for (int l = lo; l <= hi; ++l) { next[l] = code; code = (code + count[l]) << 1; }
    for (int j = 0; j < alpha; ++j) {
        int l = L[j];
        C[j] = ((uint32_t)l << 24) | next[l]++;
    }
    atomicAdd(&table_bits[b], tb);
    per_table_bits[b * HUFF_MAXTABLES + t] = tb;
}

// One CTA per block: MTF positions of the selectors (kernel.cpp:3004-3012), the bit offset of every selector's unary code and
// their total length.  Position of selector g = number of OTHER tables used more recently than sel[g]; "last use of table t
// before g" is a prefix maximum, so every thread takes a contiguous chunk of selectors: (1) last use of every table inside the
// chunk, (2) exclusive prefix maximum over the chunks, (3) the chunk again with the true history -> positions and code lengths,
// (4) exclusive sum of the chunks' bit totals -> offsets.  The serial chain is one chunk (<= 71 selectors), not the block.
constexpr int SM_THREADS = 256;
constexpr int SM_MAXSEL = 18048; // >= HUFFMAN_MAXIMUM_SELECTORS (Config.hpp:41), a multiple of 64
__global__ void __launch_bounds__(SM_THREADS) k_huff_selmtf(const HuffBlock *__restrict__ hb, const uint8_t *__restrict__ sel, uint64_t sel_stride,
                                                             uint8_t *__restrict__ selpos, uint32_t *__restrict__ selbitoff, uint32_t *__restrict__ sel_bits,
                                                             uint32_t n_blocks)
{
    __shared__ uint8_t s_sel[SM_MAXSEL];
    __shared__ int s_last[SM_THREADS / 32][HUFF_MAXTABLES];
    __shared__ uint32_t s_u[34];
    const uint32_t b = blockIdx.x;
    if (b >= n_blocks) return;
    const uint32_t ns = min(hb[b].nsel, (uint32_t)SM_MAXSEL);
    const uint8_t *s = sel + (uint64_t)b * sel_stride;
    uint8_t *o = selpos + (uint64_t)b * sel_stride;
    uint32_t *bo = selbitoff + (uint64_t)b * sel_stride;
    for (uint32_t i = threadIdx.x * 16; i < ns; i += SM_THREADS * 16) // (slots are 64-byte multiples: whole 16-byte loads)
        *reinterpret_cast<uint4 *>(s_sel + i) = __ldg(reinterpret_cast<const uint4 *>(s + i));
    __syncthreads();
    const uint32_t ch = (ns + SM_THREADS - 1) / SM_THREADS;
    const uint32_t g0 = min(ns, threadIdx.x * ch), g1 = min(ns, g0 + ch);
    // (1) last use inside the chunk
    int last[HUFF_MAXTABLES];
#pragma unroll
    for (int t = 0; t < HUFF_MAXTABLES; ++t) last[t] = -1000;
    for (uint32_t g = g0; g < g1; ++g) {
        const int v = s_sel[g];
#pragma unroll
        for (int t = 0; t < HUFF_MAXTABLES; ++t) last[t] = v == t ? (int)g : last[t];
    }
    // (2) exclusive prefix maximum over threads; the initial list 0,1,..,5 == virtual uses at -1-t
    int before[HUFF_MAXTABLES];
    const int lane = lane_id(), w = warp_id();
#pragma unroll
    for (int t = 0; t < HUFF_MAXTABLES; ++t) {
        const int inc = warp_incl_max(last[t]);
        int ex = __shfl_up_sync(FULL, inc, 1);
        if (lane == 0) ex = -1000;
        before[t] = ex;
        if (lane == 31) s_last[w][t] = inc;
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < HUFF_MAXTABLES; ++t) {
        int carry = -1 - t;
        for (int ww = 0; ww < w; ++ww) carry = max(carry, s_last[ww][t]);
        before[t] = max(before[t], carry);
    }
    // (3) positions; the code of position p has p + 1 bits
    uint32_t bits = 0;
    for (uint32_t g = g0; g < g1; ++g) {
        const int v = s_sel[g];
        int mine = 0;
#pragma unroll
        for (int t = 0; t < HUFF_MAXTABLES; ++t) mine = v == t ? before[t] : mine;
        uint32_t pos = 0;
#pragma unroll
        for (int t = 0; t < HUFF_MAXTABLES; ++t) pos += (t != v && before[t] > mine) ? 1u : 0u;
#pragma unroll
        for (int t = 0; t < HUFF_MAXTABLES; ++t) before[t] = v == t ? (int)g : before[t];
        o[g] = (uint8_t)pos;
        s_sel[g] = (uint8_t)pos; // (only this thread reads its chunk)
        bits += pos + 1;
    }
    // (4) offsets
    uint32_t total;
    uint32_t off = block_excl_sum(bits, s_u, &total);
    for (uint32_t g = g0; g < g1; ++g) {
        bo[g] = off;
        off += (uint32_t)s_sel[g] + 1u;
    }
    if (threadIdx.x == 0) sel_bits[b] = total;
}

} // namespace bz
