// util.cuh -- shared device helpers (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bz {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src)
{
    uint32_t lo = __shfl_sync(FULL, (uint32_t)v, src), hi = __shfl_sync(FULL, (uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_up_u64(uint64_t v, int d)
{
    uint32_t lo = __shfl_up_sync(FULL, (uint32_t)v, d), hi = __shfl_up_sync(FULL, (uint32_t)(v >> 32), d);
    return ((uint64_t)hi << 32) | lo;
}

// Lanes of the warp that hold the same 8-bit digit as this lane (invalid lanes match nobody): eight ballots instead of
// match.any, whose latency grows with the number of distinct values in the warp.
__device__ __forceinline__ unsigned warp_digit_peers(uint32_t d, bool valid)
{
    unsigned peers = __ballot_sync(FULL, valid);
#pragma unroll
    for (int bit = 0; bit < 8; ++bit) {
        const bool one = (d >> bit) & 1u;
        const unsigned b = __ballot_sync(FULL, one);
        peers &= one ? b : ~b;
    }
    return valid ? peers : 0u;
}

// inclusive warp scans
__device__ __forceinline__ uint32_t warp_incl_sum(uint32_t v)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(FULL, v, d);
        if (lane_id() >= d) v += o;
    }
    return v;
}
__device__ __forceinline__ int warp_incl_max(int v)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int o = __shfl_up_sync(FULL, v, d);
        if (lane_id() >= d) v = max(v, o);
    }
    return v;
}
__device__ __forceinline__ uint64_t warp_incl_sum64(uint64_t v)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint64_t o = shfl_up_u64(v, d);
        if (lane_id() >= d) v += o;
    }
    return v;
}

// Block-wide exclusive sum over one value per thread (blockDim.x <= 1024, multiple of 32).
// `scratch` needs 33 uint32_t.  Returns exclusive prefix; *total gets the block total.
__device__ __forceinline__ uint32_t block_excl_sum(uint32_t v, uint32_t *scratch, uint32_t *total)
{
    const int lane = lane_id(), w = warp_id(), nw = blockDim.x >> 5;
    uint32_t inc = warp_incl_sum(v);
    if (lane == 31) scratch[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t x = lane < nw ? scratch[lane] : 0;
        uint32_t xi = warp_incl_sum(x);
        scratch[lane] = xi - x;
        if (lane == 31) scratch[32] = xi;
    }
    __syncthreads();
    uint32_t res = scratch[w] + inc - v;
    if (total) *total = scratch[32];
    __syncthreads();
    return res;
}

// Block-wide inclusive max over one value per thread, returns the exclusive max of all PREVIOUS threads (-1 if none).
__device__ __forceinline__ int block_excl_max(int v, int *scratch, int *total)
{
    const int lane = lane_id(), w = warp_id(), nw = blockDim.x >> 5;
    int inc = warp_incl_max(v);
    if (lane == 31) scratch[w] = inc;
    __syncthreads();
    if (w == 0) {
        int x = lane < nw ? scratch[lane] : -1;
        int xi = warp_incl_max(x);
        int ex = __shfl_up_sync(FULL, xi, 1);
        if (lane == 0) ex = -1;
        scratch[lane] = ex;
        if (lane == 31) scratch[32] = xi;
    }
    __syncthreads();
    int prev = __shfl_up_sync(FULL, inc, 1);
    if (lane == 0) prev = -1;
    int res = max(scratch[w], prev);
    if (total) *total = scratch[32];
    __syncthreads();
    return res;
}

// Exclusive SUFFIX minimum over one value per thread: min of the values of all threads with a HIGHER index (INT_MAX if none).
__device__ __forceinline__ int block_excl_min_suffix(int v, int *scratch)
{
    const int lane = lane_id(), w = warp_id(), nw = blockDim.x >> 5;
    int inc = v; // min over lanes >= lane of this warp
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int o = __shfl_down_sync(FULL, inc, d);
        if (lane + d < 32) inc = min(inc, o);
    }
    if (lane == 0) scratch[w] = inc;
    __syncthreads();
    if (w == 0) {
        int x = lane < nw ? scratch[lane] : 0x7fffffff;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int o = __shfl_down_sync(FULL, x, d);
            if (lane + d < 32) x = min(x, o);
        }
        int ex = __shfl_down_sync(FULL, x, 1);
        if (lane == 31) ex = 0x7fffffff;
        scratch[lane] = ex; // min over warps > lane
    }
    __syncthreads();
    int nxt = __shfl_down_sync(FULL, inc, 1);
    if (lane == 31) nxt = 0x7fffffff;
    int res = min(scratch[w], nxt);
    __syncthreads();
    return res;
}

// Three block-wide scans in one set of barriers: exclusive prefix maxima of a and b (-1 if no earlier thread) and the
// exclusive suffix minimum of c (INT_MAX if no later thread).  `scratch` needs 3 * 32 ints; blockDim.x <= 1024, multiple of 32.
__device__ __forceinline__ void block_scan_max2_minsuffix(int a, int b, int c, int *scratch, int &prev_a, int &prev_b, int &next_c)
{
    const int lane = lane_id(), w = warp_id(), nw = blockDim.x >> 5;
    int ia = a, ib = b, ic = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int oa = __shfl_up_sync(FULL, ia, d), ob = __shfl_up_sync(FULL, ib, d), oc = __shfl_down_sync(FULL, ic, d);
        if (lane >= d) { ia = max(ia, oa); ib = max(ib, ob); }
        if (lane + d < 32) ic = min(ic, oc);
    }
    if (lane == 31) { scratch[w] = ia; scratch[32 + w] = ib; }
    if (lane == 0) scratch[64 + w] = ic;
    __syncthreads();
    if (w == 0) { // exclusive scans over the warp totals
        int xa = lane < nw ? scratch[lane] : -1, xb = lane < nw ? scratch[32 + lane] : -1, xc = lane < nw ? scratch[64 + lane] : 0x7fffffff;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int oa = __shfl_up_sync(FULL, xa, d), ob = __shfl_up_sync(FULL, xb, d), oc = __shfl_down_sync(FULL, xc, d);
            if (lane >= d) { xa = max(xa, oa); xb = max(xb, ob); }
            if (lane + d < 32) xc = min(xc, oc);
        }
        int ea = __shfl_up_sync(FULL, xa, 1), eb = __shfl_up_sync(FULL, xb, 1), ec = __shfl_down_sync(FULL, xc, 1);
        if (lane == 0) { ea = -1; eb = -1; }
        if (lane == 31) ec = 0x7fffffff;
        scratch[lane] = ea; scratch[32 + lane] = eb; scratch[64 + lane] = ec;
    }
    __syncthreads();
    const int wa = scratch[w], wb = scratch[32 + w], wc = scratch[64 + w];
    int pa = __shfl_up_sync(FULL, ia, 1), pb = __shfl_up_sync(FULL, ib, 1), nc = __shfl_down_sync(FULL, ic, 1);
    if (lane == 0) { pa = -1; pb = -1; }
    if (lane == 31) nc = 0x7fffffff;
    prev_a = max(wa, pa);
    prev_b = max(wb, pb);
    next_c = min(wc, nc);
    __syncthreads();
}

} // namespace bz
