// emit.cuh -- stage E: bit emission.  Bits are MSB-first (BitOutputStream.hpp:101-113).  The output buffer is an array of
// bytes in stream order; kernels build big-endian 32-bit words and OR their byte-swapped image into a zero-initialised
// buffer, so neighbouring producers (header thread, data tiles, the next block) may share a word without coordination.
// Every block is written directly at its final bit offset in the stream (exclusive scan of exact block bit lengths), which
// replaces the reference's bool-per-bit buffers and the host's writeFileBytes/getLeftBuffer carry loop
// (OutputStream.hpp:225-239, BitOutputStream.hpp:30-99).
#pragma once
#include "util.cuh"
#include "huff.cuh"

namespace bz {

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// Serial MSB-first bit writer used by one thread; flushes with atomicOr.
struct BitWriter {
    uint32_t *out;   // word array of the stream buffer
    uint64_t word;   // index of the word being filled
    uint32_t acc;    // bits collected for that word (big-endian image), already positioned
    uint32_t used;   // bits used in acc (0..31)
    __device__ void init(uint32_t *o, uint64_t bitpos) { out = o; word = bitpos >> 5; used = (uint32_t)(bitpos & 31); acc = 0; }
    __device__ void put(uint32_t count, uint32_t value) // count 1..32, value < 2^count
    {
        uint32_t room = 32 - used;
        if (count <= room) {
            acc |= (count == 32 ? value : (value << (room - count)));
            used += count;
            if (used == 32) flush_word();
        } else {
            uint32_t hi = value >> (count - room);
            acc |= hi;
            used = 32;
            flush_word();
            uint32_t rest = count - room;
            acc = (value & ((1u << rest) - 1u)) << (32 - rest);
            used = rest;
        }
    }
    __device__ void flush_word()
    {
        if (acc) atomicOr(out + word, bswap32(acc));
        ++word; acc = 0; used = 0;
    }
    __device__ void finish() { if (used && acc) atomicOr(out + word, bswap32(acc)); }
};

// exact bit length of every block
__global__ void __launch_bounds__(128) k_block_bits(const HuffBlock *__restrict__ hb, const uint32_t *__restrict__ present, const uint32_t *__restrict__ sel_bits,
                                                     const uint32_t *__restrict__ table_bits, const uint32_t *__restrict__ data_bits, uint32_t n_blocks,
                                                     uint64_t *__restrict__ nbits)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    uint32_t ranges = 0;
    for (int w = 0; w < 8; ++w) {
        uint32_t p = present[b * 8 + w];
        if (p & 0xffffu) ++ranges;
        if (p >> 16) ++ranges;
    }
    uint64_t bits = 48 + 32 + 1 + 24 + 16 + 16ull * ranges + 3 + 15;
    bits += sel_bits[b];
    bits += table_bits[b];
    bits += data_bits[b];
    nbits[b] = bits;
}

// single CTA: bitoff[b] = *bitpos + exclusive sum; *bitpos advances by the batch total (so batches chain without a host sync)
// A block whose bits would end behind cap_bits (the output buffer) gets the offset BIT_NOWHERE and is not written at all: the
// position still advances, so the host sees the overrun in the final bit count before anything was stored out of bounds.
constexpr uint64_t BIT_NOWHERE = ~0ull;
__global__ void __launch_bounds__(1024) k_bit_offsets(const uint64_t *__restrict__ nbits, uint32_t n_blocks, uint64_t *__restrict__ bitpos,
                                                       uint64_t *__restrict__ bitoff, uint64_t *__restrict__ batch_end, uint64_t cap_bits)
{
    __shared__ uint64_t s_w[33];
    __shared__ uint64_t s_carry;
    if (threadIdx.x == 0) s_carry = *bitpos;
    __syncthreads();
    for (uint32_t b0 = 0; b0 < n_blocks; b0 += blockDim.x) {
        uint32_t b = b0 + threadIdx.x;
        uint64_t v = b < n_blocks ? nbits[b] : 0;
        uint64_t inc = warp_incl_sum64(v);
        if (lane_id() == 31) s_w[warp_id()] = inc;
        __syncthreads();
        if (warp_id() == 0) {
            uint64_t x = s_w[lane_id()];
            uint64_t xi = warp_incl_sum64(x);
            s_w[lane_id()] = xi - x;
            if (lane_id() == 31) s_w[32] = xi;
        }
        __syncthreads();
        if (b < n_blocks) {
            const uint64_t at = s_carry + s_w[warp_id()] + inc - v;
            bitoff[b] = at + v <= cap_bits ? at : BIT_NOWHERE;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_carry += s_w[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) { *bitpos = s_carry; *batch_end = s_carry; } // batch_end: this batch's own copy (bitpos moves on with the next batch)
}

// OR `len` (1..32) bits of `value` into the stream at absolute bit position `pos`
__device__ __forceinline__ void or_bits(uint32_t *out, uint64_t pos, uint32_t len, uint32_t value)
{
    const uint64_t w = pos >> 5;
    const uint32_t off = (uint32_t)(pos & 31);
    if (off + len <= 32) {
        uint32_t x = value << (32 - off - len);
        if (x) atomicOr(out + w, bswap32(x));
    } else {
        const uint32_t rest = off + len - 32;
        uint32_t hi = value >> rest, lo = value << (32 - rest);
        if (hi) atomicOr(out + w, bswap32(hi));
        if (lo) atomicOr(out + w + 1, bswap32(lo));
    }
}

// One CTA per block: everything in front of the coded symbols.  Thread 0 writes the fixed fields, all threads write the selector
// unary codes at their precomputed offsets, warp t writes table t: the bit offset of every symbol's delta code (1 + 2 |difference to
// the previous length| bits, kernel.cpp:3015-3040) is a prefix sum, so the 258 symbols of a table are written in parallel.
constexpr int EH_THREADS = 256;
__global__ void __launch_bounds__(EH_THREADS) k_emit_header(const HuffBlock *__restrict__ hb, const uint32_t *__restrict__ present, const uint32_t *__restrict__ crc,
                                                             const uint32_t *__restrict__ orig, const uint8_t *__restrict__ selpos,
                                                             const uint32_t *__restrict__ selbitoff, uint64_t sel_stride, const uint32_t *__restrict__ sel_bits,
                                                             const uint32_t *__restrict__ per_table_bits, const uint8_t *__restrict__ lens,
                                                             const uint64_t *__restrict__ bitoff, uint32_t n_blocks, uint32_t *__restrict__ out,
                                                             uint64_t *__restrict__ data_start)
{
    const uint32_t b = blockIdx.x;
    if (b >= n_blocks) return;
    const int lane = lane_id(), w = warp_id();
    if (bitoff[b] == BIT_NOWHERE) { // does not fit the output buffer (k_bit_offsets)
        if (threadIdx.x == 0) data_start[b] = BIT_NOWHERE;
        return;
    }
    const HuffBlock h = hb[b];
    uint32_t ranges = 0;
    for (int k = 0; k < 8; ++k) {
        uint32_t p = present[b * 8 + k];
        ranges += (p & 0xffffu) ? 1u : 0u;
        ranges += (p >> 16) ? 1u : 0u;
    }
    const uint64_t sel_start = bitoff[b] + 48 + 32 + 1 + 24 + 16 + 16ull * ranges + 3 + 15;
    if (threadIdx.x == 0) {
        BitWriter bw;
        bw.init(out, bitoff[b]);
        bw.put(24, 0x314159u); bw.put(24, 0x265359u);       // OutputStream.hpp:208-209
        bw.put(32, crc[b]);                                   // :210 writeInteger
        bw.put(1, 0);                                         // :211 randomised = 0
        bw.put(24, orig[b]);                                  // kernel.cpp:3116
        // writeSymbolMap, kernel.cpp:2483-2511: value v is bit (v & 31) of present[v >> 5]
        uint32_t used16 = 0;
        for (int r = 0; r < 16; ++r) {
            uint32_t bits16 = (present[b * 8 + (r >> 1)] >> ((r & 1) * 16)) & 0xffffu;
            if (bits16) used16 |= 1u << (15 - r);
        }
        bw.put(16, used16);
        for (int r = 0; r < 16; ++r) {
            uint32_t bits16 = (present[b * 8 + (r >> 1)] >> ((r & 1) * 16)) & 0xffffu;
            if (bits16) bw.put(16, __brev(bits16) >> 16); // value k of the range goes out k-th: MSB first
        }
        bw.put(3, h.ntab);                                    // kernel.cpp:3001
        bw.put(15, h.nsel);                                   // :3002
        bw.finish();
    }
    const uint8_t *sp = selpos + (uint64_t)b * sel_stride;
    const uint32_t *so = selbitoff + (uint64_t)b * sel_stride;
    for (uint32_t g = threadIdx.x; g < h.nsel; g += EH_THREADS) { // writeUnary, :3010: p ones then a zero
        const uint32_t p = sp[g];
        or_bits(out, sel_start + so[g], p + 1, ((1u << p) - 1u) << 1);
    }
    uint64_t tstart = sel_start + sel_bits[b];
    for (uint32_t t = 0; t < h.ntab; ++t) {                   // :3015-3040
        if ((uint32_t)w == t) {
            const uint8_t *L = lens + ((uint64_t)b * HUFF_MAXTABLES + t) * HUFF_LEN_STRIDE;
            if (lane == 0) or_bits(out, tstart, 5, (uint32_t)L[0]);
            // lane l owns symbols 9 l .. 9 l + 8 (258 <= 32 * 9)
            const uint32_t j0 = 9u * (uint32_t)lane;
            int prev = j0 == 0 ? (int)L[0] : (j0 - 1 < h.alpha ? (int)L[j0 - 1] : 0);
            int len9[9];
            uint32_t mybits = 0;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const uint32_t j = j0 + (uint32_t)k;
                len9[k] = j < h.alpha ? (int)L[j] : -1;
                if (len9[k] >= 0) { const int d = len9[k] - prev; mybits += 2u * (uint32_t)(d < 0 ? -d : d) + 1u; prev = len9[k]; }
            }
            uint64_t pos = tstart + 5 + warp_incl_sum(mybits) - mybits;
            prev = j0 == 0 ? (int)L[0] : (j0 - 1 < h.alpha ? (int)L[j0 - 1] : 0);
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                if (len9[k] < 0) continue;
                const int d = len9[k] - prev;
                int steps = d < 0 ? -d : d;              // `steps` two-bit codes (10 = up, 11 = down), then a zero
                const uint32_t two = d < 0 ? 3u : 2u;
                while (steps > 0) {
                    const int take = steps < 15 ? steps : 15;
                    uint32_t v = 0;
                    for (int q = 0; q < take; ++q) v = (v << 2) | two;
                    or_bits(out, pos, 2u * (uint32_t)take, v);
                    pos += 2u * (uint32_t)take;
                    steps -= take;
                }
                pos += 1; // the terminating zero bit (the buffer is zero-initialised)
                prev = len9[k];
            }
        }
        tstart += per_table_bits[b * HUFF_MAXTABLES + t];
    }
    if (threadIdx.x == 0) data_start[b] = tstart;
}

// bits of the coded symbols per tile of HP_SYMS symbols
__global__ void __launch_bounds__(HP_THREADS) k_data_tilebits(const uint16_t *__restrict__ mtf, uint64_t mtf_stride, const HuffBlock *__restrict__ hb,
                                                               const uint8_t *__restrict__ lens, const uint8_t *__restrict__ sel, uint64_t sel_stride,
                                                               uint32_t *__restrict__ tile_bits, int tiles_per_block)
{
    __shared__ uint8_t s_len[HUFF_MAXTABLES][HUFF_LEN_STRIDE];
    __shared__ uint16_t s_sym[HP_SYMS];
    __shared__ uint32_t s_red[HP_THREADS / 32];
    const uint32_t b = blockIdx.y;
    const HuffBlock h = hb[b];
    const uint32_t sym0 = blockIdx.x * HP_SYMS;
    if (sym0 >= h.m) return;
    const uint32_t nsym = min((uint32_t)HP_SYMS, h.m - sym0);
    const uint8_t *L = lens + (uint64_t)b * HUFF_MAXTABLES * HUFF_LEN_STRIDE;
    for (uint32_t i = threadIdx.x; i < h.ntab * HUFF_LEN_STRIDE; i += HP_THREADS) (&s_len[0][0])[i] = L[i];
    const uint16_t *src = mtf + (uint64_t)b * mtf_stride + sym0;
    for (uint32_t i = threadIdx.x; i < nsym; i += HP_THREADS) s_sym[i] = src[i];
    __syncthreads();
    uint32_t bits = 0;
    const uint32_t g0 = threadIdx.x * HUFF_GROUP;
    if (g0 < nsym) {
        const uint32_t gc = min((uint32_t)HUFF_GROUP, nsym - g0);
        const uint32_t t = sel[(uint64_t)b * sel_stride + blockIdx.x * HP_THREADS + threadIdx.x];
        for (uint32_t k = 0; k < gc; ++k) bits += s_len[t][s_sym[g0 + k]];
    }
    bits = __reduce_add_sync(FULL, bits);
    if (lane_id() == 0) s_red[warp_id()] = bits;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
        for (int w = 0; w < HP_THREADS / 32; ++w) s += s_red[w];
        tile_bits[(uint64_t)b * tiles_per_block + blockIdx.x] = s;
    }
}

// one thread per block: exclusive prefix of tile bits
__global__ void __launch_bounds__(128) k_data_scan(const HuffBlock *__restrict__ hb, uint32_t *__restrict__ tile_bits, int tiles_per_block, uint32_t n_blocks)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const uint32_t nt = (hb[b].m + HP_SYMS - 1) / HP_SYMS;
    uint32_t run = 0;
    for (uint32_t t = 0; t < nt; ++t) {
        uint32_t v = tile_bits[(uint64_t)b * tiles_per_block + t];
        tile_bits[(uint64_t)b * tiles_per_block + t] = run;
        run += v;
    }
}

constexpr int ED_STAGE_WORDS = HP_SYMS * HUFF_MAXLEN / 32 + 4; // 4004 words

// one thread per 50-symbol group assembles its codes; the CTA stages the tile's bits in shared memory and stores words
__global__ void __launch_bounds__(HP_THREADS) k_emit_data(const uint16_t *__restrict__ mtf, uint64_t mtf_stride, const HuffBlock *__restrict__ hb,
                                                           const uint32_t *__restrict__ codes, const uint8_t *__restrict__ sel, uint64_t sel_stride,
                                                           const uint32_t *__restrict__ tile_bits, int tiles_per_block, const uint64_t *__restrict__ data_start,
                                                           uint32_t *__restrict__ out)
{
    __shared__ uint32_t s_code[HUFF_MAXTABLES][HUFF_MAXALPHA];
    __shared__ uint16_t s_sym[HP_SYMS];
    __shared__ uint32_t s_stage[ED_STAGE_WORDS];
    __shared__ uint32_t s_u[34];
    const uint32_t b = blockIdx.y;
    const HuffBlock h = hb[b];
    const uint32_t sym0 = blockIdx.x * HP_SYMS;
    if (sym0 >= h.m || data_start[b] == BIT_NOWHERE) return;
    const uint32_t nsym = min((uint32_t)HP_SYMS, h.m - sym0);
    const uint32_t *C = codes + (uint64_t)b * HUFF_MAXTABLES * HUFF_MAXALPHA;
    for (uint32_t i = threadIdx.x; i < h.ntab * HUFF_MAXALPHA; i += HP_THREADS) (&s_code[0][0])[i] = C[i];
    const uint16_t *src = mtf + (uint64_t)b * mtf_stride + sym0;
    for (uint32_t i = threadIdx.x; i < nsym; i += HP_THREADS) s_sym[i] = src[i];
    for (uint32_t i = threadIdx.x; i < ED_STAGE_WORDS; i += HP_THREADS) s_stage[i] = 0;
    __syncthreads();
    const uint32_t g0 = threadIdx.x * HUFF_GROUP;
    uint32_t gc = 0, t = 0, bits = 0;
    if (g0 < nsym) {
        gc = min((uint32_t)HUFF_GROUP, nsym - g0);
        t = sel[(uint64_t)b * sel_stride + blockIdx.x * HP_THREADS + threadIdx.x];
        for (uint32_t k = 0; k < gc; ++k) bits += s_code[t][s_sym[g0 + k]] >> 24;
    }
    uint32_t total;
    uint32_t off = block_excl_sum(bits, s_u, &total);
    const uint64_t tile_bit0 = data_start[b] + tile_bits[(uint64_t)b * tiles_per_block + blockIdx.x];
    const uint32_t lead = (uint32_t)(tile_bit0 & 31); // bit offset of the tile inside its first word
    if (gc) {
        uint32_t pos = lead + off; // bit position relative to the first staged word
        uint32_t wi = pos >> 5, used = pos & 31, acc = 0;
        for (uint32_t k = 0; k < gc; ++k) {
            uint32_t cs = s_code[t][s_sym[g0 + k]];
            uint32_t len = cs >> 24, code = cs & 0xffffffu;
            uint32_t room = 32 - used;
            if (len <= room) {
                acc |= code << (room - len);
                used += len;
                if (used == 32) { atomicOr(&s_stage[wi], acc); ++wi; acc = 0; used = 0; }
            } else {
                acc |= code >> (len - room);
                atomicOr(&s_stage[wi], acc);
                ++wi;
                uint32_t rest = len - room;
                acc = (code & ((1u << rest) - 1u)) << (32 - rest);
                used = rest;
            }
        }
        if (used) atomicOr(&s_stage[wi], acc);
    }
    __syncthreads();
    const uint32_t nwords = (lead + total + 31) >> 5;
    uint32_t *dst = out + (tile_bit0 >> 5);
    for (uint32_t i = threadIdx.x; i < nwords; i += HP_THREADS) {
        uint32_t v = s_stage[i];
        if (i == 0 || i == nwords - 1) { if (v) atomicOr(dst + i, bswap32(v)); }
        else dst[i] = bswap32(v);
    }
}

// ---------------------------------------------------------------------------------------------------------------- bit shift
// Multi-GPU streams: every device packs the blocks of its slice from bit 0 of its own buffer; once the bit lengths of the
// slices in front are known, the part moves to its place in the stream -- `shift` (0..7) bits to the right, the top `shift` bits of
// its first byte taken from `carry` (the trailing bits of the part in front).  This is the device-side form of the reference's
// leftBuffer hand-over between blocks (BitOutputStream.hpp:30-45, OutputStream.hpp:225-239).  src has `n_src` bytes (zero bits behind
// the part's last bit); dst gets n_dst = ceil((shift + nbits) / 8) bytes.  16 output bytes per thread.
__global__ void __launch_bounds__(256) k_shift_bits(const uint8_t *__restrict__ src, uint64_t n_src, uint8_t *__restrict__ dst, uint64_t n_dst, uint32_t shift,
                                                     uint32_t carry)
{
    const uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (i0 >= n_dst) return;
    // dst byte i = (src[i-1] << (8 - shift) | src[i] >> shift) & 0xff, src[-1] = carry bits
    uint8_t in[17];
    in[0] = i0 == 0 ? (uint8_t)0 : src[i0 - 1];
    if (i0 + 16 <= n_src && ((reinterpret_cast<uintptr_t>(src) + i0) & 15u) == 0) {
        const uint4 q = *reinterpret_cast<const uint4 *>(src + i0);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 16; ++k) in[1 + k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
    } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) in[1 + k] = i0 + k < n_src ? src[i0 + k] : (uint8_t)0;
    }
    uint8_t o[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) o[k] = (uint8_t)((((uint32_t)in[k] << 8) | in[k + 1]) >> shift);
    if (i0 == 0) o[0] = (uint8_t)((o[0] & (0xffu >> shift)) | carry);
    if (i0 + 16 <= n_dst && ((reinterpret_cast<uintptr_t>(dst) + i0) & 15u) == 0) {
        uint4 q;
        q.x = o[0] | (o[1] << 8) | (o[2] << 16) | ((uint32_t)o[3] << 24);
        q.y = o[4] | (o[5] << 8) | (o[6] << 16) | ((uint32_t)o[7] << 24);
        q.z = o[8] | (o[9] << 8) | (o[10] << 16) | ((uint32_t)o[11] << 24);
        q.w = o[12] | (o[13] << 8) | (o[14] << 16) | ((uint32_t)o[15] << 24);
        *reinterpret_cast<uint4 *>(dst + i0) = q;
    } else {
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if (i0 + k < n_dst) dst[i0 + k] = o[k];
    }
}

} // namespace bz
