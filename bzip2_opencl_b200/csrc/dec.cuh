// dec.cuh -- the decompress side: InputStream / BlockDecompressor / HuffmanStageDecoder / MoveToFront /
// BitInputStream of the reference (include/InputStream.hpp:36-159, BlockDecompressor.hpp:37-283, HuffmanStageDecoder.hpp:33-137,
// MoveToFront.hpp:31-75, BitInputStream.hpp:30-85 -- one sequential CPU loop there) as device kernels:
//   k_dec_scan        every bit position of the stream is tested for the 48-bit block / end-of-stream magics (blocks are
//                     independent once their start is known; false positives are dropped by the host's chain check)
//   k_dec_block       one warp per block: header, symbol map, selectors, delta-coded tables -> canonical decode tables + a 10-bit
//                     lookup window in shared memory; Huffman symbols -> move-to-front ranks (RUNA/RUNB runs = rank 0 stretches)
//   k_imtf_tiles<0> / k_imtf_compose / k_imtf_tiles<1>   inverse move-to-front, tile-parallel -> BWT column L
//   (k_bwt_bytehist + k_radix_bases + k_radix_onesweep with 1-byte keys from bwt.cuh: stable counting sort of L = the
//                     reference's bwtMergedPointers, BlockDecompressor.hpp:244-263)
//   k_dec_double / k_dec_double8 / k_dec_double16   four rounds of pointer doubling: every merged pointer carries sixteen bytes
//                     and the pointer sixteen steps on
//   k_dec_chase       one lane per block (a warp of its own) follows them from origPtr (:265-282): the RLE1-coded block, 16 bytes per dependent load,
//                     and the RLE1 decoded length on the way
//   k_dec_expand      RLE1 decode (:56-84) at the final offset: literal stretches copied by whole warps, only the four-equal
//                     triggers parsed in order
//   (k_crc_chunks / k_crc_finish from rle.cuh verify the block CRCs over the decoded output.)
#pragma once
#include "util.cuh"
#include "bwt.cuh"

namespace bz {

constexpr uint64_t DEC_BLOCK_MAGIC = 0x314159265359ull; // Config.hpp: BLOCK_HEADER_MARKER_1/2
constexpr uint64_t DEC_END_MAGIC = 0x177245385090ull;   // Config.hpp: STREAM_END_MARKER_1/2
constexpr int DEC_MAXLEN = 23;                          // HUFFMAN_DECODE_MAXIMUM_CODE_LENGTH (Config.hpp)

enum DecErr { // the reference's exception texts, by number (bz2b200.cu maps them back)
    DEC_OK = 0, DEC_E_TABLES = 1, DEC_E_BLOCKSIZE = 2, DEC_E_DECODE = 3, DEC_E_STARTPTR = 4, DEC_E_DATA = 5, DEC_E_RANDOMISED = 6
};

struct DecInfo {
    uint64_t end_bit;   // first bit after the block's end-of-block symbol
    uint32_t crc;       // stored block CRC
    uint32_t orig;      // origPtr
    uint32_t n;         // BWT block length
    uint32_t err;       // DecErr
};

// ---- candidates: every bit position holding one of the two magics
__global__ void __launch_bounds__(256) k_dec_scan(const uint8_t *__restrict__ in, uint64_t n_bytes, uint64_t *__restrict__ cand, uint32_t cap,
                                                   uint32_t *__restrict__ count)
{
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_bytes) return;
    uint64_t w = 0; // bytes p .. p+6, big-endian, in bits 55..0
#pragma unroll
    for (int k = 0; k < 7; ++k) w = (w << 8) | (p + k < n_bytes ? (uint64_t)__ldg(in + p + k) : 0ull);
#pragma unroll
    for (int o = 0; o < 8; ++o) {
        const uint64_t v = (w >> (8 - o)) & 0xffffffffffffull;
        if (v == DEC_BLOCK_MAGIC || v == DEC_END_MAGIC) {
            const uint32_t k = atomicAdd(count, 1u);
            if (k < cap) cand[k] = ((p * 8 + o) << 1) | (v == DEC_END_MAGIC ? 1ull : 0ull);
        }
    }
}

// ---- sequential MSB-first bit reader (BitInputStream.hpp:30-85).  Every lane of the decoding warp keeps the same copy, so
// the loads are broadcasts and the control flow stays uniform.  A 64-bit window is refilled 32 bits at a time.
struct BitReader {
    const uint8_t *in;
    uint64_t win;      // next bits, MSB first, left aligned
    uint32_t have;     // valid bits in win
    uint32_t ahead;    // the word after the window, loaded one refill early so its latency is off the decode chain
    uint64_t next;     // byte offset of the word after `ahead` (multiple of 4)
    uint64_t pos, end; // bit position, total bits
    bool overrun;
    __device__ __forceinline__ uint32_t word(uint64_t byte) const
    {
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(in + byte)); // the buffer is 16-byte aligned and zero padded
        return __byte_perm(w, 0, 0x0123);
    }
    __device__ void init(const uint8_t *p, uint64_t bitpos, uint64_t nbits, uint64_t nbytes_padded)
    {
        in = p; pos = bitpos; end = nbits; overrun = false;
        limit_byte = nbytes_padded - 4;
        const uint64_t b0 = (bitpos >> 3) & ~3ull;
        win = ((uint64_t)word(b0) << 32) | word(min(b0 + 4, limit_byte));
        const uint32_t skip = (uint32_t)(bitpos - b0 * 8);
        win <<= skip;
        have = 64 - skip;
        ahead = word(min(b0 + 8, limit_byte));
        next = b0 + 12;
    }
    uint64_t limit_byte; // last word offset that may be loaded
    __device__ __forceinline__ uint32_t peek(int count) // 1..24 bits without consuming
    {
        if (have < 32) {
            win |= (uint64_t)ahead << (32 - have);
            have += 32;
            ahead = word(min(next, limit_byte));
            next += 4;
        }
        return (uint32_t)(win >> (64 - count));
    }
    __device__ __forceinline__ void skip(int count)
    {
        win <<= count;
        have -= count;
        pos += count;
    }
    // reading past the end only yields the zero padding; callers test this at their own pace (per field / per 50-symbol group)
    __device__ __forceinline__ bool past_end() { if (pos > end) overrun = true; return overrun; }
    __device__ __forceinline__ uint32_t bits(int count) // 0..24
    {
        if (count == 0) return 0;
        const uint32_t v = peek(count);
        skip(count);
        if (pos > end) overrun = true;
        return v;
    }
    __device__ __forceinline__ uint32_t bit() { return bits(1); }
};

// canonical decode tables of one block (HuffmanStageDecoder.hpp:77-134), in the decoding warp's shared memory
constexpr int DEC_LUT_BITS = 10;
struct DecTables {
    int32_t limit[6][DEC_MAXLEN + 2];
    int32_t base[6][DEC_MAXLEN + 2];
    uint16_t perm[6][258];
    uint16_t lut[6][1 << DEC_LUT_BITS]; // next DEC_LUT_BITS bits -> (symbol << 5) | code length, 0 = code longer than the window
    uint8_t len[258];
    uint8_t minlen[6];
    uint8_t lutmin[6]; // first code length the limit search tries when the lookup window has no entry
};

constexpr int DB_WARPS = 4; // blocks per CTA of k_dec_block

// One WARP per candidate block.  The bit stream is sequential, so all lanes decode the same symbols redundantly (uniform control
// flow, broadcast loads).  Output: the move-to-front RANKS of the block (RUNA/RUNB runs = stretches of rank 0, which the
// pre-zeroed buffer already holds) and the initial list (the used byte values in order); the inverse MTF itself runs tile-parallel
// afterwards (k_imtf_*).
__global__ void __launch_bounds__(DB_WARPS * 32) k_dec_block(const uint8_t *__restrict__ in, uint64_t n_bits, const uint64_t *__restrict__ cand, uint32_t c0,
                                                              uint32_t nb, uint32_t block_size, uint8_t *__restrict__ ranks, uint64_t stride,
                                                              uint8_t *__restrict__ sel_scratch, uint64_t sel_stride, uint32_t max_selectors, uint32_t sel_cap,
                                                              uint64_t *__restrict__ list0, DecInfo *__restrict__ info)
{
    extern __shared__ __align__(16) unsigned char dec_smem[];
    DecTables *s_tab = reinterpret_cast<DecTables *>(dec_smem);
    const uint32_t bl = blockIdx.x * DB_WARPS + warp_id();
    if (bl >= nb) return;
    const int lane = lane_id();
    DecTables &T = s_tab[warp_id()];
    DecInfo I;
    I.end_bit = 0; I.crc = 0; I.orig = 0; I.n = 0; I.err = DEC_OK;
    BitReader br;
    br.init(in, (cand[c0 + bl] >> 1) + 48, n_bits, (n_bits >> 3) + 64);
    // BlockDecompressor ctor (BlockDecompressor.hpp:40-52)
    I.crc = (br.bits(16) << 16) | br.bits(16);
    const uint32_t randomised = br.bit();
    I.orig = br.bits(24);
    uint8_t *sel = sel_scratch + (uint64_t)bl * sel_stride;
    // readHuffmanTables (:137-187).  The move-to-front list starts as the used byte values in order (huffmanSymbolMap composed
    // with the identity list of MoveToFront), so its entries are output bytes directly.
    uint64_t list = 0; // entries 8*lane .. 8*lane+7, entry k in byte k
    int nsym = 0;
    const uint32_t ranges = br.bits(16);
    for (int i = 0; i < 16; ++i)
        if (ranges & (0x8000u >> i)) {
            const uint32_t m = br.bits(16);
            for (int j = 0; j < 16; ++j)
                if (m & (0x8000u >> j)) {
                    if ((nsym >> 3) == lane) list |= (uint64_t)(i * 16 + j) << (8 * (nsym & 7));
                    ++nsym;
                }
        }
    list0[(uint64_t)bl * 32 + lane] = list;
    const int eob = nsym + 1, alpha = nsym + 2;
    const uint32_t ntab = br.bits(3);
    uint32_t nsel = br.bits(15);
    if (br.overrun) I.err = DEC_E_DATA;
    else if (ntab < 2 || ntab > 6 || nsel < 1 || nsel > max_selectors) I.err = DEC_E_TABLES;
    else if (randomised) I.err = DEC_E_RANDOMISED;
    if (I.err == DEC_OK) {
        uint32_t tmtf = 0x543210u; // six 4-bit entries
        for (uint32_t i = 0; i < nsel && !br.overrun; ++i) { // selectors: unary MTF index (:164-167)
            uint32_t u = 0;
            while (br.bit()) { if (++u > 5 || br.overrun) break; }
            if (u > 5) { I.err = DEC_E_TABLES; break; } // (the reference indexes past its table list here; a 2..6-table stream never does)
            const uint32_t v = (tmtf >> (4 * u)) & 15u;
            if (v >= ntab) { I.err = DEC_E_TABLES; break; } // a table that the block does not define (the reference would index past its list)
            const uint32_t lowmask = (1u << (4 * u)) - 1u;
            tmtf = (tmtf & ~((lowmask << 4) | 15u)) | ((tmtf & lowmask) << 4) | v;
            // A block of the handle's size never needs more than sel_cap groups; selectors beyond that are read and dropped (the slot
            // holds sel_cap bytes -- a declared count above it must not reach the neighbouring blocks' slots).
            if (lane == 0 && i < sel_cap) sel[i] = (uint8_t)v;
        }
        nsel = min(nsel, sel_cap);
    }
    if (I.err == DEC_OK) {
        for (uint32_t t = 0; t < ntab && I.err == DEC_OK; ++t) { // delta-coded lengths (:169-183), then createHuffmanDecodingTables
            int cur = (int)br.bits(5);
            int mn = DEC_MAXLEN, mx = 0;
            for (int j = 0; j <= eob; ++j) {
                while (br.bit()) { cur += br.bit() ? -1 : 1; if (br.overrun) break; }
                if (cur < 1 || cur > DEC_MAXLEN) { I.err = DEC_E_TABLES; break; } // outside what the decode tables can hold
                if (lane == 0) T.len[j] = (uint8_t)cur;
                mn = min(mn, cur); mx = max(mx, cur);
            }
            if (I.err != DEC_OK) break;
            __syncwarp();
            int oversub = 0; // Kraft sum above 1: canonical codes of some length do not fit that length
            if (lane == 0) {
                int32_t *bases = T.base[t], *limits = T.limit[t];
                for (int i = 0; i < DEC_MAXLEN + 2; ++i) { bases[i] = 0; limits[i] = -1; }
                for (int i = 0; i < alpha; ++i) bases[T.len[i] + 1]++;
                for (int i = 1; i < DEC_MAXLEN + 2; ++i) bases[i] += bases[i - 1];
                int code = 0;
                for (int i = mn; i <= mx; ++i) {
                    const int b0 = code;
                    code += bases[i + 1] - bases[i];
                    if (code > (1 << i)) oversub = 1;
                    bases[i] = b0 - bases[i];
                    limits[i] = code - 1;
                    code <<= 1;
                }
                int ci = 0;
                for (int bits = mn; bits <= mx; ++bits)
                    for (int s = 0; s < alpha; ++s)
                        if (T.len[s] == bits) T.perm[t][ci++] = (uint16_t)s;
                T.minlen[t] = (uint8_t)mn;
                T.lutmin[t] = (uint8_t)(oversub ? mn : max(mn, DEC_LUT_BITS + 1));
            }
            oversub = __shfl_sync(FULL, oversub, 0);
            __syncwarp();
            // lookup window: every symbol with a code of at most DEC_LUT_BITS bits fills its 2^(window - length) slots.  An
            // over-subscribed length set (which the reference's limit search still decodes, HuffmanStageDecoder.hpp:62-74) gets no
            // window at all: its codes would run past the table, and every symbol then goes through the same limit search.
            for (int k = lane; k < (1 << DEC_LUT_BITS); k += 32) T.lut[t][k] = 0;
            __syncwarp();
            for (int ci = lane; ci < alpha && !oversub; ci += 32) { // canonical order: ci-th code
                const int sym = T.perm[t][ci];
                const int l = T.len[sym];
                if (l <= DEC_LUT_BITS) {
                    const int code = ci + T.base[t][l]; // inverse of  symbol = perm[code - base[len]]
                    const int lo = code << (DEC_LUT_BITS - l), cnt = 1 << (DEC_LUT_BITS - l);
                    for (int k = 0; k < cnt; ++k) T.lut[t][lo + k] = (uint16_t)((sym << 5) | l);
                }
            }
            __syncwarp();
        }
        if (br.overrun && I.err == DEC_OK) I.err = DEC_E_DATA;
    }
    // decodeHuffmanData (:189-241): Huffman symbol -> RUNA/RUNB run (rank 0 repeated) or rank = symbol - 1.
    // The reference reads one symbol after the other; here the warp decodes a WINDOW per step.  Lane k decodes the code that would
    // start at bit pos + k (lookup window, canonical search for the rare long codes); which of the 32 candidates are real symbol
    // starts follows from pointer doubling over "start + length" from bit 0 (log2(#symbols) rounds of one shuffle and one warp
    // OR-reduction); the real ones, in lane order, are the next symbols.  A step ends at the group's 50th symbol (the table
    // changes there, HuffmanStageDecoder.hpp:51-60), at the end-of-block symbol or at the first undecodable position.  Run
    // digits are weighted by their place inside their stretch of RUNA/RUNB symbols (carried across steps), so a prefix sum over
    // the window gives every literal its output index; zero runs are never written (the buffer is pre-zeroed).
    uint32_t n = 0;
    if (I.err == DEC_OK) {
        __syncwarp();
        uint8_t *out = ranks + (uint64_t)bl * stride;
        const uint32_t *in32 = reinterpret_cast<const uint32_t *>(in);
        const uint64_t limit_word = ((n_bits >> 3) + 64) / 4 - 1; // last word of the zero-padded buffer
        const uint64_t end = n_bits;
        uint64_t pos = br.pos;
        // the stream around pos: 3 chunks of 32 big-endian words, one word per lane; the third is loaded a chunk ahead of its use
        uint64_t chunk = pos >> 10;
        uint32_t cur, nxt, nn;
        {
            const uint64_t w0 = chunk * 32 + lane;
            cur = __byte_perm(__ldg(in32 + min(w0, limit_word)), 0, 0x0123);
            nxt = __byte_perm(__ldg(in32 + min(w0 + 32, limit_word)), 0, 0x0123);
            nn = __byte_perm(__ldg(in32 + min(w0 + 64, limit_word)), 0, 0x0123);
        }
        const unsigned lt = (1u << lane) - 1u;
        uint32_t group = 0, in_group = 0, runk = 0; // runk: RUNA/RUNB symbols at the tail of what has been decoded (open stretch)
        uint32_t tbl = sel[0], tbl_next = sel[nsel > 1 ? 1 : 0];
        const uint16_t *lutp = T.lut[tbl];
        for (;;) {
            if (in_group == 50) { // HuffmanStageDecoder::nextSymbol group switch (:51-60)
                in_group = 0;
                if (++group == nsel) { I.err = DEC_E_DECODE; break; }
                tbl = tbl_next;
                tbl_next = sel[min(group + 1, nsel - 1)];
                lutp = T.lut[tbl];
                if (pos > end) { I.err = DEC_E_DATA; break; }
            }
            const uint64_t c2 = pos >> 10;
            if (c2 != chunk) { // a step moves pos by less than two words: at most one chunk further
                chunk = c2;
                cur = nxt; nxt = nn;
                nn = __byte_perm(__ldg(in32 + min(chunk * 32 + 64 + lane, limit_word)), 0, 0x0123);
            }
            const uint32_t wi = (uint32_t)(pos >> 5) & 31u, off = (uint32_t)pos & 31u;
            const uint32_t W0 = __shfl_sync(FULL, cur, wi);
            const uint32_t W1 = wi + 1 < 32 ? __shfl_sync(FULL, cur, wi + 1) : __shfl_sync(FULL, nxt, wi + 1 - 32);
            const uint32_t W2 = wi + 2 < 32 ? __shfl_sync(FULL, cur, wi + 2) : __shfl_sync(FULL, nxt, wi + 2 - 32);
            const uint32_t hi = __funnelshift_l(W1, W0, off), lo = __funnelshift_l(W2, W1, off);
            const uint32_t v = __funnelshift_l(lo, hi, lane); // bits [pos + lane, pos + lane + 32)
            // ---- the code starting at this lane's bit
            int sym = -1;
            uint32_t len = 32;
            const uint32_t e = lutp[v >> (32 - DEC_LUT_BITS)];
            if (e) { sym = (int)(e >> 5); len = e & 31u; }
            else { // longer than the lookup window: the reference's limit search (HuffmanStageDecoder.hpp:62-74)
                for (int clen = (int)T.lutmin[tbl]; clen <= DEC_MAXLEN; ++clen) {
                    const int32_t code = (int32_t)(v >> (32 - clen));
                    if (code <= T.limit[tbl][clen]) {
                        const int32_t idx = code - T.base[tbl][clen];
                        if ((uint32_t)idx < (uint32_t)alpha) { sym = T.perm[tbl][idx]; len = (uint32_t)clen; }
                        break;
                    }
                }
            }
            const uint32_t nx1 = (uint32_t)lane + len; // where the next symbol starts (relative to pos)
            // ---- which lanes are symbol starts: reachable from bit 0
            uint32_t jump = min(nx1, 32u), mask = 1u;
            for (;;) {
                const bool mem = (mask >> lane) & 1u;
                const uint32_t add = __reduce_or_sync(FULL, (mem && jump < 32u) ? (1u << jump) : 0u);
                if ((add & ~mask) == 0u) break;
                mask |= add;
                const uint32_t j2 = __shfl_sync(FULL, jump, jump & 31u);
                jump = jump < 32u ? j2 : 32u;
            }
            bool acc = ((mask >> lane) & 1u) && (uint32_t)__popc(mask & lt) < 50u - in_group;
            const unsigned eobm = __ballot_sync(FULL, acc && sym == eob);
            if (eobm) acc = acc && lane < __ffs(eobm);
            const unsigned A = __ballot_sync(FULL, acc);
            const bool isrun = acc && (uint32_t)sym <= 1u;
            const unsigned RUN = __ballot_sync(FULL, isrun);
            uint32_t c = 0, lerr = DEC_OK, place = 0;
            if (isrun) { // digit number `place` of its stretch: weight (sym + 1) << place  (:203-214)
                const unsigned nr = A & ~RUN & lt, below = RUN & lt;
                place = nr == 0u ? (uint32_t)__popc(below) + runk : (uint32_t)__popc(below >> (32 - __clz(nr)));
                if (place >= 24u) lerr = DEC_E_BLOCKSIZE;
                else c = (uint32_t)(sym + 1) << place;
            } else if (acc && sym >= 0 && sym != eob) c = 1;
            uint32_t incl;
            if (RUN == 0u) incl = (uint32_t)__popc(__ballot_sync(FULL, c != 0u) & (lt | (1u << lane))); // literals only: count them
            else {
                incl = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(FULL, incl, d);
                    if (lane >= d) incl += t;
                }
            }
            const uint32_t at = n + incl - c; // ranks in front of this symbol
            if (acc && !isrun) {
                if (sym < 0) lerr = (pos + (uint32_t)lane + DEC_MAXLEN > end) ? DEC_E_DATA : DEC_E_DECODE;
                else if (sym == eob ? at > block_size : at >= block_size) lerr = DEC_E_BLOCKSIZE;
            }
            const unsigned errm = __ballot_sync(FULL, acc && lerr != DEC_OK);
            if (errm) { I.err = __shfl_sync(FULL, lerr, __ffs(errm) - 1); break; }
            if (acc && !isrun && sym != eob) out[at] = (uint8_t)(sym - 1);
            const int klast = 31 - __clz(A); // lane 0 always is a symbol start
            n += __shfl_sync(FULL, incl, 31);
            pos += __shfl_sync(FULL, nx1, klast);
            const uint32_t lplace = __shfl_sync(FULL, place, klast);
            runk = ((RUN >> klast) & 1u) ? lplace + 1u : 0u;
            in_group += (uint32_t)__popc(A);
            if (eobm) { if (pos > end) I.err = DEC_E_DATA; break; }
        }
        br.pos = pos;
    }
    if (I.err == DEC_OK && I.orig >= n) I.err = DEC_E_STARTPTR; // initialiseInverseBWT (:244-248)
    I.n = I.err == DEC_OK ? n : 0u;
    I.end_bit = br.pos;
    if (lane == 0) info[bl] = I;
}

// ---- inverse move-to-front, tile-parallel (the mirror of mtf.cuh): applying a tile's ranks to ANY list permutes its POSITIONS
// the same way, so (1) every tile's position permutation comes from running its ranks on the identity, (2) one warp per block
// composes them into the list at every tile start, (3) every tile re-runs from its true start list and writes the bytes.
// The 256-entry list lives in one 64-bit register per lane (entries 8*lane .. 8*lane+7), moved with shuffles
// (MoveToFront::indexToFront, MoveToFront.hpp:62-72).
constexpr int IM_TILE = 2048;
constexpr int IM_WARPS = 8;

__device__ __forceinline__ uint32_t list_index_to_front(uint64_t &list, uint32_t idx, int lane)
{
    const int src = (int)(idx >> 3), bi = (int)(idx & 7);
    const uint64_t srcw = shfl_u64(list, src);
    const uint32_t v = (uint32_t)(srcw >> (8 * bi)) & 0xffu;
    const uint64_t up = shfl_up_u64(list, 1); // previous lane's entries
    const uint64_t shifted = (list << 8) | (lane == 0 ? (uint64_t)v : (up >> 56));
    if (lane < src) list = shifted;
    else if (lane == src) {
        const uint64_t lowmask = bi == 7 ? ~0ull : ((1ull << (8 * (bi + 1))) - 1ull); // entries 0..bi of this lane change
        list = (list & ~lowmask) | (shifted & lowmask);
    }
    return v;
}

// APPLY = false: perm[tile] = positions after the tile's ranks, from the identity.  APPLY = true: lists[tile] is the true start
// list; write the bytes.
template <bool APPLY>
__global__ void __launch_bounds__(IM_WARPS * 32) k_imtf_tiles(const uint8_t *__restrict__ ranks, const uint32_t *__restrict__ counts, uint64_t stride,
                                                               uint64_t *__restrict__ lists, int tiles_per_block, uint8_t *__restrict__ L)
{
    const uint32_t bl = blockIdx.y;
    const uint32_t n = counts[bl];
    const uint32_t tile = blockIdx.x * IM_WARPS + warp_id();
    const uint32_t base = tile * IM_TILE;
    if (base >= n) return;
    const int lane = lane_id();
    const uint32_t cnt = min((uint32_t)IM_TILE, n - base);
    uint64_t *lp = lists + ((uint64_t)bl * tiles_per_block + tile) * 32 + lane;
    uint64_t list;
    if (APPLY) list = *lp;
    else {
        list = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) list |= (uint64_t)(lane * 8 + k) << (8 * k);
    }
    const uint8_t *rp = ranks + (uint64_t)bl * stride + base;
    uint8_t *op = L + (uint64_t)bl * stride + base;
    uint32_t front = (uint32_t)__shfl_sync(FULL, (uint32_t)list, 0) & 0xffu;
    for (uint32_t c0 = 0; c0 < cnt; c0 += 32) {
        const bool valid = c0 + lane < cnt;
        const uint32_t r = valid ? rp[c0 + lane] : 0u;
        unsigned nz = __ballot_sync(FULL, r != 0);
        uint32_t mine = front; // rank 0: the value at the head when this position is reached
        while (nz) {
            const int q = __ffs((int)nz) - 1;
            nz &= nz - 1;
            const uint32_t v = list_index_to_front(list, __shfl_sync(FULL, r, q), lane);
            if (lane >= q) mine = v;
            front = v;
        }
        if (APPLY && valid) op[c0 + lane] = (uint8_t)mine;
    }
    if (!APPLY) *lp = list;
}

// one warp per block: lists[t] (position permutation of tile t) -> lists[t] = list at the start of tile t
__global__ void __launch_bounds__(IM_WARPS * 32) k_imtf_compose(const uint32_t *__restrict__ counts, uint64_t *__restrict__ lists, int tiles_per_block,
                                                                 const uint64_t *__restrict__ list0, uint32_t nb)
{
    __shared__ __align__(8) uint8_t s_cur[IM_WARPS][256];
    const uint32_t bl = blockIdx.x * IM_WARPS + warp_id();
    if (bl >= nb) return;
    const int lane = lane_id();
    const uint32_t n = counts[bl];
    const uint32_t nt = (n + IM_TILE - 1) / IM_TILE;
    uint8_t *cur = s_cur[warp_id()];
    uint64_t c8 = list0[(uint64_t)bl * 32 + lane];
    for (uint32_t t = 0; t < nt; ++t) {
        uint64_t *lp = lists + ((uint64_t)bl * tiles_per_block + t) * 32 + lane;
        const uint64_t p8 = *lp; // positions: new entry k = old entry p[k]
        *lp = c8;
        *reinterpret_cast<uint64_t *>(cur + lane * 8) = c8;
        __syncwarp();
        uint64_t nx = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) nx |= (uint64_t)cur[(uint32_t)(p8 >> (8 * k)) & 0xffu] << (8 * k);
        __syncwarp();
        c8 = nx;
    }
}

// counts[bl] = block length (0 for a failed block)
__global__ void __launch_bounds__(256) k_dec_counts(const DecInfo *__restrict__ info, uint32_t nb, uint32_t *__restrict__ counts)
{
    const uint32_t bl = blockIdx.x * blockDim.x + threadIdx.x;
    if (bl < nb) counts[bl] = info[bl].n;
}

// Merged pointers (BlockDecompressor.hpp:256-262) = the sort elements of bwt.cuh after a stable sort of L by byte value:
// element j = (L[i] << 20) | i for the j-th smallest byte.  The chase (:265-282) is one dependent load per byte; two rounds of
// pointer doubling (fully parallel gathers) make every element carry FOUR bytes and the pointer four steps on:
//   bits 0..19 pointer, bits 20..27 / 28..35 / 36..43 / 44..51 the bytes at hops 0..3.
__global__ void __launch_bounds__(256) k_dec_double(const uint64_t *__restrict__ src, uint64_t *__restrict__ dst, const uint32_t *__restrict__ counts,
                                                     uint64_t stride, int have_bytes)
{
    const uint32_t bl = blockIdx.y;
    const uint32_t n = counts[bl];
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint64_t *m = src + (uint64_t)bl * stride;
    const uint64_t x = m[j];
    const uint64_t y = m[x & VAL_MASK];
    const uint64_t mask = (1ull << (8 * have_bytes)) - 1ull;
    dst[(uint64_t)bl * stride + j] = (y & VAL_MASK) | (((x >> VAL_BITS) & mask) << VAL_BITS) | (((y >> VAL_BITS) & mask) << (VAL_BITS + 8 * have_bytes));
}

// third doubling round: eight bytes per element, split over two arrays (bytes8[j], ptr8[j]) that the chase loads side by side
__global__ void __launch_bounds__(256) k_dec_double8(const uint64_t *__restrict__ src4, uint64_t *__restrict__ bytes8, uint32_t *__restrict__ ptr8,
                                                      const uint32_t *__restrict__ counts, uint64_t stride)
{
    const uint32_t bl = blockIdx.y;
    const uint32_t n = counts[bl];
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint64_t *m = src4 + (uint64_t)bl * stride;
    const uint64_t x = m[j];
    const uint64_t y = m[x & VAL_MASK];
    bytes8[(uint64_t)bl * stride + j] = ((x >> VAL_BITS) & 0xffffffffull) | ((y >> VAL_BITS) << 32);
    ptr8[(uint64_t)bl * stride + j] = (uint32_t)(y & VAL_MASK);
}

// fourth doubling round: sixteen bytes per element: lo8[j] stays, hi8[j] = bytes8[ptr8[j]], ptr16[j] = ptr8[ptr8[j]]
__global__ void __launch_bounds__(256) k_dec_double16(const uint64_t *__restrict__ bytes8, const uint32_t *__restrict__ ptr8, uint64_t *__restrict__ hi8,
                                                       uint32_t *__restrict__ ptr16, const uint32_t *__restrict__ counts, uint64_t stride)
{
    const uint32_t bl = blockIdx.y;
    const uint32_t n = counts[bl];
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint64_t o = (uint64_t)bl * stride;
    const uint32_t q = ptr8[o + j];
    hi8[o + j] = bytes8[o + q];
    ptr16[o + j] = ptr8[o + q];
}

// one lane per block follows the sixteen-byte elements from origPtr and writes the RLE1-coded block; the RLE1 state machine
// (BlockDecompressor::read, :56-84) runs along to give the decoded length (its cost hides behind the dependent loads)
__global__ void __launch_bounds__(128) k_dec_chase(const uint64_t *__restrict__ lo8, const uint64_t *__restrict__ hi8, const uint32_t *__restrict__ ptr16,
                                                   uint64_t stride, const DecInfo *__restrict__ info, uint32_t nb, uint8_t *__restrict__ rle,
                                                   uint64_t *__restrict__ out_len)
{
    // one block per WARP, lane 0 walks: threads of one warp move in lockstep, so 32 chains in a warp would each pay the slowest of
    // 32 DRAM round trips at every step
    const uint32_t bl = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (bl >= nb || (threadIdx.x & 31u) != 0u) return;
    const uint32_t n = info[bl].n;
    if (n == 0) { out_len[bl] = 0; return; }
    const uint64_t base = (uint64_t)bl * stride;
    const uint64_t *ml = lo8 + base, *mh = hi8 + base;
    const uint32_t *mp = ptr16 + base;
    uint8_t *out = rle + base;
    uint32_t p = info[bl].orig;
    uint64_t o = 0;
    int last = -1;
    uint32_t acc = 0;
    bool count_next = false;
    auto step = [&](uint32_t v) {
        if (count_next) { o += v; count_next = false; acc = 0; }
        else if ((int)v != last) { last = (int)v; acc = 1; ++o; }
        else { ++o; if (++acc == 4) count_next = true; }
    };
    uint32_t i = 0;
    for (; i + 16 <= n; i += 16) {
        const uint64_t w0 = ml[p], w1 = mh[p];
        p = mp[p];
        *reinterpret_cast<uint4 *>(out + i) = make_uint4((uint32_t)w0, (uint32_t)(w0 >> 32), (uint32_t)w1, (uint32_t)(w1 >> 32));
#pragma unroll
        for (int k = 0; k < 8; ++k) step((uint32_t)(w0 >> (8 * k)) & 0xffu);
#pragma unroll
        for (int k = 0; k < 8; ++k) step((uint32_t)(w1 >> (8 * k)) & 0xffu);
    }
    if (i < n) {
        const uint64_t w0 = ml[p], w1 = mh[p];
        for (uint32_t k = 0; i < n; ++i, ++k) {
            const uint32_t v = (uint32_t)((k < 8 ? w0 >> (8 * k) : w1 >> (8 * (k - 8))) & 0xffu);
            out[i] = (uint8_t)v;
            step(v);
        }
    }
    out_len[bl] = o;
}

// RLE1 decode (BlockDecompressor::read, :56-84) of one block at its final offset, one WARP per block.  A count byte exists only
// behind four equal literals, so the warp works on 1 KiB chunks: every lane marks, for its 32 bytes, where a byte completes three
// consecutive "equal to its predecessor" steps (a possible fourth literal); between such triggers everything is a literal and is
// copied by all lanes at once, and only the triggers themselves are parsed in order.
constexpr int DX_WARPS = 4;
constexpr int DX_CHUNK = 1024;

__global__ void __launch_bounds__(DX_WARPS * 32) k_dec_expand(const uint8_t *__restrict__ rle, uint64_t stride, const DecInfo *__restrict__ info, uint32_t nb,
                                                               const uint64_t *__restrict__ out_off, uint8_t *__restrict__ out)
{
    __shared__ __align__(16) uint8_t s_buf[DX_WARPS][DX_CHUNK + 16];
    const uint32_t bl = blockIdx.x * DX_WARPS + warp_id();
    if (bl >= nb || out_off[bl] == ~0ull) return;
    const int lane = lane_id();
    const uint32_t n = info[bl].n;
    const uint8_t *src = rle + (uint64_t)bl * stride;
    uint8_t *dst = out + out_off[bl];
    uint8_t *buf = s_buf[warp_id()] + 16; // buf[-3..-1] = last three bytes of the previous chunk
    uint64_t o = 0;
    bool count_next = false;   // the next byte is a repeat count for `runval`
    uint32_t runval = 0;
    uint32_t lit_start = 0;    // absolute index of the first byte that may belong to the current run of literals
    for (uint32_t c0 = 0; c0 < n; c0 += DX_CHUNK) {
        const uint32_t cn = min((uint32_t)DX_CHUNK, n - c0);
        __syncwarp();
        if (c0 > 0 && lane < 3) buf[-3 + lane] = buf[DX_CHUNK - 3 + lane];
        __syncwarp();
        for (uint32_t k = lane * 16; k < cn; k += 32 * 16) *reinterpret_cast<uint4 *>(buf + k) = __ldg(reinterpret_cast<const uint4 *>(src + c0 + k)); // slot padded
        __syncwarp();
        // trig bit j of lane l: bytes (32l+j-3 .. 32l+j) of the chunk are equal (indices relative to the chunk, -3.. allowed)
        uint32_t trig = 0;
        for (int j = 0; j < 32; ++j) {
            const int q = lane * 32 + j;
            if ((uint32_t)q < cn && (int64_t)c0 + q - 3 >= 0) {
                const uint8_t v = buf[q];
                if (buf[q - 1] == v && buf[q - 2] == v && buf[q - 3] == v) trig |= 1u << j;
            }
        }
        uint32_t i = 0; // next chunk byte to handle
        while (i < cn) {
            if (count_next) { // buf[i] more copies of the run value
                const uint32_t rep = buf[i];
                for (uint32_t k = lane; k < rep; k += 32) dst[o + k] = (uint8_t)runval;
                o += rep;
                count_next = false;
                ++i;
                lit_start = c0 + i;
                continue;
            }
            // first trigger at or after max(i, lit_start + 3): its four bytes are literals of the current parse
            const uint32_t from = max(c0 + i, lit_start + 3) - c0; // chunk-relative, may be >= cn
            uint32_t e = cn; // chunk-relative position of the trigger, cn = none in this chunk
            {
                const int fl = (int)(from >> 5);
                uint32_t m = trig;
                if (lane < fl) m = 0;
                else if (lane == fl) m &= ~0u << (from & 31);
                const unsigned any = __ballot_sync(FULL, m != 0 && from < cn);
                if (any) {
                    const int l = __ffs((int)any) - 1;
                    const uint32_t mm = __shfl_sync(FULL, m, l);
                    e = (uint32_t)l * 32 + (uint32_t)(__ffs((int)mm) - 1);
                }
            }
            const uint32_t upto = e < cn ? e + 1 : cn; // literals [i, upto)
            for (uint32_t k = i + lane; k < upto; k += 32) dst[o + (k - i)] = buf[k];
            o += upto - i;
            if (e < cn) { count_next = true; runval = buf[e]; }
            i = upto;
        }
    }
}

} // namespace bz
