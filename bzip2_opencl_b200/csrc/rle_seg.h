// rle_seg.h -- algebra of the reference's RLE1 / block-boundary state machine, usable on host and device.
//
// The reference decides block boundaries with a sequential per-byte state machine
// (BlockCompressor::write, /root/reference/include/BlockCompressor.hpp:69-96): a byte is refused when
// blockLength > blockSize-6 at entry (:71-74), runs are cut at 255 (:84-88), a run of L >= 4 is stored as
// 4 bytes + the count byte L-4 (writeRun, :134-154), and a new block restarts the chunking.
//
// To run that machine in parallel we summarise any byte range by a `Seg` whose composition is associative,
// and we can push the machine's state (blockLength, current value, pending run length) through a Seg in
// O(1).  Tiles of the input are summarised independently, combined in a 64-ary tree, and the (short)
// sequential walk over blocks steps through that tree instead of through bytes.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BZ_HD __host__ __device__ __forceinline__
#else
#define BZ_HD inline
#endif

namespace bz {

// encoded size of a pending run of r (0..254) bytes when it is flushed: writeRun(value, r)
BZ_HD uint32_t rle_enc_small(uint32_t r) { return r < 4u ? r : 5u; }
// x = 255 q + r.  Runs are almost always shorter than 255 bytes: keep the 64-bit division off the common path.
BZ_HD void divmod255(uint64_t x, uint64_t &q, uint32_t &r)
{
    if (x < 255ull) { q = 0; r = (uint32_t)x; }
    else if (x <= 0xffffffffull) { const uint32_t x32 = (uint32_t)x; q = x32 / 255u; r = x32 % 255u; }
    else { q = x / 255ull; r = (uint32_t)(x % 255ull); }
}
// encoded size of a maximal run of `len` equal bytes chunked from its first byte: full 255-chunks (4+1 bytes) + rest
BZ_HD uint64_t rle_enc_run(uint64_t len)
{
    uint64_t q; uint32_t r;
    divmod255(len, q, r);
    return 5ull * q + rle_enc_small(r);
}

struct Seg {
    uint64_t len;   // bytes covered (0 = empty segment, the identity)
    uint64_t lead;  // length of the leading run of equal bytes (== len when uniform)
    uint64_t trail; // length of the trailing run of equal bytes (== len when uniform)
    uint64_t mid;   // encoded bytes of every maximal run that touches neither end of the segment
    uint32_t first, last;
    uint32_t uniform; // all bytes equal
    uint32_t pad;
};

BZ_HD Seg seg_empty()
{
    Seg s;
    s.len = s.lead = s.trail = s.mid = 0;
    s.first = s.last = 0;
    s.uniform = 1;
    s.pad = 0;
    return s;
}

BZ_HD Seg seg_byte_run(uint32_t v, uint64_t n)
{
    Seg s;
    s.len = s.lead = s.trail = n;
    s.mid = 0;
    s.first = s.last = v;
    s.uniform = 1;
    s.pad = 0;
    return s;
}

// A followed by B.
BZ_HD Seg seg_combine(const Seg &A, const Seg &B)
{
    if (A.len == 0) return B;
    if (B.len == 0) return A;
    Seg R;
    R.pad = 0;
    R.len = A.len + B.len;
    R.first = A.first;
    R.last = B.last;
    const bool join = (A.last == B.first);
    if (A.uniform && B.uniform && join) {
        R.uniform = 1;
        R.lead = R.trail = R.len;
        R.mid = 0;
        return R;
    }
    R.uniform = 0;
    R.mid = A.mid + B.mid;
    if (join) {
        R.lead = A.uniform ? A.len + B.lead : A.lead;
        R.trail = B.uniform ? A.trail + B.len : B.trail;
        if (!A.uniform && !B.uniform) R.mid += rle_enc_run(A.trail + B.lead);
    } else {
        R.lead = A.lead;   // == A.len when A is uniform
        R.trail = B.trail; // == B.len when B is uniform
        if (!A.uniform) R.mid += rle_enc_run(A.trail);
        if (!B.uniform) R.mid += rle_enc_run(B.lead);
    }
    return R;
}

// State of BlockCompressor between two write() calls.
struct RleState {
    uint64_t bl;  // blockLength
    uint32_t cur; // rleCurrentValue (only meaningful when run > 0)
    uint32_t run; // rleLength, 0..254
};

BZ_HD RleState state_fresh()
{
    RleState s;
    s.bl = 0;
    s.cur = 0xffffffffu;
    s.run = 0;
    return s;
}

// One accepted byte (BlockCompressor.hpp:76-95). The refusal test (:71-74) is the caller's.
BZ_HD void state_step(RleState &s, uint32_t v)
{
    if (s.run == 0) {
        s.cur = v;
        s.run = 1;
    } else if (s.cur == v) {
        if (++s.run > 254) {
            s.bl += 5;
            s.run = 0;
        }
    } else {
        s.bl += rle_enc_small(s.run);
        s.cur = v;
        s.run = 1;
    }
}

// State after accepting every byte of S (valid only while no refusal happens inside S; since blockLength is
// monotone, "bl after <= limit" proves that).
BZ_HD RleState state_apply(const Seg &S, RleState s)
{
    if (S.len == 0) return s;
    if (S.uniform) {
        uint64_t tot;
        if (s.run > 0 && s.cur == S.first) tot = s.run + S.len;
        else { s.bl += rle_enc_small(s.run); tot = S.len; }
        uint64_t q; uint32_t r;
        divmod255(tot, q, r);
        s.bl += 5ull * q;
        s.run = r;
        s.cur = S.first;
    } else {
        uint64_t tot;
        if (s.run > 0 && s.cur == S.first) tot = s.run + S.lead;
        else { s.bl += rle_enc_small(s.run); tot = S.lead; }
        uint64_t q; uint32_t r;
        divmod255(S.trail, q, r);
        s.bl += rle_enc_run(tot) + S.mid + 5ull * q;
        s.run = r;
        s.cur = S.last;
    }
    return s;
}

// blockLength after finishRLE() (BlockCompressor.hpp:112-118)
BZ_HD uint64_t state_final_length(const RleState &s) { return s.bl + rle_enc_small(s.run); }

// Summary of `n` bytes read through `get(i)`.
template <typename Get>
BZ_HD Seg seg_from_bytes(Get get, uint32_t n)
{
    if (n == 0) return seg_empty();
    Seg s;
    s.pad = 0;
    s.len = n;
    s.mid = 0;
    uint32_t cur = get(0);
    s.first = cur;
    uint32_t runlen = 1;
    bool changed = false;
    uint32_t lead = 0;
    for (uint32_t i = 1; i < n; ++i) {
        uint32_t v = get(i);
        if (v == cur) {
            ++runlen;
        } else {
            if (!changed) { lead = runlen; changed = true; }
            else s.mid += rle_enc_run(runlen);
            cur = v;
            runlen = 1;
        }
    }
    s.last = cur;
    if (!changed) {
        s.uniform = 1;
        s.lead = s.trail = n;
    } else {
        s.uniform = 0;
        s.lead = lead;
        s.trail = runlen;
    }
    return s;
}

// ---- geometry of the summary tree ----
constexpr uint32_t RLE_TILE = 2048;     // bytes per level-0 tile (one warp, 64 bytes per lane)
constexpr uint32_t RLE_LANE_BYTES = 64;
constexpr uint32_t RLE_FANOUT = 64;
constexpr int RLE_MAX_LEVELS = 6;       // 2 KiB * 64^5 = 2 PiB

} // namespace bz

namespace bz {

// Result of scanning bytes [lo, hi) of one tile with the machine in state `st` at lo.
struct TileScanResult {
    uint64_t pos;   // first q in [lo,hi) refused (st.bl > limit at entry), else hi
    RleState st;    // state at pos
};

// One block of the sequential walk: starting with a fresh machine at `start`, find where the block ends.
//   load(level, index)            -> Seg of tree node (level 0 = RLE_TILE bytes)
//   scan(lo, hi, state, limit)    -> TileScanResult (lo, hi inside one tile)
// Returns the end position e (== n_bytes, or the refused byte that starts the next block) and leaves the machine
// state at e in `st_end` and the state at the first tile boundary after `start` in `st_tile` (for the emit pass).
template <typename Load, typename Scan>
BZ_HD uint64_t walk_block(uint64_t start, uint64_t n_bytes, uint64_t limit, int n_levels, Load load, Scan scan, RleState &st_end)
{
    RleState st = state_fresh();
    uint64_t pos = start;
    // 1. finish the partial tile
    if (pos % RLE_TILE != 0) {
        uint64_t hi = (pos / RLE_TILE + 1) * RLE_TILE;
        if (hi > n_bytes) hi = n_bytes;
        TileScanResult r = scan(pos, hi, st, limit);
        if (r.pos < hi || r.st.bl > limit || hi == n_bytes) { st_end = r.st; return r.pos; }
        st = r.st;
        pos = hi;
    }
    // 2. climb: take the largest aligned node that fits; 3. once a node does not fit, only descend
    bool descending = false;
    int level = 0;
    while (pos < n_bytes) {
        if (!descending) {
            level = 0;
            uint64_t size = RLE_TILE;
            while (level + 1 < n_levels && pos % (size * RLE_FANOUT) == 0) { size *= RLE_FANOUT; ++level; }
        }
        uint64_t size = RLE_TILE;
        for (int l = 0; l < level; ++l) size *= RLE_FANOUT;
        uint64_t hi = pos + size;
        if (hi > n_bytes) hi = n_bytes;
        Seg node = load(level, pos / size);
        RleState nx = state_apply(node, st);
        if (nx.bl <= limit) {
            st = nx;
            pos = hi;
            continue;
        }
        if (level > 0) {
            descending = true;
            --level;
            continue;
        }
        TileScanResult r = scan(pos, hi, st, limit);
        st_end = r.st; // r.pos < hi, or refusal exactly at hi (r.st.bl > limit), or hi == n_bytes
        return r.pos;
    }
    st_end = st;
    return pos;
}

// Same walk, but every tree level is consumed a sibling group (<= 64 nodes) at a time so that a warp can evaluate the
// whole group with one coalesced load and one scan:
//   step(level, i0, iend, st, limit, found) -> true if pushing the machine through nodes [i0, iend) of `level` (at most 64)
//       overflows; then `found` is the first overflowing node and `st` the state in front of it; otherwise `st` is the state
//       behind node iend-1.
//   cnt[level] = number of nodes at that level.
template <typename Step, typename Scan>
BZ_HD uint64_t walk_block_wide(uint64_t start, uint64_t n_bytes, uint64_t limit, int n_levels, const uint64_t *cnt, Step step, Scan scan,
                               RleState &st_end)
{
    RleState st = state_fresh();
    uint64_t pos = start;
    if (pos % RLE_TILE != 0) {
        uint64_t hi = (pos / RLE_TILE + 1) * RLE_TILE;
        if (hi > n_bytes) hi = n_bytes;
        TileScanResult r = scan(pos, hi, st, limit);
        if (r.pos < hi || r.st.bl > limit || hi == n_bytes) { st_end = r.st; return r.pos; }
        st = r.st;
        pos = hi;
    }
    int level = 0;
    uint64_t found = 0;
    bool hit = false;
    while (!hit) {
        if (pos >= n_bytes) { st_end = st; return n_bytes; }
        level = 0;
        uint64_t i0 = pos / RLE_TILE;
        while (level + 1 < n_levels && (i0 % RLE_FANOUT) == 0) { ++level; i0 /= RLE_FANOUT; }
        uint64_t gend = (level == n_levels - 1) ? cnt[level] : (i0 / RLE_FANOUT + 1) * RLE_FANOUT;
        if (gend > cnt[level]) gend = cnt[level];
        uint64_t size = RLE_TILE;
        for (int l = 0; l < level; ++l) size *= RLE_FANOUT;
        while (i0 < gend && !hit) {
            uint64_t iend = i0 + RLE_FANOUT < gend ? i0 + RLE_FANOUT : gend;
            hit = step(level, i0, iend, st, limit, found);
            i0 = iend;
        }
        if (!hit) {
            pos = gend * size;
            if (pos > n_bytes) pos = n_bytes;
        }
    }
    while (level > 0) {
        --level;
        uint64_t i0 = found * RLE_FANOUT;
        uint64_t iend = i0 + RLE_FANOUT < cnt[level] ? i0 + RLE_FANOUT : cnt[level];
        step(level, i0, iend, st, limit, found); // must hit: the parent overflowed
    }
    pos = found * RLE_TILE;
    uint64_t hi = pos + RLE_TILE < n_bytes ? pos + RLE_TILE : n_bytes;
    TileScanResult r = scan(pos, hi, st, limit);
    st_end = r.st;
    return r.pos;
}

// ---- a full lane (64 bytes) from its equality mask ----
#if defined(__CUDA_ARCH__)
BZ_HD uint32_t bz_ffs64(uint64_t x) { return (uint32_t)__ffsll((long long)x); }
BZ_HD uint32_t bz_clz64(uint64_t x) { return (uint32_t)__clzll((long long)x); }
BZ_HD uint32_t bz_popc64(uint64_t x) { return (uint32_t)__popcll(x); }
#else
BZ_HD uint32_t bz_ffs64(uint64_t x) { return (uint32_t)__builtin_ffsll((long long)x); }
BZ_HD uint32_t bz_clz64(uint64_t x) { return x ? (uint32_t)__builtin_clzll(x) : 64u; }
BZ_HD uint32_t bz_popc64(uint64_t x) { return (uint32_t)__builtin_popcountll(x); }
#endif
// eq: bit j (1..63) set when byte j equals byte j-1, bit 0 clear.  Inside 64 bytes a run of L equal bytes is encoded as L bytes
// (L < 4) or as 5 (4 <= L <= 64), so the encoded length of all runs is 64 - sum over long runs of (L - 5); a long run shows as L - 3
// consecutive bits of eq & eq>>1 & eq>>2, hence sum (L - 3) = popcount of that mask and the number of long runs = its rising edges.
BZ_HD Seg seg_from_eq_mask(uint32_t first, uint32_t last, uint64_t eq)
{
    Seg s;
    s.pad = 0;
    s.len = RLE_LANE_BYTES;
    s.first = first;
    s.last = last;
    const uint32_t lead = bz_ffs64(~(eq >> 1)); // 64 when all bytes are equal (eq >> 1 has 63 ones, bit 63 is clear)
    if (lead == RLE_LANE_BYTES) {
        s.uniform = 1;
        s.lead = s.trail = RLE_LANE_BYTES;
        s.mid = 0;
        return s;
    }
    const uint32_t trail = 1u + bz_clz64(~eq);
    const uint64_t m3 = eq & (eq >> 1) & (eq >> 2);
    const uint32_t all = RLE_LANE_BYTES - bz_popc64(m3) + 2u * bz_popc64(m3 & ~(m3 << 1));
    s.uniform = 0;
    s.lead = lead;
    s.trail = trail;
    s.mid = all - (lead < 4u ? lead : 5u) - (trail < 4u ? trail : 5u);
    return s;
}

// ---- summaries of at most 4096 bytes in 32-bit fields (the walk's tile simulation scans 32 of them per block: two shuffles per step
// instead of eleven) ----
struct TSeg {
    uint32_t len, lead, trail, mid, first, last; // as in Seg; uniform <=> lead == len
};
BZ_HD TSeg tseg_of(const Seg &s)
{
    TSeg t;
    t.len = (uint32_t)s.len; t.lead = (uint32_t)s.lead; t.trail = (uint32_t)s.trail; t.mid = (uint32_t)s.mid; t.first = s.first; t.last = s.last;
    return t;
}
BZ_HD uint32_t rle_enc_run32(uint32_t len) { return 5u * (len / 255u) + rle_enc_small(len % 255u); }
BZ_HD TSeg tseg_combine(const TSeg &A, const TSeg &B)
{
    if (A.len == 0) return B;
    if (B.len == 0) return A;
    TSeg R;
    R.len = A.len + B.len;
    R.first = A.first;
    R.last = B.last;
    const bool au = A.lead == A.len, bu = B.lead == B.len, join = A.last == B.first;
    if (au && bu && join) { R.lead = R.trail = R.len; R.mid = 0; return R; }
    R.mid = A.mid + B.mid;
    if (join) {
        R.lead = au ? A.len + B.lead : A.lead;
        R.trail = bu ? A.trail + B.len : B.trail;
        if (!au && !bu) R.mid += rle_enc_run32(A.trail + B.lead);
    } else {
        R.lead = A.lead;
        R.trail = B.trail;
        if (!au) R.mid += rle_enc_run32(A.trail);
        if (!bu) R.mid += rle_enc_run32(B.lead);
    }
    return R;
}
// state_apply for a TSeg (the carried run is at most 254 long, the segment at most 4096)
BZ_HD RleState tstate_apply(const TSeg &S, RleState s)
{
    if (S.len == 0) return s;
    const bool uni = S.lead == S.len;
    uint32_t add = 0, tot;
    if (s.run > 0 && s.cur == S.first) tot = s.run + S.lead;
    else { add = rle_enc_small(s.run); tot = S.lead; }
    if (uni) {
        add += 5u * (tot / 255u);
        s.run = tot % 255u;
        s.cur = S.first;
    } else {
        add += rle_enc_run32(tot) + S.mid + 5u * (S.trail / 255u);
        s.run = S.trail % 255u;
        s.cur = S.last;
    }
    s.bl += add;
    return s;
}

// First refused byte of a segment in which NO run of four or more equal bytes forms, the run carried in by `in` included
// (promise of the caller): every flushed run then counts as its own length, so the machine's state in front of byte i is
// closed form -- bl = in.bl + in.run + i - (pending run), pending run = the 1..3 equal bytes ending at i-1 (+ in.run when they reach
// the segment's start and continue the carried run) -- and the first i with bl > limit is one of three neighbours.
//   get(j) = byte j of the segment (0 <= j < n); looks at positions i_min <= i < n; pos = n if no byte is refused.
// Returns the state in front of byte pos (pos == n: behind the segment).
template <typename Get>
BZ_HD RleState short_run_search(Get get, uint32_t n, const RleState &in, uint64_t limit, uint32_t i_min, uint32_t &pos)
{
    auto state_at = [&](uint32_t i) {
        if (i == 0) return in;
        const uint32_t v = get(i - 1);
        uint32_t r = 1;
        while (r < 3 && r < i && get(i - 1 - r) == v) ++r;
        if (r == i && in.run > 0 && in.cur == v) r += in.run;
        RleState st;
        st.bl = in.bl + in.run + i - r;
        st.cur = v;
        st.run = r;
        return st;
    };
    if (i_min == 0 && in.bl > limit) { pos = 0; return in; }
    // i - pending > K with pending in 1..3: the first such i is K+2, K+3 or K+4
    const int64_t K = (int64_t)limit - (int64_t)(in.bl + in.run);
    int64_t start = K + 2;
    if (start < 1) start = 1;
    if (start < (int64_t)i_min) start = (int64_t)i_min;
    for (int64_t i = start; i < start + 3 && i < (int64_t)n; ++i) {
        const RleState st = state_at((uint32_t)i);
        if (st.bl > limit) { pos = (uint32_t)i; return st; }
    }
    pos = n;
    return state_at(n);
}

// ---- the same walk on PREFIX states: one tile simulation per block instead of a tree descent ----
// Let the GLOBAL machine be the one that never starts a new block (it runs from byte 0), and pstate(t) its state at the first
// byte of tile t (a parallel scan of the tile summaries).  A block's fresh machine, started at s, and the global machine agree
// on (cur, run) as soon as the run that contains s has ended (both flush it when the next different byte arrives), and from
// then on their blockLengths differ by a constant D.  So "first byte refused by the fresh machine" = "first byte where the
// global blockLength exceeds limit + D": a search in the monotone pstate[].bl plus one exact tile simulation.
//   gstate       in: global state at s; out: global state at the returned position
//   head(s, v, r, vr)   v = in[s]; r = first q > s with in[q] != in[s], or n_bytes; vr = in[r] when r < n_bytes (one look at the
//                bytes behind s answers all three)
//   pstate(t)    global state at byte t * RLE_TILE for t <= n_tiles (pstate(n_tiles) = state at n_bytes)
//   first_over(t_lo, lim, before)  smallest t > t_lo with pstate(t).bl > lim, or n_tiles + 1 if none; when it is <= n_tiles,
//                before = pstate(t - 1) (the search has usually just read it)
//   scan(lo, hi, state, lim) as above (any state / limit)
// Returns the end position; st_end = FRESH machine state there (for the block's RLE1 length).
template <typename Head, typename PState, typename FirstOver, typename Scan>
BZ_HD uint64_t walk_block_prefix(uint64_t s, uint64_t n_bytes, uint64_t limit, RleState &gstate, Head head, PState pstate, FirstOver first_over, Scan scan,
                                 RleState &st_end)
{
    const uint64_t n_tiles = (n_bytes + RLE_TILE - 1) / RLE_TILE;
    uint32_t v, vr;
    uint64_t r;
    head(s, v, r, vr);
    const uint64_t L1 = r - s;
    // the block ends inside its first run: 255-byte chunks of 5 encoded bytes each
    const uint64_t k = limit / 5 + 1;
    if (255ull * k <= L1) {
        const uint64_t e = s + 255ull * k;
        st_end = state_apply(seg_byte_run(v, 255ull * k), state_fresh());
        gstate = state_apply(seg_byte_run(v, 255ull * k), gstate);
        return e;
    }
    const Seg first_run = seg_byte_run(v, L1);
    RleState f = state_apply(first_run, state_fresh());
    RleState g = state_apply(first_run, gstate);
    if (r >= n_bytes) { st_end = f; gstate = g; return n_bytes; }
    // both machines take byte r: the first run is flushed in both, (cur, run) agree from here on
    state_step(f, vr);
    state_step(g, vr);
    const uint64_t D = g.bl - f.bl;
    const uint64_t limD = limit + D;
    const uint64_t pos = r + 1; // next byte to enter, global state g
    if (g.bl > limD) { st_end = f; gstate = g; return pos; }
    if (pos >= n_bytes) { st_end = f; gstate = g; return n_bytes; }
    const uint64_t t = pos / RLE_TILE; // tile of pos
    // first tile boundary behind pos whose global blockLength is over the limit: the refusal is in the tile in front of it
    RleState before;
    const uint64_t ts = first_over(t, limD, before);
    if (ts > n_tiles) { // never: the block runs to the end of the input
        const RleState pe = pstate(n_tiles);
        gstate = pe; st_end = pe; st_end.bl -= D;
        return n_bytes;
    }
    const uint64_t hi = ts * RLE_TILE < n_bytes ? ts * RLE_TILE : n_bytes;
    if (ts == t + 1) { // the rest of the tile pos lies in
        if (pos < hi) {
            TileScanResult res = scan(pos, hi, g, limD);
            gstate = res.st;
            st_end = res.st; st_end.bl -= D;
            return res.pos; // refusal inside [pos, hi) or exactly at hi (then res.st.bl > limD)
        }
        const RleState pn = pstate(ts);
        gstate = pn; st_end = pn; st_end.bl -= D;
        return hi;
    }
    TileScanResult res = scan((ts - 1) * RLE_TILE, hi, before, limD);
    gstate = res.st;
    st_end = res.st; st_end.bl -= D;
    return res.pos;
}

} // namespace bz
