// Host check of the RLE1 segment algebra + block walk (bzip2_opencl_b200/csrc/rle_seg.h) against a literal
// byte-at-a-time run of the reference state machine (BlockCompressor.hpp:69-96).  Pure CPU; run by tests/test_host_logic.py.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstring>
#include "../../bzip2_opencl_b200/csrc/rle_seg.h"
using namespace bz;

static std::vector<Seg> level0(const std::vector<uint8_t> &in)
{
    size_t nt = (in.size() + RLE_TILE - 1) / RLE_TILE;
    std::vector<Seg> out(nt);
    for (size_t t = 0; t < nt; ++t) {
        Seg acc = seg_empty();
        for (uint32_t lane = 0; lane < 32; ++lane) {
            size_t lo = t * RLE_TILE + lane * RLE_LANE_BYTES, hi = lo + RLE_LANE_BYTES;
            if (hi > in.size()) hi = in.size();
            if (lo >= hi) break;
            const uint8_t *p = in.data() + lo;
            acc = seg_combine(acc, seg_from_bytes([&](uint32_t i) { return (uint32_t)p[i]; }, (uint32_t)(hi - lo)));
        }
        out[t] = acc;
    }
    return out;
}

struct Blocks { std::vector<uint64_t> starts, lens; };

static Blocks reference_walk(const std::vector<uint8_t> &in, uint64_t limit)
{
    Blocks b;
    RleState st = state_fresh();
    uint64_t start = 0;
    for (size_t i = 0; i < in.size(); ++i) {
        if (st.bl > limit) {
            b.starts.push_back(start); b.lens.push_back(state_final_length(st));
            st = state_fresh(); start = i;
        }
        state_step(st, in[i]);
    }
    if (!(st.bl == 0 && st.run == 0)) { b.starts.push_back(start); b.lens.push_back(state_final_length(st)); }
    return b;
}

static Blocks tree_walk(const std::vector<uint8_t> &in, uint64_t limit)
{
    std::vector<std::vector<Seg>> lv;
    lv.push_back(level0(in));
    while (lv.back().size() > 1 && (int)lv.size() < RLE_MAX_LEVELS) {
        const auto &c = lv.back();
        std::vector<Seg> up((c.size() + RLE_FANOUT - 1) / RLE_FANOUT);
        for (size_t j = 0; j < up.size(); ++j) {
            Seg acc = seg_empty();
            for (size_t k = j * RLE_FANOUT; k < (j + 1) * RLE_FANOUT && k < c.size(); ++k) acc = seg_combine(acc, c[k]);
            up[j] = acc;
        }
        lv.push_back(up);
    }
    auto load = [&](int level, uint64_t idx) { return lv[level][idx]; };
    auto scan = [&](uint64_t lo, uint64_t hi, RleState st, uint64_t lim) {
        // emulate the warp: 32 lanes x 64 bytes from lo
        Seg segs[32]; RleState sts[32];
        Seg acc = seg_empty();
        TileScanResult res; res.pos = hi;
        bool found = false;
        for (uint32_t lane = 0; lane < 32; ++lane) {
            uint64_t a = lo + lane * RLE_LANE_BYTES, e = a + RLE_LANE_BYTES;
            if (e > hi) e = hi;
            sts[lane] = state_apply(acc, st);
            if (a < e) {
                const uint8_t *p = in.data() + a;
                segs[lane] = seg_from_bytes([&](uint32_t i) { return (uint32_t)p[i]; }, (uint32_t)(e - a));
                RleState s = sts[lane];
                for (uint64_t q = a; q < e && !found; ++q) {
                    if (s.bl > lim) { found = true; res.pos = q; res.st = s; break; }
                    state_step(s, in[q]);
                }
                // consistency of apply vs step
                if (!found) {
                    RleState viaSeg = state_apply(segs[lane], sts[lane]);
                    if (viaSeg.bl != s.bl || viaSeg.run != s.run || (s.run && viaSeg.cur != s.cur)) { printf("apply/step mismatch\n"); exit(2); }
                }
                acc = seg_combine(acc, segs[lane]);
            }
        }
        if (!found) res.st = state_apply(acc, st);
        return res;
    };
    std::vector<uint64_t> cnt;
    for (auto &v : lv) cnt.push_back(v.size());
    auto step = [&](int level, uint64_t i0, uint64_t iend, RleState &st, uint64_t lim, uint64_t &found) {
        for (uint64_t i = i0; i < iend; ++i) {
            RleState nx = state_apply(lv[level][i], st);
            if (nx.bl > lim) { found = i; return true; }
            st = nx;
        }
        return false;
    };
    Blocks b;
    uint64_t s = 0, n = in.size();
    while (s < n) {
        RleState st, st2;
        uint64_t e = walk_block(s, n, limit, (int)lv.size(), load, scan, st);
        uint64_t e2 = walk_block_wide(s, n, limit, (int)lv.size(), cnt.data(), step, scan, st2);
        if (e != e2 || state_final_length(st) != state_final_length(st2)) { printf("walk_block_wide disagrees at start %llu: %llu vs %llu\n", (unsigned long long)s, (unsigned long long)e, (unsigned long long)e2); exit(3); }
        b.starts.push_back(s); b.lens.push_back(state_final_length(st));
        s = e;
    }
    return b;
}

// the prefix-state walk (walk_block_prefix) with literal host callbacks
static Blocks prefix_walk(const std::vector<uint8_t> &in, uint64_t limit)
{
    const uint64_t n = in.size();
    const uint64_t nt = (n + RLE_TILE - 1) / RLE_TILE;
    std::vector<Seg> l0 = level0(in);
    std::vector<RleState> ps(nt + 1);
    {
        Seg acc = seg_empty();
        for (uint64_t t = 0; t <= nt; ++t) {
            ps[t] = state_apply(acc, state_fresh());
            if (t < nt) acc = seg_combine(acc, l0[t]);
        }
    }
    auto head = [&](uint64_t s, uint32_t &v, uint64_t &r, uint32_t &vr) {
        v = in[s];
        uint64_t q = s + 1;
        while (q < n && in[q] == in[s]) ++q;
        r = q;
        vr = q < n ? in[q] : 0u;
    };
    auto pstate = [&](uint64_t t) { return ps[t]; };
    auto first_over = [&](uint64_t t_lo, uint64_t lim, RleState &before) {
        uint64_t t = t_lo + 1;
        while (t <= nt && ps[t].bl <= lim) ++t;
        if (t <= nt) before = ps[t - 1];
        return t;
    };
    auto scan = [&](uint64_t lo, uint64_t hi, RleState st, uint64_t lim) {
        TileScanResult res; res.pos = hi;
        RleState s = st;
        for (uint64_t q = lo; q < hi; ++q) {
            if (s.bl > lim) { res.pos = q; break; }
            state_step(s, in[q]);
        }
        res.st = s;
        return res;
    };
    Blocks b;
    uint64_t s = 0;
    RleState g = state_fresh();
    while (s < n) {
        RleState st;
        uint64_t e = walk_block_prefix(s, n, limit, g, head, pstate, first_over, scan, st);
        b.starts.push_back(s); b.lens.push_back(state_final_length(st));
        if (e <= s) { printf("prefix walk made no progress at %llu\n", (unsigned long long)s); exit(4); }
        s = e;
    }
    return b;
}

static uint64_t rng = 88172645463325252ull;
static uint32_t rnd();

// TSeg (32-bit summaries) against Seg, and the closed-form search against the literal machine
static void check_tseg_and_short_runs()
{
    for (int iter = 0; iter < 20000; ++iter) {
        const uint32_t n = rnd() % 300 + 1;
        std::vector<uint8_t> b(n);
        const int mode = iter % 4;
        for (uint32_t i = 0; i < n;) {
            uint32_t v = rnd() & (mode == 0 ? 1 : 0xff), run = mode == 3 ? rnd() % 3 + 1 : (mode == 2 ? rnd() % 600 + 1 : rnd() % 5 + 1);
            for (uint32_t k = 0; k < run && i < n; ++k) b[i++] = (uint8_t)v;
        }
        // split in up to 5 pieces, combine as Seg and as TSeg
        Seg acc = seg_empty();
        TSeg tacc = tseg_of(seg_empty());
        uint32_t a = 0;
        while (a < n) {
            uint32_t e = a + rnd() % (n - a) + 1;
            const uint8_t *p = b.data() + a;
            Seg sg = seg_from_bytes([&](uint32_t i) { return (uint32_t)p[i]; }, e - a);
            acc = seg_combine(acc, sg);
            tacc = tseg_combine(tacc, tseg_of(sg));
            a = e;
        }
        RleState in = state_fresh();
        in.bl = rnd() % 100000;
        if (rnd() % 2) { in.run = rnd() % 254 + 1; in.cur = (rnd() % 2) ? b[0] : (rnd() & 0xff); }
        RleState x = state_apply(acc, in), y = tstate_apply(tacc, in);
        if (x.bl != y.bl || x.run != y.run || (x.run && x.cur != y.cur) || tacc.len != acc.len || tacc.mid != acc.mid) { printf("TSeg mismatch iter %d\n", iter); exit(5); }
        // closed-form search where the promise holds
        RleState in2 = in;
        if (in2.run > 3) in2.run = rnd() % 4;
        bool longrun = false;
        {
            uint32_t r = (in2.run > 0 && in2.cur == b[0]) ? in2.run + 1 : 1;
            if (r >= 4) longrun = true;
            for (uint32_t i = 1; i < n; ++i) { r = b[i] == b[i - 1] ? r + 1 : 1; if (r >= 4) longrun = true; }
        }
        if (!longrun) {
            for (int rep = 0; rep < 3; ++rep) {
                const uint64_t limit = in2.bl + (rep == 0 ? 0 : rnd() % (n + 8)) - (rep == 2 && in2.bl > 3 ? rnd() % 3 : 0);
                const uint32_t i_min = rnd() % 3 == 0 ? 1 : 0;
                // literal
                RleState st = in2;
                uint32_t lit = n;
                RleState lit_st;
                bool found = false;
                for (uint32_t i = 0; i < n; ++i) {
                    if (i >= i_min && st.bl > limit) { lit = i; lit_st = st; found = true; break; }
                    state_step(st, b[i]);
                }
                if (!found) lit_st = st;
                uint32_t pos;
                RleState got = short_run_search([&](uint32_t j) { return (uint32_t)b[j]; }, n, in2, limit, i_min, pos);
                if (pos != lit || got.bl != lit_st.bl || got.run != lit_st.run || (got.run && got.cur != lit_st.cur)) {
                    printf("short_run_search mismatch iter %d rep %d: pos %u vs %u, bl %llu vs %llu, run %u vs %u\n", iter, rep, pos, lit, (unsigned long long)got.bl,
                           (unsigned long long)lit_st.bl, got.run, lit_st.run);
                    exit(6);
                }
            }
        }
    }
}

static uint32_t rnd() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return (uint32_t)(rng >> 11); }

// a full lane's summary from its equality mask (seg_from_eq_mask) against the byte-by-byte summary
static void check_eq_mask_summary()
{
    for (int it = 0; it < 300000; ++it) {
        uint8_t p[64];
        const int mode = it % 4;
        for (int i = 0; i < 64;) {
            const uint32_t v = rnd() & (mode == 0 ? 1 : 255), run = mode == 1 ? rnd() % 3 + 1 : (mode == 2 ? rnd() % 70 + 1 : rnd() % 7 + 1);
            for (uint32_t k = 0; k < run && i < 64; ++k) p[i++] = (uint8_t)v;
        }
        uint64_t eq = 0;
        for (int j = 1; j < 64; ++j)
            if (p[j] == p[j - 1]) eq |= 1ull << j;
        const Seg ref = seg_from_bytes([&](uint32_t i) { return (uint32_t)p[i]; }, 64), s = seg_from_eq_mask(p[0], p[63], eq);
        if (s.uniform != ref.uniform || s.lead != ref.lead || s.trail != ref.trail || s.mid != ref.mid || s.first != ref.first || s.last != ref.last) { printf("eq-mask summary mismatch %d\n", it); exit(7); }
    }
}

int main()
{
    check_eq_mask_summary();
    check_tseg_and_short_runs();
    int cases = 0;
    for (int iter = 0; iter < 400; ++iter) {
        size_t n = rnd() % 300000 + 1;
        if (iter % 7 == 0) n = rnd() % 5000 + 1;
        if (iter % 50 == 3) n = 9000000 + rnd() % 12000000; // deep tree: 4 levels
        std::vector<uint8_t> in(n);
        int mode = iter % 6;
        size_t i = 0;
        while (i < n) {
            uint32_t v = rnd() & 0xff;
            size_t run;
            switch (mode) {
            case 0: run = 1; break;
            case 1: run = rnd() % 6 + 1; break;
            case 2: run = rnd() % 600 + 1; break;
            case 3: run = (rnd() % 3 == 0) ? rnd() % 70000 + 1 : rnd() % 4 + 1; break;
            case 4: run = n; v = 0; break;
            default: run = (rnd() % 50 == 0) ? 255 * (rnd() % 5 + 1) + (rnd() % 3) - 1 : rnd() % 9 + 1; break;
            }
            if (mode == 1) v &= 3;
            for (size_t k = 0; k < run && i < n; ++k) in[i++] = (uint8_t)v;
        }
        uint64_t bss[] = {20, 100, 1000, 10000, 90000, 300000};
        for (uint64_t bs : bss) {
            uint64_t limit = bs - 6;
            Blocks a = reference_walk(in, limit), b = tree_walk(in, limit), p = prefix_walk(in, limit);
            if (a.starts != p.starts || a.lens != p.lens) {
                printf("PREFIX MISMATCH iter %d mode %d n %zu bs %llu: ref %zu blocks, prefix %zu blocks\n", iter, mode, n, (unsigned long long)bs, a.starts.size(), p.starts.size());
                for (size_t k = 0; k < a.starts.size() && k < p.starts.size(); ++k)
                    if (a.starts[k] != p.starts[k] || a.lens[k] != p.lens[k]) { printf(" first diff at block %zu: ref (%llu,%llu) prefix (%llu,%llu)\n", k, (unsigned long long)a.starts[k], (unsigned long long)a.lens[k], (unsigned long long)p.starts[k], (unsigned long long)p.lens[k]); break; }
                return 1;
            }
            if (a.starts != b.starts || a.lens != b.lens) {
                printf("MISMATCH iter %d mode %d n %zu bs %llu: ref %zu blocks, tree %zu blocks\n", iter, mode, n, (unsigned long long)bs, a.starts.size(), b.starts.size());
                for (size_t k = 0; k < a.starts.size() && k < b.starts.size(); ++k)
                    if (a.starts[k] != b.starts[k] || a.lens[k] != b.lens[k]) { printf(" first diff at block %zu: ref (%llu,%llu) tree (%llu,%llu)\n", k, (unsigned long long)a.starts[k], (unsigned long long)a.lens[k], (unsigned long long)b.starts[k], (unsigned long long)b.lens[k]); break; }
                return 1;
            }
            ++cases;
        }
    }
    printf("OK %d cases\n", cases);
    return 0;
}
