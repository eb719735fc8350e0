// Host check of the round-0 key arithmetic (csrc/packed_math.h) used by k_bwt_count_packed and the first sweep of k_radix_onesweep:
// for every alphabet size class, the 40-bit keys of all cyclic rotations of a random text must (1) order the rotations by their
// first h0 bytes, (2) never contradict the full rotation order (the partial symbol only refines), (3) come out of the sliding
// window exactly as from a direct evaluation.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../bzip2_opencl_b200/csrc/packed_math.h"
using namespace bz;

static uint64_t rng_state = 20261018;
static uint32_t rnd() { rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(rng_state >> 33); }

int main()
{
    const int alphas[] = {1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 32, 33, 63, 64, 65, 127, 128, 129, 200, 255, 256};
    for (int alpha : alphas) {
        const PackedShape s = packed_shape((uint32_t)alpha);
        int want_bits = 1;
        while ((1 << want_bits) < alpha) ++want_bits;
        if (s.bits != want_bits || s.width != s.hc * s.bits || s.width > 47 || s.hc > PK_MAX_SYMS || s.h0 > s.hc || s.h0 < 5 ||
            s.h0 != (40 / s.bits < 16 ? 40 / s.bits : 16) || (s.hc < PK_MAX_SYMS && s.width < PK_KEY_BITS)) {
            printf("shape wrong for alpha %d: bits %d hc %d width %d h0 %d\n", alpha, s.bits, s.hc, s.width, s.h0);
            return 1;
        }
        // used byte values -> dense map (byte order), as packed_setup builds it from the presence bitmap
        std::vector<int> used(256, 0);
        for (int k = 0; k < alpha;) { int v = (int)(rnd() & 255u); if (!used[v]) { used[v] = 1; ++k; } }
        int map[256], r = 0;
        for (int v = 0; v < 256; ++v) { map[v] = r; r += used[v]; }
        std::vector<int> vals;
        for (int v = 0; v < 256; ++v) if (used[v]) vals.push_back(v);
        for (int rep = 0; rep < 3; ++rep) {
            const int n = rep == 0 ? 7 : 150 + (int)(rnd() % 90u); // a block shorter than a key wraps around several times
            std::vector<uint8_t> T(n);
            for (auto &b : T) b = (uint8_t)vals[rnd() % vals.size()];
            for (int k = 40; k + 30 < n; ++k) T[k + 25] = T[k - 30 + (k % 3)]; // repeats: ties on the first bytes
            std::vector<uint64_t> key(n);
            // sliding window over consecutive rotations, as packed_keys16 does
            uint64_t c = 0;
            for (int j = 0; j + 1 < s.hc; ++j) c = (c << s.bits) | (uint64_t)map[T[j % n]];
            for (int i = 0; i < n; ++i) {
                c = packed_push(c, (uint32_t)map[T[(i + s.hc - 1) % n]], s);
                key[i] = packed_key_of(c, s);
                uint64_t d = 0; // direct
                for (int j = 0; j < s.hc; ++j) d = (d << s.bits) | (uint64_t)map[T[(i + j) % n]];
                if (packed_key_of(d, s) != key[i] || (key[i] >> PK_KEY_BITS)) { printf("window mismatch alpha %d i %d\n", alpha, i); return 1; }
            }
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) {
                    int ch = 0, cf = 0; // rotation order on h0 bytes / on all n bytes
                    for (int k = 0; k < n && !cf; ++k) {
                        const int a = T[(i + k) % n], b = T[(j + k) % n];
                        if (a != b) { cf = a < b ? -1 : 1; if (k < s.h0) ch = cf; }
                    }
                    const int ck = key[i] < key[j] ? -1 : (key[i] > key[j] ? 1 : 0);
                    if (ch != 0 && ck != ch) { printf("alpha %d: keys do not order the first %d bytes (i %d j %d)\n", alpha, s.h0, i, j); return 1; }
                    if (ck != 0 && ck != cf) { printf("alpha %d: key order contradicts rotation order (i %d j %d)\n", alpha, i, j); return 1; }
                }
        }
    }
    printf("OK\n");
    return 0;
}
