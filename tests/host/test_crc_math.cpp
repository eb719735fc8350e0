// Host check of the CRC combine arithmetic used by k_crc_chunks / k_crc_finish (csrc/crc_math.h, rle.cuh).
#include <cstdio>
#include <vector>
#include <cstdlib>
#include "../../bzip2_opencl_b200/csrc/crc_math.h"
using namespace bz;
int main()
{
    uint32_t tab[256], pow8[48], lane_shift[256];
    crc_build_table(tab);
    crc_build_pow8(pow8, 48);
    auto P = [&](int j) { return pow8[j]; };
    for (int t = 0; t < 256; ++t) lane_shift[t] = crc_xpow8_t(64ull * (255 - t), P);
    if (tab[1] != 0x04c11db7u || tab[255] != 0xb1f740b4u) { printf("table mismatch\n"); return 1; }
    uint64_t rng = 12345;
    for (int iter = 0; iter < 60; ++iter) {
        rng = rng * 6364136223846793005ull + 1442695040888963407ull;
        size_t n = (size_t)((rng >> 33) % 200000) + 1, off = (size_t)((rng >> 20) % 77);
        std::vector<uint8_t> d(off + n);
        for (auto &b : d) { rng = rng * 6364136223846793005ull + 1442695040888963407ull; b = (uint8_t)(rng >> 56); }
        uint64_t s = off, e = off + n;
        uint32_t ref = 0xffffffffu;
        for (uint64_t q = s; q < e; ++q) ref = (ref << 8) ^ tab[((ref >> 24) ^ d[q]) & 0xff];
        ref = ~ref;
        const int CH = 256 * 64;
        uint32_t acc = 0;
        for (uint64_t c = 0; c * CH < e - s; ++c) {
            uint64_t chunk_end = e - c * CH;
            uint32_t x = 0;
            for (int t = 0; t < 256; ++t) {
                int64_t hi = (int64_t)chunk_end - 64 * (255 - t), lo = hi - 64;
                if (lo < (int64_t)s) lo = (int64_t)s;
                uint32_t crc = 0;
                for (int64_t q = lo; q < hi; ++q) crc = (crc << 8) ^ tab[((crc >> 24) ^ d[q]) & 0xff];
                if (crc) crc = crc_mulmod(crc, lane_shift[t]);
                x ^= crc;
            }
            if (x) acc ^= crc_mulmod(x, crc_xpow8_t(c * CH, P));
        }
        uint32_t got = ~(acc ^ crc_mulmod(0xffffffffu, crc_xpow8_t(e - s, P)));
        if (got != ref) { printf("MISMATCH n=%zu got %08x ref %08x\n", n, got, ref); return 1; }
    }
    printf("OK\n");
    return 0;
}
