// crc_math.h -- GF(2) arithmetic for combining bzip2 CRC32s (CRC32.hpp:30-92: MSB-first, polynomial 0x04C11DB7).
// The CRC register after n bytes from start value I is  I*x^(8n) + sum_i b_i * x^(32 + 8(n-1-i))  (mod P), so CRCs of
// chunks computed independently (from I = 0) combine with multiplications by x^(8k) mod P.  Host + device.
#pragma once
#include <stdint.h>
#if defined(__CUDACC__)
#define BZ_HD2 __host__ __device__ __forceinline__
#else
#define BZ_HD2 inline
#endif
namespace bz {
constexpr uint32_t CRC_POLY = 0x04C11DB7u;
BZ_HD2 uint32_t crc_mulmod(uint32_t a, uint32_t b)
{
    uint32_t r = 0;
    for (int i = 31; i >= 0; --i) {
        r = (r << 1) ^ ((r & 0x80000000u) ? CRC_POLY : 0u);
        if ((b >> i) & 1u) r ^= a;
    }
    return r;
}
// pow8[j] = x^(8 * 2^j) mod P
inline void crc_build_pow8(uint32_t *pow8, int n)
{
    pow8[0] = 0x100u;
    for (int j = 1; j < n; ++j) pow8[j] = crc_mulmod(pow8[j - 1], pow8[j - 1]);
}
template <typename Pow>
BZ_HD2 uint32_t crc_xpow8_t(uint64_t k, Pow pow8) // x^(8k) mod P
{
    uint32_t r = 1;
    for (int j = 0; k; ++j, k >>= 1)
        if (k & 1) r = crc_mulmod(r, pow8(j));
    return r;
}
inline uint32_t crc_xpow8_host(uint64_t k, const uint32_t *pow8) // x^(8k) mod P, host tables
{
    uint32_t r = 1;
    for (int j = 0; k; ++j, k >>= 1)
        if (k & 1) r = crc_mulmod(r, pow8[j]);
    return r;
}
inline void crc_build_table(uint32_t *table)
{
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i << 24;
        for (int k = 0; k < 8; ++k) c = (c & 0x80000000u) ? (c << 1) ^ CRC_POLY : (c << 1);
        table[i] = c;
    }
}
} // namespace bz
