// CRC32.hpp -- facade of the reference's CRC32 class (/root/reference/include/CRC32.hpp:30-92): same public surface
// (getCRC / updateCRC(int) / updateCRC(int,int) / reset) and the same arithmetic: MSB-first CRC-32, polynomial
// 0x04C11DB7, register starts at 0xFFFFFFFF, result is the complement.  The 256-entry table the reference spells out
// (:33-65) is generated at compile time here.  Host utility only; block CRCs on the hot path come from the GPU
// (k_crc_chunks) through bz2b200_compress_blocks.
#ifndef CRC32_HPP
#define CRC32_HPP

#include <array>
#include <cstdint>

namespace bz2b200_detail
{
    constexpr std::array<uint32_t, 256> makeCrcTable()
    {
        std::array<uint32_t, 256> t{};
        for (uint32_t i = 0; i < 256; ++i)
        {
            uint32_t c = i << 24;
            for (int k = 0; k < 8; ++k)
                c = (c & 0x80000000u) ? (c << 1) ^ 0x04c11db7u : (c << 1);
            t[i] = c;
        }
        return t;
    }
}

class CRC32
{
private:
    static constexpr std::array<uint32_t, 256> Crc32Lookup = bz2b200_detail::makeCrcTable();
    uint32_t crc = 0xffffffffu;

public:
    int getCRC() const { return static_cast<int>(~crc); }
    void updateCRC(int value) { crc = (crc << 8) ^ Crc32Lookup[((crc >> 24) ^ static_cast<uint32_t>(value)) & 0xff]; }
    void updateCRC(int value, int count)
    {
        while (count-- > 0)
            updateCRC(value);
    }
    void reset() { crc = 0xffffffffu; }
};
#endif
