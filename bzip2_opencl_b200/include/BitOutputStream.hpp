// BitOutputStream.hpp -- facade of the reference's free bit-writer functions
// (/root/reference/include/BitOutputStream.hpp:30-135): same names, same signatures, same semantics on a bool-per-bit
// buffer with an external size_t count.  OutputStream only uses them for the few bits it writes itself (stream header,
// end marker, stream CRC, padding: OutputStream.hpp:126-128,169-172); the block bits arrive already packed from the GPU
// (csrc/emit.cuh) -- the reference's 16*blockSize bools per block (OutputStream.hpp:70,83) are gone.
#ifndef BIT_OUTPUT_STREAM_HPP
#define BIT_OUTPUT_STREAM_HPP

#include <ostream>
#include <stdexcept>
#include <vector>
#include <cstddef>

inline void writeBoolean(bool *bitBuffer, size_t *bitCount, bool value) { bitBuffer[(*bitCount)++] = value; }

inline void writeUnary(bool *bitBuffer, size_t *bitCount, int value)
{
    for (; value > 0; --value)
        writeBoolean(bitBuffer, bitCount, true);
    writeBoolean(bitBuffer, bitCount, false);
}

inline void writeBits(bool *bitBuffer, size_t *bitCount, int count, int value)
{
    for (int bit = count - 1; bit >= 0; --bit)
        writeBoolean(bitBuffer, bitCount, (static_cast<unsigned>(value) >> bit) & 1u);
}

inline void writeInteger(bool *bitBuffer, size_t *bitCount, int value)
{
    writeBits(bitBuffer, bitCount, 16, (value >> 16) & 0xffff);
    writeBits(bitBuffer, bitCount, 16, value & 0xffff);
}

inline void padding(bool *bitBuffer, size_t *bitCount)
{
    while (*bitCount % 8 != 0)
        writeBoolean(bitBuffer, bitCount, false);
}

// Takes the (< 8) bits left in the buffer as the carry for the next writeFileBytes call (reference :30-45).
inline std::vector<bool> getLeftBuffer(bool *bitBuffer, size_t *bitCount)
{
    if (*bitCount >= 8)
        throw std::runtime_error("Couldn't write all bytes properly during last step.");
    std::vector<bool> left(bitBuffer, bitBuffer + *bitCount);
    *bitCount = 0;
    return left;
}

// Writes every whole byte of (leftBuffer ++ bitBuffer) to `out`, MSB first, and keeps the remaining (< 8) bits at the
// front of bitBuffer (reference :47-99).
inline void writeFileBytes(bool *bitBuffer, size_t *bitCount, std::ostream &out, const std::vector<bool> &leftBuffer)
{
    const size_t carry = leftBuffer.size();
    const size_t total = carry + *bitCount;
    auto bitAt = [&](size_t i) -> bool { return i < carry ? leftBuffer[i] : bitBuffer[i - carry]; };
    const size_t whole = total / 8;
    for (size_t b = 0; b < whole; ++b)
    {
        unsigned char byte = 0;
        for (int j = 0; j < 8; ++j)
            byte = static_cast<unsigned char>((byte << 1) | (bitAt(b * 8 + j) ? 1 : 0));
        out.put(static_cast<char>(byte));
    }
    const size_t rest = total % 8;
    bool tail[8];
    for (size_t i = 0; i < rest; ++i)
        tail[i] = bitAt(whole * 8 + i);
    for (size_t i = 0; i < rest; ++i)
        bitBuffer[i] = tail[i];
    *bitCount = rest;
}
#endif
