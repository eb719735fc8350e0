// BlockCompressor.hpp -- facade of the reference's host-side BlockCompressor
// (/root/reference/include/BlockCompressor.hpp:35-155): same constructor and methods, same observable behaviour
// (RLE1 into the caller's buffer, presence flags for values and count bytes, CRC over the un-RLE'd bytes, refusal when
// blockLength > blockSize-6).  It exists so reference call sites keep compiling; OutputStream (the hot path) does NOT use
// it: there RLE1, presence and CRC run on the GPU (csrc/rle.cuh) and this per-byte loop disappears.
// Deviation: the vector<char> overload masks each byte to 0..255 instead of indexing with a negative char.
#ifndef BLOCK_COMPRESSOR_HPP
#define BLOCK_COMPRESSOR_HPP

#include <vector>
#include <cstdint>

#include "Config.hpp"
#include "CRC32.hpp"

class BlockCompressor
{
private:
    CRC32 crc{};
    unsigned char *block;
    bool *blockValuesPresent;
    int blockLengthLimit;
    int blockLength = 0;
    int rleCurrentValue = -1;
    int rleLength = 0;

    void writeRun(int value, int runLength) // reference :134-154
    {
        blockValuesPresent[value] = true;
        crc.updateCRC(value, runLength);
        const int copies = runLength < 4 ? runLength : 4;
        for (int i = 0; i < copies; ++i)
            block[blockLength++] = static_cast<unsigned char>(value);
        if (runLength >= 4)
        {
            blockValuesPresent[runLength - 4] = true;
            block[blockLength++] = static_cast<unsigned char>(runLength - 4);
        }
    }

public:
    BlockCompressor(unsigned char *blockPtr, bool *valuesPresentPtr, int blockSize)
        : block(blockPtr), blockValuesPresent(valuesPresentPtr), blockLengthLimit(blockSize - 6) {}

    bool isEmpty() { return blockLength == 0 && rleLength == 0; }
    int getCRC() const { return crc.getCRC(); }
    int getBlockLength() const { return blockLength; }

    bool write(int value) // reference :69-96
    {
        if (blockLength > blockLengthLimit)
            return false;
        if (rleLength == 0)
        {
            rleCurrentValue = value;
            rleLength = 1;
        }
        else if (rleCurrentValue != value)
        {
            writeRun(rleCurrentValue, rleLength);
            rleCurrentValue = value;
            rleLength = 1;
        }
        else if (++rleLength > 254)
        {
            writeRun(rleCurrentValue, 255);
            rleLength = 0;
        }
        return true;
    }

    int write(const std::vector<char> &data, int offset, int length) // reference :98-110
    {
        int written = 0;
        for (; written < length; ++written)
            if (!write(static_cast<unsigned char>(data[offset + written])))
                break;
        return written;
    }

    void finishRLE() // reference :112-118
    {
        if (rleLength > 0)
            writeRun(rleCurrentValue & 0xff, rleLength);
    }

    void reset() // reference :120-131
    {
        crc.reset();
        blockLength = 0;
        rleCurrentValue = -1;
        rleLength = 0;
        for (int i = 0; i < 256; ++i)
            blockValuesPresent[i] = false;
    }
};
#endif
