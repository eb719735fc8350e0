// OutputStream.hpp -- facade of the reference's OutputStream (/root/reference/include/OutputStream.hpp:35-241) on top
// of the C ABI in include/bz2b200.h.  Same constructor, write(int), write(vector,off,len), close(); same exceptions
// (std::invalid_argument "Invalid block size" / "Invalid parallel block count", :73-81; std::runtime_error
// "Write beyond end of stream", :133-136).  An unmodified app.cpp compress branch (app.cpp:93-125) compiles against it.
//
// What changed underneath: the reference pushes every byte through a host BlockCompressor and ships 16 bool-bytes per
// input byte to and from the device each batch (:131-142,190-240).  Here write() only appends to a host staging buffer;
// when `parallelBlockCnt` blocks' worth of input has accumulated (and at close()) the raw bytes go to
// bz2b200_compress_blocks, which finds the block boundaries, does RLE1/CRC/BWT/MTF/Huffman on the GPU and returns the
// packed bit stream of all complete blocks plus how many input bytes they consumed.  The carry of < 8 bits between calls
// is the reference's leftBuffer (:225-239); header, end marker, folded stream CRC and padding stay here (:126-128,169-172,202).
#ifndef OUTPUT_STREAM_HPP
#define OUTPUT_STREAM_HPP

#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>
#include <cstdint>
#include <cstring>

#include "Config.hpp"
#include "BitOutputStream.hpp"
#include "BlockCompressor.hpp"
#include "bz2b200.h"

class OutputStream
{
private:
    std::ostream &outputStream;
    bool streamFinished = false;
    int streamBlockSize;
    int parallelBlockCnt;
    int streamCRC = 0;
    bz2b200_t *handle = nullptr;
    std::vector<uint8_t> staging;   // raw input not yet handed to the device
    size_t flushThreshold;
    std::vector<uint8_t> outBits;
    std::vector<uint32_t> blockCRCs;
    bool carryBits[8];              // < 8 bits waiting for the next byte (the reference's leftBuffer)
    size_t carryCount = 0;

    [[noreturn]] void fail(int rc)
    {
        std::string msg = std::string(bz2b200_strerror(rc)) + ": " + bz2b200_last_error(handle);
        throw std::runtime_error(msg);
    }

    // closeBlocks(), reference :190-240: every complete block of the staged input -> bits -> outputStream
    void closeBlocks(bool final)
    {
        if (staging.empty() && !final)
            return;
        const size_t cap = bz2b200_max_output_bytes(handle, staging.size()) + 16;
        if (outBits.size() < cap)
            outBits.resize(cap);
        const size_t maxBlocks = staging.size() / static_cast<size_t>((streamBlockSize - 5) * 4 / 5 > 0 ? (streamBlockSize - 5) * 4 / 5 : 1) + 4;
        if (blockCRCs.size() < maxBlocks)
            blockCRCs.resize(maxBlocks);
        uint64_t nbits = 0, nblocks = 0;
        size_t consumed = 0;
        int rc = bz2b200_compress_blocks(handle, staging.data(), staging.size(), final ? 1 : 0, static_cast<unsigned>(carryCount), outBits.data(),
                                         outBits.size(), &nbits, &consumed, blockCRCs.data(), nullptr, blockCRCs.size(), &nblocks);
        if (rc != BZ2B200_OK)
            fail(rc);
        for (uint64_t i = 0; i < nblocks; ++i) // reference :201-202
            streamCRC = ((streamCRC << 1) | (static_cast<unsigned int>(streamCRC) >> 31)) ^ static_cast<int>(blockCRCs[i]);
        if (nbits)
        {
            // OR the carried bits into the top of the first byte, write whole bytes, keep the new carry
            uint8_t first = 0;
            for (size_t i = 0; i < carryCount; ++i)
                first = static_cast<uint8_t>(first | ((carryBits[i] ? 1 : 0) << (7 - i)));
            outBits[0] = static_cast<uint8_t>(outBits[0] | first);
            const uint64_t total = carryCount + nbits;
            const size_t whole = static_cast<size_t>(total / 8);
            outputStream.write(reinterpret_cast<const char *>(outBits.data()), static_cast<std::streamsize>(whole));
            carryCount = static_cast<size_t>(total % 8);
            for (size_t i = 0; i < carryCount; ++i)
                carryBits[i] = (outBits[whole] >> (7 - i)) & 1;
        }
        staging.erase(staging.begin(), staging.begin() + static_cast<std::ptrdiff_t>(consumed));
    }

public:
    OutputStream(std::ostream &out, int blockSizeMultiplier, int parallelBlockCnt)
        : outputStream(out), streamBlockSize(BLOCKSIZE_DEFAULT * blockSizeMultiplier), parallelBlockCnt(parallelBlockCnt)
    {
        if (blockSizeMultiplier < 1 || blockSizeMultiplier > 9)
            throw std::invalid_argument("Invalid block size");
        if (parallelBlockCnt < 1)
            throw std::invalid_argument("Invalid parallel block count");
        int rc = bz2b200_create(&handle, 0, BLOCKSIZE_DEFAULT, blockSizeMultiplier, parallelBlockCnt);
        if (rc == BZ2B200_E_BLOCKSIZE)
            throw std::invalid_argument("Invalid block size");
        if (rc == BZ2B200_E_PARALLEL)
            throw std::invalid_argument("Invalid parallel block count");
        if (rc != BZ2B200_OK)
            throw std::runtime_error(std::string("bz2b200_create: ") + bz2b200_strerror(rc) + " (a CUDA device is required; there is no CPU fallback)");
        flushThreshold = static_cast<size_t>(streamBlockSize) * static_cast<size_t>(parallelBlockCnt) + static_cast<size_t>(streamBlockSize);
        // stream header, reference :126-128
        const char hdr[4] = {'B', 'Z', 'h', static_cast<char>('0' + blockSizeMultiplier)};
        outputStream.write(hdr, 4);
    }

    OutputStream(const OutputStream &) = delete;
    OutputStream &operator=(const OutputStream &) = delete;

    ~OutputStream() { bz2b200_destroy(handle); }

    void write(int value) // reference :131-142
    {
        if (streamFinished)
            throw std::runtime_error("Write beyond end of stream");
        staging.push_back(static_cast<uint8_t>(value & 0xff));
        if (staging.size() >= flushThreshold)
            closeBlocks(false);
    }

    void write(const std::vector<char> &data, int offset, int length) // reference :144-161
    {
        if (streamFinished)
            throw std::runtime_error("Write beyond end of stream");
        const uint8_t *p = reinterpret_cast<const uint8_t *>(data.data()) + offset;
        staging.insert(staging.end(), p, p + length);
        if (staging.size() >= flushThreshold)
            closeBlocks(false);
    }

    void close() // reference :163-176
    {
        if (streamFinished)
            return;
        streamFinished = true;
        closeBlocks(true);
        bool bits[8 + 48 + 32 + 8];
        size_t cnt = 0;
        for (size_t i = 0; i < carryCount; ++i)
            writeBoolean(bits, &cnt, carryBits[i]);
        carryCount = 0;
        writeBits(bits, &cnt, 24, STREAM_END_MARKER_1);
        writeBits(bits, &cnt, 24, STREAM_END_MARKER_2);
        writeInteger(bits, &cnt, streamCRC);
        padding(bits, &cnt);
        writeFileBytes(bits, &cnt, outputStream, {});
        outputStream.flush();
    }
};
#endif
