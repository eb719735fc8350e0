// InputStream.hpp -- facade of the reference's InputStream (/root/reference/include/InputStream.hpp:36-159) on top of the
// C ABI in include/bz2b200.h.  Same constructor, read(), read(vector,off,len), close(); the same std::runtime_error texts
// ("Invalid BZip2 header", "BZip2 stream format error", "BZip2 stream CRC error", "BZip2 block CRC error", "block Huffman
// tables invalid", "BZip2 block exceeds declared block size", "Error decoding  block", "BZip2 start pointer invalid",
// "Insufficient data").  An unmodified app.cpp decompress branch (app.cpp:126-160) compiles against it.
//
// What changed underneath: the reference decodes one block at a time on the host (BlockDecompressor.hpp:37-283).  Here the
// whole compressed stream is read at the first read() and handed to bz2b200_decompress_stream, which finds every block by
// its magic and decodes all of them on the GPU; read() then serves bytes from the decoded buffer.
#ifndef INPUT_STREAM_HPP
#define INPUT_STREAM_HPP

#include <istream>
#include <iterator>
#include <stdexcept>
#include <string>
#include <vector>
#include <cstdint>
#include <cstring>

#include "Config.hpp"
#include "bz2b200.h"

class InputStream
{
private:
    std::istream &inputStream;
    bool streamComplete = false;
    bool decoded = false;
    bz2b200_t *handle = nullptr;
    std::vector<uint8_t> data; // decoded bytes
    size_t position = 0;
    int parallelBlockCnt;
    bool allStreams; // decode every concatenated stream (bzip2 -d behaviour) instead of the reference's first stream only

    void decodeAll()
    {
        decoded = true;
        std::vector<uint8_t> in((std::istreambuf_iterator<char>(inputStream)), std::istreambuf_iterator<char>());
        if (in.size() < 4 || in[0] != 'B' || in[1] != 'Z' || in[2] != 'h' || in[3] < '1' || in[3] > '9')
            throw std::runtime_error("Invalid BZip2 header"); // InputStream.hpp:118-123
        // blocks decoded per device batch: the entropy decode is one warp per block, so throughput grows with the blocks in flight --
        // parallelBlockCnt is a minimum, the device's memory the limit (half as many again as the compressor keeps in flight)
        const int mult = allStreams ? 9 : in[3] - '0';
        int blocks = bz2b200_recommended_blocks(0, BLOCKSIZE_DEFAULT, mult) * 3 / 2;
        if (blocks < parallelBlockCnt)
            blocks = parallelBlockCnt;
        int rc = bz2b200_create(&handle, 0, BLOCKSIZE_DEFAULT, mult, blocks); // later streams may use larger blocks
        if (rc != BZ2B200_OK)
            throw std::runtime_error(std::string(bz2b200_strerror(rc)) + " (no CUDA device? there is no CPU path)");
        size_t cap = in.size() * 6 + (1u << 16), n = 0;
        for (;;)
        {
            data.resize(cap);
            rc = allStreams ? bz2b200_decompress_concat(handle, in.data(), in.size(), data.data(), cap, &n, nullptr)
                            : bz2b200_decompress_stream(handle, in.data(), in.size(), data.data(), cap, &n);
            if (rc != BZ2B200_E_OUTPUT)
                break;
            cap = n > cap ? n : cap * 2; // n = what the stream decodes to (bz2b200.h): one retry
        }
        if (rc == BZ2B200_E_FORMAT)
            throw std::runtime_error(bz2b200_last_error(handle)); // the reference's own exception text
        if (rc != BZ2B200_OK)
            throw std::runtime_error(std::string(bz2b200_strerror(rc)) + ": " + bz2b200_last_error(handle));
        data.resize(n);
    }

public:
    // parallelBlockCnt: blocks decoded per device batch; allStreams: go on behind the first stream while another "BZh" header follows
    // (SURVEY 8f N4).  Neither is in the reference's signature; both are defaulted to its behaviour.
    explicit InputStream(std::istream &in, int parallelBlockCnt = 256, bool allStreams = false)
        : inputStream(in), parallelBlockCnt(parallelBlockCnt), allStreams(allStreams) {}
    InputStream(const InputStream &) = delete;
    InputStream &operator=(const InputStream &) = delete;
    ~InputStream()
    {
        if (handle)
            bz2b200_destroy(handle);
    }

    int read() // InputStream.hpp:51-72
    {
        if (streamComplete)
            return -1;
        if (!decoded)
            decodeAll();
        if (position < data.size())
            return data[position++];
        return -1;
    }

    int read(std::vector<uint8_t> &buffer, int offset, int length) // :74-95
    {
        if (streamComplete)
            return -1;
        if (!decoded)
            decodeAll();
        if (position >= data.size())
            return -1;
        size_t n = data.size() - position;
        if (n > static_cast<size_t>(length))
            n = static_cast<size_t>(length);
        std::memcpy(buffer.data() + offset, data.data() + position, n);
        position += n;
        return static_cast<int>(n);
    }

    void close() // :97-101
    {
        streamComplete = true;
        data.clear();
        data.shrink_to_fit();
    }
};

#endif
