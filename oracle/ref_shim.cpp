// TEST / BASELINE INFRASTRUCTURE ONLY -- never linked into or loaded by the product (bzip2_opencl_b200/).
//
// CPU build of the REAL reference: the reference's host headers (BlockCompressor.hpp, BitOutputStream.hpp,
// CRC32.hpp, InputStream.hpp ...) are #included from /root/reference/include, and the OpenCL C text of
// kernel.cpp (generated + patched P1/P3/P4 by build_ref.py into _ref/cfg_<base>/kernel.cl) is compiled as
// C++ behind the macro shim below.  This file only re-creates the driver that
// include/OutputStream.hpp:126-128,163-176,190-240 implements around the OpenCL wrapper:
//   - bytes go one at a time through BlockCompressor::write        (app.cpp:112-115, OutputStream.hpp:131-142)
//   - per block: finishRLE, CRC fold, 48+32+1 header bits          (OutputStream.hpp:192-213)
//   - close_block() per block (kernel.cpp:3099-3122) instead of kernel_close->run(); blocks are independent
//     work-items in the reference, here they run on `threads` host threads
//   - bit streams concatenated in block order                       (OutputStream.hpp:225-239)
// Patch P2: every block gets its own zeroed freq[258] and a bs+1 byte slot.
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <vector>
#include <string>
#include <sstream>
#include <thread>
#include <atomic>
#include <memory>
#include <stdexcept>

#include "Config.hpp" // patched copy in _ref/cfg_<base>/ (include guard keeps the reference's own copy out)
#include "BlockCompressor.hpp"
#include "BitOutputStream.hpp"
#include "InputStream.hpp"

namespace dev
{
    typedef unsigned int uint;
#define constant static const
#define global
#define kernel
#define private
#define get_global_id(x) 0
#include "kernel.cl"
#undef constant
#undef global
#undef kernel
#undef private
#undef get_global_id
}

namespace
{
    thread_local std::string g_err;

    struct RleBlock
    {
        std::vector<unsigned char> data; // bs + 1
        bool present[256];
        int length = 0;
        int crc = 0;
        size_t inputStart = 0;
    };

    // app.cpp:112-115 + OutputStream::write(int) + getNextCompressor, with an unbounded list of compressors
    std::vector<std::unique_ptr<RleBlock>> rle1_split(const uint8_t *in, size_t n, int bs)
    {
        std::vector<std::unique_ptr<RleBlock>> blocks;
        auto fresh = [&](size_t start) {
            auto b = std::make_unique<RleBlock>();
            b->data.assign((size_t)bs + 1, 0);
            std::memset(b->present, 0, sizeof(b->present));
            b->inputStart = start;
            return b;
        };
        std::unique_ptr<RleBlock> cur = fresh(0);
        auto bc = std::make_unique<BlockCompressor>(cur->data.data(), cur->present, bs);
        auto finish = [&]() {
            if (!bc->isEmpty())
            {
                bc->finishRLE();
                cur->length = bc->getBlockLength();
                cur->crc = bc->getCRC();
                blocks.push_back(std::move(cur));
            }
        };
        for (size_t i = 0; i < n; ++i)
        {
            int v = in[i] & 0xff;
            if (!bc->write(v))
            {
                finish();
                cur = fresh(i);
                bc = std::make_unique<BlockCompressor>(cur->data.data(), cur->present, bs);
                bc->write(v);
            }
        }
        finish();
        return blocks;
    }

    struct Scratch
    {
        std::vector<int> SA, bucketA, bucketB, tmp, freq, symMap, mtf, sel;
        std::unique_ptr<bool[]> bits;
        explicit Scratch(int bs) : SA((size_t)bs + 1), bucketA(BWT_BUCKET_A_SIZE), bucketB(BWT_BUCKET_B_SIZE), tmp(256),
                                   freq(HUFFMAN_MAXIMUM_ALPHABET_SIZE), symMap(256), mtf(256),
                                   sel((bs + HUFFMAN_GROUP_RUN_LENGTH - 1) / HUFFMAN_GROUP_RUN_LENGTH + 1),
                                   bits(new bool[(size_t)16 * bs + 128]) {}
    };

    // One block through the reference device code; returns the packed bits (MSB first) and the bit count.
    void close_one(RleBlock &b, Scratch &s, std::vector<uint8_t> &packed, size_t &nbits)
    {
        size_t cnt = 0;
        bool *bits = s.bits.get();
        writeBits(bits, &cnt, 24, BLOCK_HEADER_MARKER_1); // OutputStream.hpp:208-211
        writeBits(bits, &cnt, 24, BLOCK_HEADER_MARKER_2);
        writeInteger(bits, &cnt, b.crc);
        writeBoolean(bits, &cnt, false);
        std::fill(s.freq.begin(), s.freq.end(), 0); // P2
        dev::close_block(b.data.data(), s.SA.data(), b.length, s.bucketA.data(), s.bucketB.data(), s.tmp.data(),
                         bits, &cnt, b.present, s.freq.data(), s.symMap.data(), s.mtf.data(), s.sel.data());
        nbits = cnt;
        packed.assign((cnt + 7) / 8, 0);
        for (size_t i = 0; i < cnt; ++i)
            if (bits[i])
                packed[i >> 3] |= (uint8_t)(0x80u >> (i & 7));
    }

    struct BitSink
    {
        uint8_t *out;
        size_t cap;
        size_t nbits = 0;
        bool overflow = false;
        void put(const uint8_t *src, size_t bits)
        {
            if ((nbits + bits + 7) / 8 > cap)
            {
                overflow = true;
                return;
            }
            unsigned sh = nbits & 7;
            size_t o = nbits >> 3;
            size_t nb = (bits + 7) / 8;
            if (sh == 0)
            {
                std::memcpy(out + o, src, nb);
            }
            else
            {
                for (size_t i = 0; i < nb; ++i)
                {
                    out[o + i] |= (uint8_t)(src[i] >> sh);
                    if (o + i + 1 < cap)
                        out[o + i + 1] = (uint8_t)(src[i] << (8 - sh));
                }
            }
            nbits += bits;
            // clear anything past the end
            size_t endb = nbits >> 3;
            unsigned r = nbits & 7;
            if (r && endb < cap)
                out[endb] &= (uint8_t)(0xff00u >> r);
            else if (!r && endb < cap)
                out[endb] = 0;
        }
        void bits(int count, uint32_t value)
        {
            uint8_t tmp[4];
            uint32_t v = value << (32 - count);
            tmp[0] = v >> 24;
            tmp[1] = v >> 16;
            tmp[2] = v >> 8;
            tmp[3] = v;
            put(tmp, count);
        }
    };
}

extern "C"
{
    int ref_base_block_size() { return BLOCKSIZE_DEFAULT; }
    const char *ref_last_error() { return g_err.c_str(); }

    // Whole-stream compress. Returns bytes written, -1 on error.
    long ref_compress(const uint8_t *in, size_t n, int mult, int threads, uint8_t *out, size_t cap)
    {
        try
        {
            if (mult < 1 || mult > 9)
                throw std::invalid_argument("Invalid block size");
            if (threads < 1)
                throw std::invalid_argument("Invalid parallel block count");
            const int bs = BLOCKSIZE_DEFAULT * mult;
            auto blocks = rle1_split(in, n, bs);
            const size_t B = blocks.size();
            std::vector<std::vector<uint8_t>> packed(B);
            std::vector<size_t> nbits(B);
            std::atomic<size_t> next{0};
            auto worker = [&]() {
                Scratch s(bs);
                for (;;)
                {
                    size_t i = next.fetch_add(1);
                    if (i >= B)
                        break;
                    close_one(*blocks[i], s, packed[i], nbits[i]);
                    blocks[i]->data = std::vector<unsigned char>(); // free
                }
            };
            int T = (int)std::min<size_t>((size_t)threads, B ? B : 1);
            std::vector<std::thread> th;
            for (int t = 1; t < T; ++t)
                th.emplace_back(worker);
            worker();
            for (auto &t : th)
                t.join();

            if (cap)
                out[0] = 0;
            BitSink sink{out, cap};
            sink.bits(16, STREAM_START_MARKER_1); // OutputStream.hpp:126-128
            sink.bits(8, STREAM_START_MARKER_2);
            sink.bits(8, '0' + mult);
            int streamCRC = 0;
            for (size_t i = 0; i < B; ++i)
            {
                int blockCRC = blocks[i]->crc; // OutputStream.hpp:201-202
                streamCRC = ((streamCRC << 1) | (static_cast<unsigned int>(streamCRC) >> 31)) ^ blockCRC;
                sink.put(packed[i].data(), nbits[i]);
            }
            sink.bits(24, STREAM_END_MARKER_1); // OutputStream.hpp:169-172
            sink.bits(24, STREAM_END_MARKER_2);
            sink.bits(16, ((uint32_t)streamCRC >> 16) & 0xffff);
            sink.bits(16, (uint32_t)streamCRC & 0xffff);
            if (sink.overflow)
                throw std::runtime_error("output buffer too small");
            return (long)((sink.nbits + 7) / 8);
        }
        catch (const std::exception &e)
        {
            g_err = e.what();
            return -1;
        }
    }

    // RLE1 split only: per block input start, RLE1 length, CRC. Returns number of blocks (<= max_blocks filled).
    long ref_rle1_blocks(const uint8_t *in, size_t n, int mult, uint64_t *starts, int32_t *lengths, uint32_t *crcs,
                         long max_blocks)
    {
        const int bs = BLOCKSIZE_DEFAULT * mult;
        auto blocks = rle1_split(in, n, bs);
        for (size_t i = 0; i < blocks.size() && (long)i < max_blocks; ++i)
        {
            starts[i] = blocks[i]->inputStart;
            lengths[i] = blocks[i]->length;
            crcs[i] = (uint32_t)blocks[i]->crc;
        }
        return (long)blocks.size();
    }

    // Stage dump of block `index` of the stream: RLE1 bytes, presence, BWT last column + origPtr, MTF/RLE2 symbols,
    // frequencies, selectors (from the final refinement pass), total block bit count.
    // Buffers: rle[bs+1], present[256], L[bs], mtf[bs+1] (uint16), freq[258] (int32), sel[bs/50+2] (uint8).
    // info[0]=length info[1]=origPtr info[2]=mtfLength info[3]=alphabetSize info[4]=nSelectors info[5]=crc info[6]=nbits(lo32)
    int ref_block_stages(const uint8_t *in, size_t n, int mult, long index, uint8_t *rle, uint8_t *present, uint8_t *L,
                         uint16_t *mtf, int32_t *freq, uint8_t *sel, int64_t *info)
    {
        try
        {
            const int bs = BLOCKSIZE_DEFAULT * mult;
            auto blocks = rle1_split(in, n, bs);
            if (index < 0 || (size_t)index >= blocks.size())
                throw std::out_of_range("block index");
            RleBlock &b = *blocks[index];
            Scratch s(bs);
            std::memcpy(rle, b.data.data(), b.length);
            for (int i = 0; i < 256; ++i)
                present[i] = b.present[i];
            b.data[b.length] = b.data[0];
            int orig = dev::DivSufSortBWT(b.data.data(), s.SA.data(), s.bucketA.data(), s.bucketB.data(), s.tmp.data(), b.length);
            for (int i = 0; i < b.length; ++i)
                L[i] = (uint8_t)(s.SA[i] & 0xff);
            std::fill(s.freq.begin(), s.freq.end(), 0);
            dev::MTFResult r = dev::MTFAndRLE2StageEncoder(s.SA.data(), b.length, b.present, s.freq.data(), s.symMap.data(), s.mtf.data());
            for (int i = 0; i < r.mtfLength; ++i)
                mtf[i] = (uint16_t)s.SA[i];
            for (int i = 0; i < HUFFMAN_MAXIMUM_ALPHABET_SIZE; ++i)
                freq[i] = s.freq[i];
            size_t cnt = 0;
            dev::HuffmanStageEncoder(s.bits.get(), &cnt, s.SA.data(), r.mtfLength, r.alphabetSize, s.freq.data(), s.sel.data());
            int nsel = (r.mtfLength + HUFFMAN_GROUP_RUN_LENGTH - 1) / HUFFMAN_GROUP_RUN_LENGTH;
            for (int i = 0; i < nsel; ++i)
                sel[i] = (uint8_t)s.sel[i];
            info[0] = b.length;
            info[1] = orig;
            info[2] = r.mtfLength;
            info[3] = r.alphabetSize;
            info[4] = nsel;
            info[5] = (uint32_t)b.crc;
            info[6] = (int64_t)cnt; // Huffman stage bits only (tables+selectors+data)
            return 0;
        }
        catch (const std::exception &e)
        {
            g_err = e.what();
            return -1;
        }
    }

    // Reference decoder (InputStream.hpp:51-72 loop as in app.cpp:148-153). Returns bytes, -1 on error, -2 on overflow.
    long ref_decompress(const uint8_t *in, size_t n, uint8_t *out, size_t cap)
    {
        try
        {
            std::string s(reinterpret_cast<const char *>(in), n);
            std::istringstream is(s, std::ios::binary);
            InputStream bz2in(is);
            size_t o = 0;
            int ch;
            while ((ch = bz2in.read()) != -1)
            {
                if (o >= cap)
                    return -2;
                out[o++] = (uint8_t)ch;
            }
            bz2in.close();
            return (long)o;
        }
        catch (const std::exception &e)
        {
            g_err = e.what();
            return -1;
        }
    }
}
