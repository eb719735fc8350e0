// OutputStream.hpp -- facade of the reference's OutputStream (/root/reference/include/OutputStream.hpp:35-241) on top
// of the C ABI in include/bz2b200.h.  Same constructor, write(int), write(vector,off,len), close(); same exceptions
// (std::invalid_argument "Invalid block size" / "Invalid parallel block count", :73-81; std::runtime_error
// "Write beyond end of stream", :133-136).  The unmodified app.cpp compiles against it (tests/test_host_logic.py).
//
// What changed underneath: the reference pushes every byte through a host BlockCompressor, and once parallelBlockCnt
// blocks are full it ships 16 bool-bytes per input byte to and from ONE device and waits (:131-142,190-240, device picked
// in :61).  Here:
//   * write() only appends to one of TWO pinned staging buffers (bz2b200_host_alloc).  When a buffer holds a batch, a worker
//     thread hands it to bz2b200_compress_blocks -- which finds the block boundaries and does RLE1/CRC/BWT/MTF/Huffman on
//     the GPUs, copying the input in segments while earlier blocks are already being compressed -- and writes the packed
//     bits to the ostream, while the caller keeps filling the other buffer (file reading and compression overlap);
//   * every visible GPU is used: one multi-device handle (bz2b200_create_multi) instead of the reference's single Device;
//   * parallelBlockCnt is a MINIMUM: batches are as large as the devices can hold (bz2b200_recommended_blocks), because a
//     five-block device batch -- the reference CLI's default -p 10 at two lanes -- leaves a B200 idle.  The output never depends on it.
// The bytes of the last, unfinished block of a batch are carried into the next batch (each buffer keeps `head` bytes of room
// in front: no unfinished block is longer than that).  The < 8 bits between calls are the reference's leftBuffer (:225-239);
// header, end marker, folded stream CRC and padding stay here (:126-128,169-172,202).
// Extensions (not in the reference): write(const void*, size_t) and readFrom(fd) fill the staging buffers without the
// per-byte call of app.cpp:112-115; BZ2B200_DEVICES=0,1,.. / BZ2B200_FACADE_BLOCKS=n override device list and batch size.
#ifndef OUTPUT_STREAM_HPP
#define OUTPUT_STREAM_HPP

#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <mutex>
#include <ostream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include <sys/stat.h>
#include <unistd.h>

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
    int deviceCount = 1;

    // two pinned staging buffers: [head room for the carried-over bytes | new input]
    uint8_t *stage[2] = {nullptr, nullptr};
    size_t head = 0, cap = 0; // new input goes to [head, cap)
    size_t target = 0;        // bytes of new input that make a batch: small at first (the GPUs start early), then doubling up to cap - head
    size_t limit() const { return head + target; }
    size_t maxTarget() const { return cap - head; }
    int readers = 8;          // threads readFrom() copies a regular file with (BZ2B200_FACADE_READERS)
    size_t parallelReadMin = size_t(8) << 20; // shorter reads are not worth the threads (BZ2B200_FACADE_READ_MIN, bytes; tests lower it)
    size_t firstTarget = 0;   // readFrom() on a regular file: the first batch (short: the GPUs start early), the rest in equal batches
    bool trace = false;       // BZ2B200_FACADE_TRACE=1: a host-clock timeline of batches on stderr (development aid)
    double t0ms = 0;
    static double nowMs() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
    int cur = 0;              // buffer being filled
    size_t fill = 0;          // next free byte in stage[cur]
    uint8_t *outBuf[2] = {nullptr, nullptr}; // pinned: packed bits of a batch; one is written to the ostream while the next batch fills the other
    uint8_t *outBits = nullptr;               // the one the running batch fills
    int outCur = 0;
    size_t outCap = 0;
    // the writer: the ostream takes the bits of batch k while batch k+1 is being compressed
    std::thread writer;
    std::mutex wmu;
    std::condition_variable wcv;
    bool wPosted = false, wDone = true, wQuit = false;
    const uint8_t *wPtr = nullptr;
    size_t wLen = 0;
    std::exception_ptr wFailure;
    std::vector<uint32_t> blockCRCs;
    bool carryBits[8]; // < 8 bits waiting for the next byte (the reference's leftBuffer)
    size_t carryCount = 0;

    // the batch in flight
    std::thread worker;
    std::mutex mu;
    std::condition_variable cv;
    bool jobPosted = false, jobDone = true, quit = false;
    int jobBuf = 0;
    size_t jobBegin = 0, jobEnd = 0;
    bool jobFinal = false;
    size_t leftBegin = 0, leftEnd = 0; // bytes of the finished job that belong to the next batch
    int leftBuf = 0;
    std::exception_ptr failure;
    bool broken = false; // a batch failed: the exception has been thrown once, the stream is unusable from then on

    static std::vector<int> deviceList()
    {
        std::vector<int> ids;
        if (const char *e = std::getenv("BZ2B200_DEVICES"))
        {
            for (const char *p = e; *p;)
            {
                char *q = nullptr;
                long v = std::strtol(p, &q, 10);
                if (q == p)
                    break;
                ids.push_back(static_cast<int>(v));
                p = (*q == ',') ? q + 1 : q;
            }
        }
        if (ids.empty())
        {
            const int n = bz2b200_device_count();
            for (int i = 0; i < n; ++i)
                ids.push_back(i);
        }
        return ids;
    }

    // closeBlocks(), reference :190-240: every complete block of the batch -> bits -> outputStream.  Worker thread.
    void closeBlocks(uint8_t *buf, size_t begin, size_t end, bool final)
    {
        uint64_t nbits = 0, nblocks = 0;
        size_t consumed = 0;
        const double tc = trace ? nowMs() : 0;
        int rc = bz2b200_compress_blocks(handle, buf + begin, end - begin, final ? 1 : 0, static_cast<unsigned>(carryCount), outBits, outCap, &nbits,
                                         &consumed, blockCRCs.data(), nullptr, blockCRCs.size(), &nblocks);
        if (trace)
            std::fprintf(stderr, "[facade %8.2f] batch of %zu bytes compressed in %.2f ms (%llu blocks, %zu bytes carried over)\n", nowMs() - t0ms, end - begin,
                         nowMs() - tc, static_cast<unsigned long long>(nblocks), end - begin - consumed);
        if (rc != BZ2B200_OK)
            throw std::runtime_error(std::string(bz2b200_strerror(rc)) + ": " + bz2b200_last_error(handle));
        for (uint64_t i = 0; i < nblocks; ++i) // reference :201-202
            streamCRC = static_cast<int>(((static_cast<unsigned int>(streamCRC) << 1) | (static_cast<unsigned int>(streamCRC) >> 31)) ^ blockCRCs[i]);
        if (nbits)
        {
            // OR the carried bits into the top of the first byte, write whole bytes, keep the new carry
            uint8_t first = 0;
            for (size_t i = 0; i < carryCount; ++i)
                first = static_cast<uint8_t>(first | ((carryBits[i] ? 1 : 0) << (7 - i)));
            outBits[0] = static_cast<uint8_t>(outBits[0] | first);
            const uint64_t total = carryCount + nbits;
            const size_t whole = static_cast<size_t>(total / 8);
            carryCount = static_cast<size_t>(total % 8);
            for (size_t i = 0; i < carryCount; ++i)
                carryBits[i] = (outBits[whole] >> (7 - i)) & 1;
            postWrite(outBits, whole); // (waits for the previous batch's bytes to be out, then returns: the other buffer is free)
            outCur ^= 1;
            outBits = outBuf[outCur];
        }
        leftBegin = begin + consumed;
        leftEnd = end;
    }

    void writerMain()
    {
        for (;;)
        {
            std::unique_lock<std::mutex> lk(wmu);
            wcv.wait(lk, [&] { return wPosted || wQuit; });
            if (!wPosted)
                return;
            wPosted = false;
            const uint8_t *p = wPtr;
            const size_t n = wLen;
            lk.unlock();
            try
            {
                outputStream.write(reinterpret_cast<const char *>(p), static_cast<std::streamsize>(n));
            }
            catch (...)
            {
                wFailure = std::current_exception();
            }
            lk.lock();
            wDone = true;
            wcv.notify_all();
        }
    }

    void waitWritten()
    {
        std::unique_lock<std::mutex> lk(wmu);
        wcv.wait(lk, [&] { return wDone; });
        if (wFailure)
        {
            std::exception_ptr f = wFailure;
            wFailure = nullptr;
            std::rethrow_exception(f);
        }
    }

    void postWrite(const uint8_t *p, size_t n)
    {
        waitWritten();
        {
            std::lock_guard<std::mutex> lk(wmu);
            wPtr = p;
            wLen = n;
            wPosted = true;
            wDone = false;
        }
        wcv.notify_all();
    }

    void workerMain()
    {
        for (;;)
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return jobPosted || quit; });
            if (!jobPosted)
                return;
            jobPosted = false;
            const int b = jobBuf;
            const size_t jb = jobBegin, je = jobEnd;
            const bool fin = jobFinal;
            lk.unlock();
            try
            {
                closeBlocks(stage[b], jb, je, fin);
                leftBuf = b;
            }
            catch (...)
            {
                failure = std::current_exception();
                leftBegin = leftEnd = 0;
            }
            lk.lock();
            jobDone = true;
            cv.notify_all();
        }
    }

    void waitIdle()
    {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return jobDone; });
        if (failure)
        {
            std::exception_ptr f = failure;
            failure = nullptr;
            broken = true;
            std::rethrow_exception(f);
        }
        if (broken)
            throw std::runtime_error("the stream failed earlier");
    }

    // hands the filled buffer to the worker (after the previous batch has finished: its unfinished tail goes in front) and
    // goes on with the other buffer
    void submit(bool final)
    {
        const double tw = trace ? nowMs() : 0;
        waitIdle();
        if (trace)
            std::fprintf(stderr, "[facade %8.2f] submit: %zu new bytes, waited %.2f ms for the batch in front\n", nowMs() - t0ms, fill - head, nowMs() - tw);
        size_t begin = head;
        const size_t left = leftEnd - leftBegin;
        if (left)
        {
            if (left > head)
                throw std::runtime_error("internal: an unfinished block is longer than the staging head"); // cannot happen: bz2b200.h, slice primitives
            begin = head - left;
            std::memcpy(stage[cur] + begin, stage[leftBuf] + leftBegin, left);
            leftBegin = leftEnd = 0;
        }
        if (begin == fill && !final)
            return;
        {
            std::lock_guard<std::mutex> lk(mu);
            jobBuf = cur;
            jobBegin = begin;
            jobEnd = fill;
            jobFinal = final;
            jobPosted = true;
            jobDone = false;
        }
        cv.notify_all();
        cur ^= 1;
        fill = head;
        target = target * 2 < cap - head ? target * 2 : cap - head;
    }

    // threads joined (after their work is done), pinned buffers and the handle released; also the way out of a constructor that fails half way
    void release()
    {
        if (worker.joinable())
        {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return jobDone; });
                quit = true;
            }
            cv.notify_all();
            worker.join();
        }
        if (writer.joinable())
        {
            {
                std::unique_lock<std::mutex> lk(wmu);
                wcv.wait(lk, [&] { return wDone; });
                wQuit = true;
            }
            wcv.notify_all();
            writer.join();
        }
        for (int i = 0; i < 2; ++i)
        {
            bz2b200_host_free(stage[i]);
            bz2b200_host_free(outBuf[i]);
            stage[i] = outBuf[i] = nullptr;
        }
        bz2b200_destroy(handle);
        handle = nullptr;
    }


public:
    OutputStream(std::ostream &out, int blockSizeMultiplier, int parallelBlockCnt)
        : outputStream(out), streamBlockSize(BLOCKSIZE_DEFAULT * blockSizeMultiplier), parallelBlockCnt(parallelBlockCnt)
    {
        if (blockSizeMultiplier < 1 || blockSizeMultiplier > 9)
            throw std::invalid_argument("Invalid block size");
        if (parallelBlockCnt < 1)
            throw std::invalid_argument("Invalid parallel block count");
        try
        {
            construct(blockSizeMultiplier);
        }
        catch (...)
        {
            release(); // (a constructor that throws runs no destructor)
            throw;
        }
    }

private:
    void construct(int blockSizeMultiplier)
    {
        const std::vector<int> ids = deviceList();
        if (ids.empty())
            throw std::runtime_error("bz2b200_create: no CUDA device (a CUDA device is required; there is no CPU fallback)");
        // -p is a minimum: the device batch is what the devices can hold
        int blocks = parallelBlockCnt;
        if (const char *e = std::getenv("BZ2B200_FACADE_BLOCKS"))
            blocks = std::atoi(e) > 0 ? std::atoi(e) : blocks;
        else
        {
            int rec = bz2b200_recommended_blocks(ids[0], BLOCKSIZE_DEFAULT, blockSizeMultiplier);
            if (rec > blocks)
                blocks = rec;
        }
        int rc = ids.size() > 1 ? bz2b200_create_multi(&handle, ids.data(), static_cast<int>(ids.size()), BLOCKSIZE_DEFAULT, blockSizeMultiplier, blocks)
                                : bz2b200_create(&handle, ids[0], BLOCKSIZE_DEFAULT, blockSizeMultiplier, blocks);
        if (rc == BZ2B200_E_BLOCKSIZE)
            throw std::invalid_argument("Invalid block size");
        if (rc == BZ2B200_E_PARALLEL)
            throw std::invalid_argument("Invalid parallel block count");
        if (rc != BZ2B200_OK)
            throw std::runtime_error(std::string("bz2b200_create: ") + bz2b200_strerror(rc) + " (a CUDA device is required; there is no CPU fallback)");
        deviceCount = static_cast<int>(ids.size());
        // staging: no unfinished block is longer than `head` raw bytes (255-byte run chunks of 5 encoded bytes each)
        head = ((static_cast<size_t>(streamBlockSize) / 5 + 2) * 255 + 2 * static_cast<size_t>(streamBlockSize) + 4095) / 4096 * 4096;
        // a batch: what the devices hold in flight, but not more than 384 MiB per device -- beyond that a longer batch gains little, while
        // the pinned buffers cost set-up time and the first batch waits for its bytes to be read
        size_t batch = static_cast<size_t>(streamBlockSize) * static_cast<size_t>(blocks);
        if (batch > (static_cast<size_t>(384) << 20))
            batch = static_cast<size_t>(384) << 20;
        batch *= static_cast<size_t>(deviceCount);
        cap = head + batch;
        target = batch < (static_cast<size_t>(128) << 20) * static_cast<size_t>(deviceCount) ? batch : (static_cast<size_t>(128) << 20) * static_cast<size_t>(deviceCount);
        // readFrom(): a regular file is copied out of the page cache by several threads (one thread moves 3-4 GB/s, a GPU compresses 10);
        // its first batch is short, the rest of the file goes in equal batches
        {
            const unsigned hw = std::thread::hardware_concurrency();
            // (half the cores, eight at most: measured on a 16-core B200 host, 8 threads already read 2-4 times faster than the GPU compresses,
            // and 16 get in the way of the worker, lane and writer threads -- profiles/r02_facade_sweep.json)
            readers = hw == 0 ? 8 : static_cast<int>(hw / 2 < 2 ? 2 : hw / 2 > 8 ? 8 : hw / 2);
            if (const char *e = std::getenv("BZ2B200_FACADE_READERS"))
                readers = std::atoi(e) > 0 ? (std::atoi(e) > 64 ? 64 : std::atoi(e)) : readers;
            if (const char *e = std::getenv("BZ2B200_FACADE_READ_MIN"))
                parallelReadMin = std::atol(e) > 0 ? static_cast<size_t>(std::atol(e)) : parallelReadMin;
            size_t firstMb = 128;
            if (const char *e = std::getenv("BZ2B200_FACADE_FIRST_MB"))
                firstMb = std::atoi(e) > 0 ? static_cast<size_t>(std::atoi(e)) : firstMb;
            firstTarget = (firstMb << 20) * static_cast<size_t>(deviceCount);
            if (firstTarget > batch)
                firstTarget = batch;
            if (const char *e = std::getenv("BZ2B200_FACADE_TRACE"))
                trace = std::atoi(e) != 0;
            t0ms = nowMs();
        }
        for (int i = 0; i < 2; ++i)
        {
            stage[i] = static_cast<uint8_t *>(bz2b200_host_alloc(cap));
            if (!stage[i])
                throw std::runtime_error("bz2b200_host_alloc failed");
        }
        outCap = bz2b200_max_output_bytes(handle, cap) + 16;
        for (int i = 0; i < 2; ++i)
        {
            outBuf[i] = static_cast<uint8_t *>(bz2b200_host_alloc(outCap));
            if (!outBuf[i])
                throw std::runtime_error("bz2b200_host_alloc failed");
        }
        outBits = outBuf[0];
        blockCRCs.resize(cap / static_cast<size_t>((streamBlockSize - 5) * 4 / 5 > 0 ? (streamBlockSize - 5) * 4 / 5 : 1) + 4);
        fill = head;
        {
            // an empty call makes the handle allocate its per-batch device buffers now instead of inside the first batch
            uint64_t nbits = 0, nblocks = 0;
            size_t consumed = 0;
            rc = bz2b200_compress_blocks(handle, nullptr, 0, 0, 0, outBits, outCap, &nbits, &consumed, nullptr, nullptr, 0, &nblocks);
            if (rc != BZ2B200_OK)
                throw std::runtime_error(std::string(bz2b200_strerror(rc)) + ": " + bz2b200_last_error(handle));
        }
        // stream header, reference :126-128 (before the threads exist: a throwing ostream must not leave them behind)
        const char hdr[4] = {'B', 'Z', 'h', static_cast<char>('0' + blockSizeMultiplier)};
        outputStream.write(hdr, 4);
        worker = std::thread([this] { workerMain(); });
        writer = std::thread([this] { writerMain(); });
    }

public:
    OutputStream(const OutputStream &) = delete;
    OutputStream &operator=(const OutputStream &) = delete;

    ~OutputStream() { release(); }

    void write(int value) // reference :131-142
    {
        if (streamFinished)
            throw std::runtime_error("Write beyond end of stream");
        stage[cur][fill++] = static_cast<uint8_t>(value & 0xff);
        if (fill == limit())
            submit(false);
    }

    void write(const void *data, size_t length) // extension: any number of bytes at once
    {
        if (streamFinished)
            throw std::runtime_error("Write beyond end of stream");
        const uint8_t *p = static_cast<const uint8_t *>(data);
        while (length)
        {
            const size_t n = length < limit() - fill ? length : limit() - fill;
            std::memcpy(stage[cur] + fill, p, n);
            fill += n;
            p += n;
            length -= n;
            if (fill == limit())
                submit(false);
        }
    }

    void write(const std::vector<char> &data, int offset, int length) // reference :144-161
    {
        write(data.data() + offset, static_cast<size_t>(length));
    }

    // extension: everything the file descriptor still holds, read straight into the staging buffers (no intermediate copy).
    // A regular file is read by `readers` threads at once (pread on disjoint ranges: one thread copies out of the page cache at a
    // few GB/s, less than a GPU compresses).
    size_t readFrom(int fd)
    {
        if (streamFinished)
            throw std::runtime_error("Write beyond end of stream");
        size_t total = 0;
        struct stat sb;
        const off_t at = ::lseek(fd, 0, SEEK_CUR);
        if (::fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode) && at >= 0)
        {
            size_t pos = static_cast<size_t>(at);
            const size_t end = static_cast<size_t>(sb.st_size);
            // the size is known: a short first batch, then what is left in equal batches (no small batch at the end, none larger than the buffers)
            auto plan = [&](bool first) {
                const size_t left = end > pos ? end - pos : 0;
                if (fill != head || left == 0)
                    return; // a batch is being filled already: it keeps its size
                size_t t = first && firstTarget ? firstTarget : maxTarget();
                if (!first || left <= t)
                {
                    const size_t k = (left + maxTarget() - 1) / maxTarget();
                    t = ((left + k - 1) / k + 4095) / 4096 * 4096;
                }
                target = t < maxTarget() ? t : maxTarget();
            };
            plan(true);
            while (pos < end)
            {
                const size_t n = end - pos < limit() - fill ? end - pos : limit() - fill;
                const int T = n >= parallelReadMin ? readers : 1;
                std::vector<std::thread> th;
                bool ok = true;
                uint8_t *dst = stage[cur] + fill;
                for (int t = 0; t < T; ++t)
                    th.emplace_back([=, &ok] {
                        size_t a = n * static_cast<size_t>(t) / T;
                        const size_t e = n * static_cast<size_t>(t + 1) / T;
                        while (a < e)
                        {
                            const ssize_t r = ::pread(fd, dst + a, e - a, static_cast<off_t>(pos + a));
                            if (r <= 0) { ok = false; return; }
                            a += static_cast<size_t>(r);
                        }
                    });
                for (std::thread &x : th)
                    x.join();
                if (!ok)
                    throw std::runtime_error("read failed");
                fill += n;
                pos += n;
                total += n;
                if (fill == limit())
                {
                    submit(false);
                    plan(false);
                }
            }
            ::lseek(fd, static_cast<off_t>(end), SEEK_SET);
            return total;
        }
        for (;;)
        {
            const ssize_t n = ::read(fd, stage[cur] + fill, limit() - fill);
            if (n < 0)
                throw std::runtime_error("read failed");
            if (n == 0)
                return total;
            fill += static_cast<size_t>(n);
            total += static_cast<size_t>(n);
            if (fill == limit())
                submit(false);
        }
    }

    int devices() const { return deviceCount; } // extension

    void close() // reference :163-176
    {
        if (streamFinished)
            return;
        streamFinished = true;
        submit(true);
        waitIdle();
        waitWritten();
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
