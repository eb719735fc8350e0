// The OutputStream facade over tests/host/stub_abi.c (the oracle behind the C ABI's names): its host pipeline -- batches, the bytes
// and bits carried between them, worker and writer threads -- on the CPU, built with -fsanitize=thread / address by
// tests/test_host_logic.py.  TEST INFRASTRUCTURE; the product facade links libbz2b200.so.
// usage: facade_stub_driver <mode> <input> <output.bz2> <multiplier> <parallel>
//   mode: byte (app.cpp:112-115: write(int) per byte) | vec (write(vector, off, len) in ragged pieces) | ptr (write(ptr, n)) |
//         fd (readFrom) | mixed (all four in turn) | fail (expects a std::runtime_error from the pipeline; exit 0 if it came)
#include <fcntl.h>
#include <unistd.h>
#include <cstring>
#include <fstream>
#include <iostream>
#include <vector>
#include "OutputStream.hpp"

static std::vector<char> slurp(const char *path)
{
    std::ifstream in(path, std::ios::binary);
    return std::vector<char>((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
}

int main(int argc, char **argv)
{
    if (argc < 6)
        return 2;
    const std::string mode = argv[1];
    std::ofstream out(argv[3], std::ios::binary);
    const int mult = std::atoi(argv[4]), par = std::atoi(argv[5]);
    if (mode == "fail")
    {
        const std::vector<char> data = slurp(argv[2]);
        try
        {
            OutputStream o(out, mult, par);
            o.write(data.data(), data.size());
            o.close();
        }
        catch (const std::runtime_error &e)
        {
            std::cout << "caught: " << e.what() << std::endl;
            return 0; // (and the destructor has joined both threads, or we would not be here)
        }
        return 1;
    }
    OutputStream o(out, mult, par);
    if (mode == "fd")
    {
        const int fd = ::open(argv[2], O_RDONLY);
        if (fd < 0)
            return 2;
        o.readFrom(fd);
        ::close(fd);
    }
    else
    {
        const std::vector<char> data = slurp(argv[2]);
        size_t pos = 0;
        unsigned piece = 1;
        int turn = 0;
        while (pos < data.size())
        {
            // ragged pieces: 1, 3, 7, 15, ... up to 64 KiB - 1, then again
            piece = piece * 2 + 1 > 65535 ? 1 : piece * 2 + 1;
            const size_t n = std::min<size_t>(piece, data.size() - pos);
            const int how = mode == "mixed" ? turn++ % 3 : mode == "byte" ? 0 : mode == "vec" ? 1 : 2;
            if (how == 0)
                for (size_t i = 0; i < n; ++i)
                    o.write(static_cast<unsigned char>(data[pos + i]));
            else if (how == 1)
                o.write(data, static_cast<int>(pos), static_cast<int>(n));
            else
                o.write(data.data() + pos, n);
            pos += n;
        }
    }
    o.close();
    o.close(); // idempotent (reference :163-165)
    try
    {
        o.write(1);
        return 3;
    }
    catch (const std::runtime_error &)
    {
    }
    return 0;
}
