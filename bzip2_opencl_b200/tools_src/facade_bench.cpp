// facade_bench -- times file -> OutputStream -> .bz2 bytes through the reference-named facade (bench.py's "e2e_facade").
//   facade_bench <input> <output> <multiplier> <parallel> <mode>     mode: fd   readFrom(fd): the file is read straight into the pinned staging buffers
//                                                                           vec  write(vector, off, len) in 1 MiB pieces (reference API, OutputStream.hpp:144-161)
//                                                                           byte write(int) per byte, the loop of app.cpp:105-116
// Prints one JSON object: setup_ms (constructor: handle, pinned buffers), ms (first write .. close() returned), bytes in / out.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <fstream>
#include <iostream>
#include <unistd.h>
#include <vector>
#include "OutputStream.hpp"

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv)
{
    if (argc < 6) { std::fprintf(stderr, "usage: facade_bench <input> <output> <multiplier> <parallel> fd|vec|byte\n"); return 2; }
    const std::string mode = argv[5];
    try
    {
        std::ofstream out(argv[2], std::ios::binary);
        if (!out.is_open()) { std::fprintf(stderr, "cannot open output\n"); return 1; }
        const double t0 = now_ms();
        OutputStream bz2out(out, std::atoi(argv[3]), std::atoi(argv[4]));
        const double t1 = now_ms();
        size_t n_in = 0;
        if (mode == "fd")
        {
            const int fd = ::open(argv[1], O_RDONLY);
            if (fd < 0) { std::fprintf(stderr, "cannot open input\n"); return 1; }
            n_in = bz2out.readFrom(fd);
            ::close(fd);
        }
        else
        {
            std::ifstream in(argv[1], std::ios::binary);
            if (!in.is_open()) { std::fprintf(stderr, "cannot open input\n"); return 1; }
            std::vector<char> buffer(mode == "vec" ? (1u << 20) : 131072u);
            while (in)
            {
                in.read(buffer.data(), (std::streamsize)buffer.size());
                const std::streamsize n = in.gcount();
                if (n <= 0) break;
                n_in += (size_t)n;
                if (mode == "vec") bz2out.write(buffer, 0, (int)n);
                else for (std::streamsize i = 0; i < n; ++i) bz2out.write(buffer[(size_t)i]);
            }
        }
        bz2out.close();
        out.flush();
        const double t2 = now_ms();
        const long long n_out = (long long)out.tellp();
        std::printf("{\"mode\": \"%s\", \"setup_ms\": %.3f, \"ms\": %.3f, \"bytes_in\": %zu, \"bytes_out\": %lld, \"devices\": %d}\n", mode.c_str(), t1 - t0, t2 - t1,
                    n_in, n_out, bz2out.devices());
    }
    catch (const std::exception &e)
    {
        std::fprintf(stderr, "facade_bench: %s\n", e.what());
        return 3;
    }
    return 0;
}
