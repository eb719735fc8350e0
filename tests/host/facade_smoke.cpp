// Compiles the reference-named facade exactly the way app.cpp's compress branch uses it (app.cpp:93-125) and runs it.
// usage: facade_smoke errors | facade_smoke <input> <output.bz2> <multiplier> <parallel> [bulk]
#include <fstream>
#include <iostream>
#include <vector>
#include <cstring>
#include "OutputStream.hpp"

int main(int argc, char **argv)
{
    if (argc >= 2 && std::strcmp(argv[1], "errors") == 0)
    {
        // exception parity without needing a GPU (argument checks come first, OutputStream.hpp:73-81)
        std::ofstream sink("/dev/null", std::ios::binary);
        int ok = 0;
        try { OutputStream o(sink, 0, 4); } catch (const std::invalid_argument &e) { ok += std::strcmp(e.what(), "Invalid block size") == 0; }
        try { OutputStream o(sink, 10, 4); } catch (const std::invalid_argument &e) { ok += std::strcmp(e.what(), "Invalid block size") == 0; }
        try { OutputStream o(sink, 9, 0); } catch (const std::invalid_argument &e) { ok += std::strcmp(e.what(), "Invalid parallel block count") == 0; }
        // host-only pieces of the surface
        CRC32 c; c.updateCRC('a'); c.updateCRC('b', 3);
        unsigned char blk[64]; bool present[256] = {false};
        BlockCompressor bc(blk, present, 30);
        for (int i = 0; i < 10; ++i) bc.write('z');
        bc.write('q'); bc.finishRLE();
        ok += bc.getBlockLength() == 6 && blk[4] == 6 && present[6] && present['z'] && present['q'];
        bool bits[64]; size_t cnt = 0;
        writeBits(bits, &cnt, 5, 0x15); writeUnary(bits, &cnt, 2); padding(bits, &cnt);
        ok += cnt == 8 && bits[0] && !bits[1] && bits[2] && !bits[3] && bits[4] && bits[5] && bits[6] && !bits[7];
        std::cout << "errors-ok " << ok << " crc " << std::hex << (unsigned)c.getCRC() << std::endl;
        return ok == 5 ? 0 : 1;
    }
    if (argc < 5) return 2;
    std::ifstream in(argv[1], std::ios::binary);
    std::ofstream out(argv[2], std::ios::binary);
    OutputStream bz2out(out, std::atoi(argv[3]), std::atoi(argv[4]));
    const size_t bufferSize = 131072;
    std::vector<char> buffer(bufferSize);
    bool bulk = argc > 5;
    while (in)
    {
        in.read(buffer.data(), bufferSize);
        std::streamsize n = in.gcount();
        if (bulk) bz2out.write(buffer, 0, (int)n);
        else for (int i = 0; i < n; ++i) bz2out.write(buffer[i]);
    }
    bz2out.close();
    try { bz2out.write(1); return 3; } catch (const std::runtime_error &) {}
    return 0;
}
