// Drives the reference-named InputStream facade the way app.cpp's decompress branch does (app.cpp:126-160).
// usage: facade_decompress <input.bz2> <output> [bulk|byte] [multi]     exit 0 = ok, 4 = std::runtime_error (text on stderr)
#include <fstream>
#include <iostream>
#include <vector>
#include <cstring>
#include "InputStream.hpp"

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    std::ifstream in(argv[1], std::ios::binary);
    std::ofstream out(argv[2], std::ios::binary);
    try
    {
        InputStream bz2in(in, 256, argc > 4 && !std::strcmp(argv[4], "multi"));
        if (argc > 3 && !std::strcmp(argv[3], "bulk"))
        {
            std::vector<uint8_t> buffer(131072);
            int n;
            while ((n = bz2in.read(buffer, 0, (int)buffer.size())) != -1) out.write((const char *)buffer.data(), n);
        }
        else
        {
            int c;
            while ((c = bz2in.read()) != -1) out.put((char)c);
        }
        bz2in.close();
        if (bz2in.read() != -1) return 3;
    }
    catch (const std::runtime_error &e)
    {
        std::cerr << e.what();
        return 4;
    }
    return 0;
}
