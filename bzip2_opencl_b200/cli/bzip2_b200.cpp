// bzip2_b200 -- command line front end with the reference's flags and file rules (/root/reference/app.cpp:31-179), on top of
// the reference-named facades (OutputStream.hpp / InputStream.hpp -> libbz2b200.so).  SURVEY 8(f) row N3.
//
//   bzip2_b200 <file> [--dec|-d] [--keep|-k] [--check|-c] [--size|-s 1..9] [--parallel|-p N] [--help|-h]
//
// Same behaviour: compress <file> -> <file>.bz2 and remove <file> unless -k (app.cpp:93-125); -d needs a ".bz2" name, writes the
// name without it, or "<name>.out" with a warning when that exists (:128-141); -c decodes and prints "Integrity check passed!"
// (:161-168); exit status 1 with the reference's messages for a missing file / unknown flag.  Deliberate differences:
// the file name may stand anywhere on the line (the reference always takes argv[1], app.cpp:82), the input is removed once, not
// twice (:119-122 and :173-176), input is fed through write(vector, off, len) instead of byte by byte (:112-115), and a failed
// (de)compression reports the exception text and keeps the input.
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <unistd.h>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "OutputStream.hpp"
#include "InputStream.hpp"

int main(int argc, char *argv[])
{
    const char *flags = "\n\n  [--help|-h]              print help\n  [--dec|-d]               decompress file\n  [--keep|-k]              keep original (de)compressed file\n  [--check|-c]             check compressed file integrity\n  [--size|-s <1-9>]        set block size (x BLOCKSIZE_DEFAULT)\n  [--parallel|-p <1+>]     blocks per device batch\n  [--all-streams|-a]       with -d / -c: decode every concatenated stream (not in the reference)\n";
    if (argc < 2)
    {
        std::cerr << "\n  Usage: bzip2_b200 [file_path] [flags]" << flags << std::endl;
        return 1;
    }
    std::string filename;
    bool decompress = false, keepFile = false, checkCRC = false, allStreams = false;
    int blockSize = 9;     // app.cpp:46
    int parallelCnt = 10;  // app.cpp:47
    for (int i = 1; i < argc; ++i)
    {
        const char *a = argv[i];
        if (!std::strcmp(a, "--dec") || !std::strcmp(a, "-d")) decompress = true;
        else if (!std::strcmp(a, "--keep") || !std::strcmp(a, "-k")) keepFile = true;
        else if ((!std::strcmp(a, "--size") || !std::strcmp(a, "-s")) && i + 1 < argc) blockSize = std::atoi(argv[++i]);
        else if ((!std::strcmp(a, "--parallel") || !std::strcmp(a, "-p")) && i + 1 < argc) parallelCnt = std::atoi(argv[++i]);
        else if (!std::strcmp(a, "--check") || !std::strcmp(a, "-c")) checkCRC = true;
        else if (!std::strcmp(a, "--all-streams") || !std::strcmp(a, "-a")) allStreams = true;
        else if (!std::strcmp(a, "--help") || !std::strcmp(a, "-h"))
        {
            std::cout << "\n  Usage: bzip2_b200 [file_path] [flags]" << flags << std::endl;
            return 0;
        }
        else if (a[0] == '-')
        {
            std::cerr << "  Unknown flag!\n\n  Usage: bzip2_b200 [file_path] [flags]" << flags << std::endl;
            return 1;
        }
        else filename = a;
    }
    std::ifstream inputFile(filename, std::ios::binary);
    if (!inputFile.is_open())
    {
        std::cerr << "Failed to open input file." << std::endl;
        return 1;
    }
    try
    {
        if (!decompress && !checkCRC)
        {
            const std::string outputFilename = filename + ".bz2";
            std::ofstream outputFile(outputFilename, std::ios::binary);
            if (!outputFile.is_open())
            {
                std::cerr << "Failed to open output file.";
                return 1;
            }
            OutputStream bz2out(outputFile, blockSize, parallelCnt);
            // the file goes straight into the facade's pinned staging buffers (no per-byte write(int) as in app.cpp:112-115)
            const int fd = ::open(filename.c_str(), O_RDONLY);
            if (fd < 0)
            {
                std::cerr << "Failed to open input file." << std::endl;
                return 1;
            }
            bz2out.readFrom(fd);
            ::close(fd);
            bz2out.close();
            outputFile.close();
        }
        else if (!checkCRC)
        {
            if (filename.size() < 4 || filename.substr(filename.size() - 4) != ".bz2")
            {
                std::cerr << "Input file doesn't have right .bz2 extenstion for decompression!" << std::endl;
                return 1;
            }
            std::string outputFilename = filename.substr(0, filename.size() - 4);
            if (std::ifstream(outputFilename))
            {
                outputFilename += ".out";
                std::cerr << "Warning: Output file exists. Writing to " << outputFilename << std::endl;
            }
            std::ofstream outputFile(outputFilename, std::ios::binary);
            if (!outputFile.is_open())
            {
                std::cerr << "Failed to open output file." << std::endl;
                return 1;
            }
            InputStream bz2in(inputFile, parallelCnt < 64 ? 64 : parallelCnt, allStreams);
            std::vector<uint8_t> buffer(1 << 20);
            int n;
            while ((n = bz2in.read(buffer, 0, (int)buffer.size())) != -1) outputFile.write((const char *)buffer.data(), n);
            bz2in.close();
            outputFile.close();
        }
        else
        {
            InputStream bz2in(inputFile, parallelCnt < 64 ? 64 : parallelCnt, allStreams);
            std::vector<uint8_t> buffer(1 << 20);
            while (bz2in.read(buffer, 0, (int)buffer.size()) != -1)
                ;
            bz2in.close();
            std::cout << "  Integrity check passed!" << std::endl;
        }
    }
    catch (const std::exception &e)
    {
        std::cerr << "bzip2_b200: " << e.what() << std::endl;
        return 2;
    }
    inputFile.close();
    if (!keepFile && !checkCRC) std::remove(filename.c_str());
    return 0;
}
