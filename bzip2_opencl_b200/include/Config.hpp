// Config.hpp -- facade copy of the reference's format constants (/root/reference/include/Config.hpp:27-47), same names
// and values so code written against the reference compiles unchanged.  BLOCKSIZE_DEFAULT is the one knob: the stock
// reference uses 10000 ("-9" = 90 kB blocks, Config.hpp:30); define BZ2B200_BLOCKSIZE_DEFAULT=100000 for real bzip2 sizes.
#ifndef CONFIG_HPP
#define CONFIG_HPP

#ifndef BZ2B200_BLOCKSIZE_DEFAULT
#define BZ2B200_BLOCKSIZE_DEFAULT 10000
#endif

static constexpr int BLOCK_HEADER_MARKER_1 = 0x314159;
static constexpr int BLOCK_HEADER_MARKER_2 = 0x265359;
static constexpr int ALPHABET_SIZE = 256;
static constexpr int BLOCKSIZE_DEFAULT = BZ2B200_BLOCKSIZE_DEFAULT;
static constexpr int MAX_BLOCK_SIZE = BLOCKSIZE_DEFAULT * 9;
static constexpr int BWT_BUCKET_A_SIZE = 256;
static constexpr int BWT_BUCKET_B_SIZE = 65536;
static constexpr int HUFFMAN_GROUP_RUN_LENGTH = 50;
static constexpr int HUFFMAN_MAXIMUM_ALPHABET_SIZE = 258;
static constexpr int HUFFMAN_HIGH_SYMBOL_COST = 15;
static constexpr int HUFFMAN_ENCODE_MAXIMUM_CODE_LENGTH = 20;
static constexpr int HUFFMAN_DECODE_MAXIMUM_CODE_LENGTH = 23;
static constexpr int HUFFMAN_MINIMUM_TABLES = 2;
static constexpr int HUFFMAN_MAXIMUM_TABLES = 6;
static constexpr int HUFFMAN_MAXIMUM_SELECTORS = (MAX_BLOCK_SIZE / HUFFMAN_GROUP_RUN_LENGTH) + 1;
static constexpr int HUFFMAN_SYMBOL_RUNA = 0;
static constexpr int HUFFMAN_SYMBOL_RUNB = 1;
static constexpr int STREAM_END_MARKER_1 = 0x177245;
static constexpr int STREAM_END_MARKER_2 = 0x385090;
static constexpr int STREAM_START_MARKER_1 = 0x425a;
static constexpr int STREAM_START_MARKER_2 = 0x68;

#endif
