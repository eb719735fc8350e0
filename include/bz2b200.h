/*
 * bz2b200.h -- C ABI of the B200-native bzip2 block compressor (libbz2b200.so).
 *
 * This is the drop-in boundary for the compress path of Stan1slav337/Bzip2-OpenCL.  It replaces everything
 * that reference keeps behind include/opencl.hpp -- Device / Memory<T> / Kernel (opencl.hpp:145-539), the
 * 14 mirrored buffers and the blocking upload / kernel_close->run() / download sequence of
 * OutputStream::closeBlocks() (OutputStream.hpp:83-116,215-223) -- AND the per-byte host RLE1 loop that feeds
 * it (BlockCompressor.hpp:69-154 driven by app.cpp:108-116): raw input bytes go in, the concatenated
 * bit-exact block bit streams come out.  The reference-named C++ facade (Config.hpp, CRC32.hpp,
 * BlockCompressor.hpp, BitOutputStream.hpp, OutputStream.hpp in bzip2_opencl_b200/include/) sits on top of
 * these entry points; INTEGRATION.md shows the binding.
 *
 * Plain pointers and sizes only.  Every `input` / `out` pointer may be a host pointer (pageable or pinned)
 * or a device pointer on the handle's device; the library detects which.  All work is CUDA (sm_100a); there
 * is no CPU fallback: without a usable device bz2b200_create fails with BZ2B200_E_CUDA.
 *
 * Threading: one handle = one host thread at a time.  bz2b200_create makes a handle for one GPU; bz2b200_create_multi makes one
 * handle that fans a stream out over several GPUs of the box inside the call (blocks share no state -- SURVEY 8(e)) and is
 * accepted by every function below.  Jobs with one PROCESS per GPU build the same stream from the slice primitives
 * (bz2b200_scan_prepare / bz2b200_scan_blocks_ex / bz2b200_compress_scanned / bz2b200_shift_bits).
 */
#ifndef BZ2B200_H
#define BZ2B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bz2b200 bz2b200_t;

enum {
    BZ2B200_OK = 0,
    BZ2B200_E_BLOCKSIZE = -1, /* multiplier outside 1..9 / bad base -> facade throws std::invalid_argument("Invalid block size") (OutputStream.hpp:73-76) */
    BZ2B200_E_PARALLEL = -2,  /* max_blocks_in_flight < 1 -> std::invalid_argument("Invalid parallel block count") (OutputStream.hpp:78-81) */
    BZ2B200_E_ARG = -3,       /* null pointer, bad start_bit, bad stage id ... */
    BZ2B200_E_CUDA = -4,      /* no device / CUDA runtime error (text in bz2b200_last_error) */
    BZ2B200_E_NOMEM = -5,     /* device or pinned allocation failed */
    BZ2B200_E_OUTPUT = -6,    /* out_capacity / block_capacity too small; nothing useful was written */
    BZ2B200_E_STATE = -7,     /* stage dump asked for a block that is not in the last batch */
    BZ2B200_E_FORMAT = -8     /* decompress: the input is not a valid stream; bz2b200_last_error holds the reference's exception text */
};

/* Number of CUDA devices visible (0 when there is none or the driver is unusable). */
int bz2b200_device_count(void);

/*
 * Create a compressor on CUDA device `device_id`.
 *   base_block_size       BLOCKSIZE_DEFAULT of the reference (Config.hpp:30): 10000 for the stock reference,
 *                         100000 for real bzip2 block sizes (BASELINE configs).  Any value in [1000, 100000].
 *   multiplier            1..9 (OutputStream.hpp:65-76); block size = base_block_size * multiplier.
 *   max_blocks_in_flight  the reference's parallelBlockCnt (-p, OutputStream.hpp:67): how many blocks one
 *                         device batch holds.  Output never depends on it.
 * Replaces: Device{select_device_with_most_flops()} + the Memory<> / Kernel constructors (OutputStream.hpp:61,83-116).
 */
int bz2b200_create(bz2b200_t **handle, int device_id, int base_block_size, int multiplier, int max_blocks_in_flight);

/*
 * One handle over several GPUs: device_ids[0..n_devices) (a device may be named more than once), max_blocks_in_flight per device.
 * Replaces the reference's single `Device device{select_device_with_most_flops()}` (OutputStream.hpp:61): where the reference
 * batches parallelBlockCnt blocks on one device (OutputStream.hpp:190-240), a multi handle cuts the input of a call into one
 * contiguous byte slice per device, finds the block boundaries with a chain of walks (device d starts where device d-1's last
 * complete block ended), compresses every slice on its own device and places each device's bits at their bit offset in the
 * caller's buffer.  The bytes produced are identical to those of a one-GPU handle.  Inputs of less than 16 blocks per extra
 * device use fewer devices.  BZ2B200_E_ARG for a bad list; otherwise as bz2b200_create.
 */
int bz2b200_create_multi(bz2b200_t **handle, const int *device_ids, int n_devices, int base_block_size, int multiplier, int max_blocks_in_flight);

/* Devices behind a handle (1 for bz2b200_create). */
int bz2b200_device_count_of(const bz2b200_t *handle);

/* Blocks per device batch that keep `device_id` busy without taking more than a third of its free memory (0: no such device).
 * The facade uses it so that the reference's small default of -p (app.cpp:41) does not turn into five-block device batches. */
int bz2b200_recommended_blocks(int device_id, int base_block_size, int multiplier);

/* Pinned (page-locked, portable) host memory for staging buffers handed to the calls below: copies from and to it are
 * asynchronous and overlap with the kernels.  Replaces the host arrays of Memory<T> (opencl.hpp:250-262). NULL on failure. */
void *bz2b200_host_alloc(size_t bytes);
void bz2b200_host_free(void *p);

/* Releases every device and pinned buffer (Memory<T>::~Memory, opencl.hpp:276-278). NULL is allowed. */
void bz2b200_destroy(bz2b200_t *handle);

/* Upper bound for the bits bz2b200_compress_blocks can produce for n_bytes of input, in BYTES. */
size_t bz2b200_max_output_bytes(const bz2b200_t *handle, size_t n_bytes);

/*
 * Find the block boundaries the reference's sequential RLE1 state machine would produce
 * (BlockCompressor::write refusal rule, BlockCompressor.hpp:71-74; OutputStream::getNextCompressor,
 * OutputStream.hpp:179-188) for `n_bytes` of input, without compressing.
 *   block_starts[i] = input offset of the first byte of block i; *n_blocks = number of blocks (0 for empty input).
 * If more than `capacity` blocks exist only the first `capacity` starts are stored and BZ2B200_E_OUTPUT is returned
 * (*n_blocks is still the true count).  Used by multi-GPU callers to shard block ranges.
 */
int bz2b200_scan_blocks(bz2b200_t *handle, const uint8_t *input, size_t n_bytes, uint64_t *block_starts, size_t capacity,
                        uint64_t *n_blocks);

/*
 * The slice primitives: one stream over several handles (one per GPU, in one or in several processes).
 *
 * bz2b200_scan_prepare     queue the staging of input[0..n_bytes) (host input: segmented copies) and the summary tree over it;
 *                          returns without waiting, so every device of a job prepares its slice at the same time.
 * bz2b200_scan_blocks_ex   the walk of bz2b200_scan_blocks over a slice: starts with a fresh machine at input[first_start]
 *                          (the position where the previous slice's last complete block ended; the bytes in front of it are only
 *                          context) and, with is_final == 0, stops in front of the first block whose refusing byte is not in the
 *                          slice; *consumed = offset of that block's first byte = the next slice's first_start, relative to
 *                          `input`.  Uses the preparation of the same (input, n_bytes) if there is one.  block_starts may be NULL.
 *                          A slice must begin at most one unfinished block (255 * (block size / 5 + 2) raw bytes at the very most)
 *                          in front of first_start, but may begin anywhere before that.
 * bz2b200_compress_scanned compress exactly the blocks the last bz2b200_scan_blocks_ex on this handle found (the input stays
 *                          staged from the scan); arguments as for bz2b200_compress_blocks.
 * bz2b200_shift_bits       move a part to its bit offset in the stream: dst = src (nbits bits, device memory) shifted right by
 *                          shift = 0..7 bits, the top `shift` bits of dst[0] taken from `carry` (the part in front's trailing
 *                          bits, already in the top bits of the byte).  This is the reference's leftBuffer hand-over
 *                          (BitOutputStream.hpp:30-45, OutputStream.hpp:225-239) between devices.
 *
 * bz2b200_compress_blocks_from   scan + compress of a slice in one call, the form the chain of a multi-GPU stream uses: the first block
 *                          starts at input[first_start]; on_walked(next_start, user) (may be NULL) is called from inside the call as
 *                          soon as the walk has been through the slice -- next_start is what bz2b200_scan_blocks_ex would return in
 *                          *consumed -- so that the next device's walk starts while this one still compresses (the walk never waits
 *                          for the kernels; batches start as soon as blocks are known).  Uses the preparation of the same
 *                          (input, n_bytes) if there is one; otherwise host input is copied in segments under the kernels.
 */
typedef void (*bz2b200_walked_fn)(uint64_t next_start, void *user);
int bz2b200_compress_blocks_from(bz2b200_t *handle, const uint8_t *input, size_t n_bytes, int is_final, uint64_t first_start, unsigned start_bit,
                                 uint8_t *out, size_t out_capacity, uint64_t *out_nbits, size_t *consumed, uint32_t *block_crcs, uint64_t *block_nbits,
                                 size_t block_capacity, uint64_t *n_blocks, bz2b200_walked_fn on_walked, void *user);
int bz2b200_scan_prepare(bz2b200_t *handle, const uint8_t *input, size_t n_bytes);
int bz2b200_scan_blocks_ex(bz2b200_t *handle, const uint8_t *input, size_t n_bytes, int is_final, uint64_t first_start, uint64_t *block_starts,
                           size_t capacity, uint64_t *n_blocks, size_t *consumed);
int bz2b200_compress_scanned(bz2b200_t *handle, unsigned start_bit, uint8_t *out, size_t out_capacity, uint64_t *out_nbits, uint32_t *block_crcs,
                             uint64_t *block_nbits, size_t block_capacity, uint64_t *n_blocks);
int bz2b200_shift_bits(bz2b200_t *handle, const uint8_t *src, uint64_t nbits, unsigned shift, unsigned carry, uint8_t *dst, size_t dst_capacity);

/*
 * The hot path.  Compress input[0..n_bytes) into bzip2 blocks: RLE1 + CRC32 + symbol map, BWT, MTF/RLE2,
 * 6-table Huffman, bit packing -- everything closeBlocks() + kernel_close produce for every block
 * (OutputStream.hpp:190-240, kernel.cpp:3099-3159), i.e. per block
 *   48b magic | 32b block CRC | 1b 0 | 24b origPtr | symbol map | tables | selectors | codes     
 * concatenated at BIT granularity with no padding between blocks.  The stream header ("BZh9") and trailer
 * (end magic, stream CRC, padding) are NOT written -- OutputStream does that (OutputStream.hpp:126-128,169-172),
 * or use bz2b200_compress_stream.
 *
 *   is_final     non-zero: the input ends here; the last partial block is flushed (finishRLE, BlockCompressor.hpp:112-118).
 *                zero: only blocks that are complete (their refusing byte was seen) are produced; *consumed tells the
 *                caller how many input bytes they covered -- the rest must be presented again at the start of the next call.
 *   start_bit    0..7: bit offset inside out[0] at which the first block starts (the carry of the previous call, the
 *                reference's leftBuffer, BitOutputStream.hpp:30-45).  The top start_bit bits of out[0] are written as 0
 *                for the caller to OR its carry bits in.
 *   out          receives ceil((start_bit + *out_nbits)/8) bytes; unused low bits of the last byte are 0.
 *   out_nbits    number of bits produced (excluding start_bit).
 *   block_crcs / block_nbits   per-block CRC (as written in the header) and bit length; may be NULL.  block_capacity
 *                entries each; BZ2B200_E_OUTPUT if there are more blocks.
 */
int bz2b200_compress_blocks(bz2b200_t *handle, const uint8_t *input, size_t n_bytes, int is_final, unsigned start_bit,
                            uint8_t *out, size_t out_capacity, uint64_t *out_nbits, size_t *consumed, uint32_t *block_crcs,
                            uint64_t *block_nbits, size_t block_capacity, uint64_t *n_blocks);

/*
 * Whole .bz2 stream in one call: "BZh" + ('0'+multiplier), all blocks, end-of-stream magic, combined CRC
 * (crc = rotl(crc,1) ^ blockCRC, OutputStream.hpp:202), zero padding.  == OutputStream ctor + write* + close().
 */
int bz2b200_compress_stream(bz2b200_t *handle, const uint8_t *input, size_t n_bytes, uint8_t *out, size_t out_capacity,
                            size_t *out_bytes);

/*
 * The decompress side: one whole .bz2 stream in, the original bytes out.  == InputStream ctor + read* until -1
 * (InputStream.hpp:36-159) with everything it drives: BlockDecompressor (BlockDecompressor.hpp:37-283), HuffmanStageDecoder
 * (HuffmanStageDecoder.hpp:33-137), MoveToFront::indexToFront (MoveToFront.hpp:62-72), BitInputStream (BitInputStream.hpp:30-85).
 * Blocks are found by scanning for the 48-bit block magic at every bit position and decoded independently on the device; block CRCs
 * and the stream CRC are verified.  The stream's block size (header digit x base_block_size of the handle) must not exceed the
 * handle's.  Errors: BZ2B200_E_FORMAT with the reference's exception text in bz2b200_last_error ("Invalid BZip2 header",
 * "BZip2 stream format error", "BZip2 stream CRC error", "BZip2 block CRC error", "block Huffman tables invalid",
 * "BZip2 block exceeds declared block size", "Error decoding  block", "BZip2 start pointer invalid", "Insufficient data");
 * BZ2B200_E_OUTPUT when out_capacity is too small: *out_bytes = the number of bytes the input decodes to (the decoder goes on without
 * writing, so one retry with a buffer of that size succeeds).
 */
int bz2b200_decompress_stream(bz2b200_t *handle, const uint8_t *input, size_t n_bytes, uint8_t *out, size_t out_capacity, size_t *out_bytes);

/*
 * Concatenated streams (interoperability beyond the reference's decoder): like bz2b200_decompress_stream, but
 * after a complete stream decoding goes on while another "BZh1".."BZh9" header starts at the next byte -- what `bzip2 -d` / `cat a.bz2
 * b.bz2 | bzip2 -d` accept.  The reference reads exactly one stream and never looks behind its CRC (InputStream.hpp:146-153);
 * bytes behind the last complete stream that do not start a stream are ignored here in the same way.  Every stream has its own
 * block-size digit and combined CRC.  *n_streams (may be NULL) = streams decoded.  Errors as for bz2b200_decompress_stream; an
 * error in a later stream fails the call.
 */
int bz2b200_decompress_concat(bz2b200_t *handle, const uint8_t *input, size_t n_bytes, uint8_t *out, size_t out_capacity, size_t *out_bytes,
                              uint64_t *n_streams);

/* Per-stage device time (CUDA events on the handle's stream) and counters of the LAST compress call. */
typedef struct {
    float ms_total;       /* first kernel to last kernel, device side, excludes host<->device copies */
    float ms_h2d, ms_d2h; /* copies done by the call (0 for device-resident buffers) */
    float ms_rle, ms_bwt_sort0, ms_bwt_rounds, ms_bwt_finish, ms_mtf, ms_huffman, ms_emit;
    float ms_walk;        /* block-boundary walk (side stream, overlapped with the previous batch) */
    uint64_t n_input, n_rle, n_mtf, n_out_bits, n_blocks;
    uint64_t bwt_rounds;        /* prefix-doubling rounds executed (max over batches) */
    uint64_t bwt_active_volume; /* sum over rounds of active (unresolved) rotations: rho * n_rle of SURVEY 8(d) */
    uint64_t kernel_launches;   /* kernels launched by the call */
} bz2b200_stats_t;
int bz2b200_get_stats(const bz2b200_t *handle, bz2b200_stats_t *stats);

/*
 * Optional per-kernel timing: with profiling on, every kernel launch is bracketed by CUDA events on the launching stream;
 * bz2b200_kernel_times(index) returns the name, summed milliseconds, launch count and (for the sort / MTF kernels) summed algorithmic bytes of kernel `index` since profiling was
 * switched on (BZ2B200_E_STATE when index is past the last kernel).  Used by bench.py for the roofline line.
 */
int bz2b200_set_profiling(bz2b200_t *handle, int on);
int bz2b200_kernel_times(const bz2b200_t *handle, int index, const char **name, double *ms, uint64_t *launches, uint64_t *algorithmic_bytes);

/*
 * Test hook: copy one intermediate of block `block_index` (index within the last
 * compress call) to host memory.  Only blocks of the LAST device batch are available.
 */
enum {
    BZ2B200_STAGE_RLE1 = 1,      /* uint8  [n]      RLE1'd block bytes          (BlockCompressor.hpp:134-154) */
    BZ2B200_STAGE_PRESENT = 2,   /* uint8  [256]    blockValuesPresent          (BlockCompressor.hpp:136,148) */
    BZ2B200_STAGE_BWT = 3,       /* uint8  [n]      BWT last column L           (kernel.cpp:2429) */
    BZ2B200_STAGE_ORIGPTR = 4,   /* uint32 [1]      origPtr                     (kernel.cpp:3114-3116) */
    BZ2B200_STAGE_MTF = 5,       /* uint16 [m]      MTF/RLE2 symbols incl. EOB  (kernel.cpp:2561-2649) */
    BZ2B200_STAGE_FREQ = 6,      /* int32  [258]    mtfSymbolFrequencies        (kernel.cpp:2613,2641-2643) */
    BZ2B200_STAGE_SELECTORS = 7, /* uint8  [nsel]   final selectors             (kernel.cpp:2940-2943) */
    BZ2B200_STAGE_LENGTHS = 8,   /* uint8  [T*258]  final code lengths          (kernel.cpp:2947-2950) */
    BZ2B200_STAGE_RANK = 9       /* uint32 [n]      rank[i] = first SA position of rotation i's tie group (debug) */
};
int bz2b200_stage_dump(bz2b200_t *handle, uint64_t block_index, int stage, void *dst, size_t dst_capacity, size_t *len_bytes);

const char *bz2b200_strerror(int code);
/* Detail text of the last failure on this handle (CUDA error string ...); "" if none. */
const char *bz2b200_last_error(const bz2b200_t *handle);

#ifdef __cplusplus
}
#endif
#endif /* BZ2B200_H */
