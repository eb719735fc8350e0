// bz2b200.cu -- host orchestration + C ABI (include/bz2b200.h) of the B200-native bzip2 block compressor.
// Replaces OutputStream::closeBlocks()'s device boundary (OutputStream.hpp:190-240) and include/opencl.hpp.
// No CPU fallback: every stage is a CUDA kernel from rle.cuh / bwt.cuh / mtf.cuh / huff.cuh / emit.cuh.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <chrono>

#include "../../include/bz2b200.h"
#include "rle.cuh"
#include "bwt.cuh"
#include "mtf.cuh"
#include "huff.cuh"
#include "emit.cuh"
#include "dec.cuh"

using namespace bz;

enum KernelId { KID_k_radix_hist, KID_k_radix_scatter, KID_k_radix_scan, KID_k_radix_count, KID_k_radix_bases, KID_k_radix_onesweep, KID_k_bwt_bytehist, KID_k_dec_scan, KID_k_dec_block, KID_k_imtf_tiles, KID_k_imtf_compose, KID_k_dec_bases, KID_k_dec_double, KID_k_dec_chase, KID_k_dec_expand, KID_k_rle_tile_seg, KID_k_rle_reduce, KID_k_rle_walk, KID_k_rle_btile_seg, KID_k_rle_btile_scan, KID_k_rle_emit, KID_k_crc_chunks, KID_k_crc_finish, KID_k_bwt_keys0, KID_k_bwt_regroup, KID_k_bwt_roundstats, KID_k_bwt_gather, KID_k_bwt_gather_local, KID_k_bwt_local, KID_k_bwt_finish, KID_k_bwt_period, KID_k_bwt_expand, KID_k_mtf_recency, KID_k_mtf_ranks, KID_k_mtf_compose, KID_k_rle2_scan, KID_k_rle2_emit, KID_k_huff_seeds, KID_k_huff_pass, KID_k_huff_lengths, KID_k_huff_codes, KID_k_huff_selmtf, KID_k_block_bits, KID_k_bit_offsets, KID_k_emit_header, KID_k_data_tilebits, KID_k_data_scan, KID_k_emit_data, KID_COUNT };
static const char *const kKernelNames[KID_COUNT] = { "k_radix_hist", "k_radix_scatter", "k_radix_scan", "k_radix_count", "k_radix_bases", "k_radix_onesweep", "k_bwt_bytehist", "k_dec_scan", "k_dec_block", "k_imtf_tiles", "k_imtf_compose", "k_dec_counts", "k_dec_double", "k_dec_chase", "k_dec_expand", "k_rle_tile_seg", "k_rle_reduce", "k_rle_walk", "k_rle_btile_seg", "k_rle_btile_scan", "k_rle_emit", "k_crc_chunks", "k_crc_finish", "k_bwt_keys0", "k_bwt_regroup", "k_bwt_roundstats", "k_bwt_gather", "k_bwt_gather_local", "k_bwt_local", "k_bwt_finish", "k_bwt_period", "k_bwt_expand", "k_mtf_recency", "k_mtf_ranks", "k_mtf_compose", "k_rle2_scan", "k_rle2_emit", "k_huff_seeds", "k_huff_pass", "k_huff_lengths", "k_huff_codes", "k_huff_selmtf", "k_block_bits", "k_bit_offsets", "k_emit_header", "k_data_tilebits", "k_data_scan", "k_emit_data" };

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct CudaErr {
    cudaError_t e;
    const char *what;
};

} // namespace

// Lanes: a handle compresses its batches on `n_lanes` lanes.  A lane is a handle of its own (per-batch buffers for G / n_lanes
// blocks, its own stream) driven by its own host thread, so the kernels of one batch that keep few SMs busy (one warp per block:
// list composition, code-length allocation, selector coding, headers) and the host round trips of the prefix-doubling rounds
// overlap with the wide kernels of the other batch.  What the batches share lives in the parent: input, block table, output
// stream and its bit position (handed from batch to batch in order, LaneCtl).
struct LaneCtl {
    std::mutex mu;
    std::condition_variable cv;
    uint64_t bits_issued = 0; // batches whose bit offsets have been queued (events ev_bits recorded)
    uint64_t done = 0;        // batches completed and drained, in order
    int rc = 0;               // first failure of any lane
    std::string err;
    cudaEvent_t ev_bits[4] = {};
    cudaEvent_t ev_setup = nullptr;
};

static double now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct bz2b200 {
    int trace = 0;        // BZ2B200_TRACE=1: host-clock timeline of walk chunks and batches on stderr (development aid)
    double t_call = 0;
    bz2b200 *parent = nullptr;      // lane: the handle it works for
    std::vector<bz2b200 *> children; // multi-device handle (bz2b200_create_multi): one single-device handle per entry; the parent owns no device state
    int usable_children = 0;        // BZ2B200_MULTI_MAX: upper bound for the children a call may use (0 = all)
    std::vector<bz2b200 *> lanes;   // parent: its lanes (empty: batches run on the handle itself)
    bool lanes_wanted = false;      // lanes are created at the first compress call (a handle used for decompression only never needs them)
    std::vector<std::thread> lane_thr;
    std::vector<char> lane_busy;
    std::vector<uint32_t> lane_nb;  // blocks of the batch a busy lane is working on
    LaneCtl *ctl = nullptr;
    int n_lanes = 2, last_lane = 0;
    uint64_t call_batches = 0;      // batches handed to lanes by the current call
    bool batch_ready = false;       // per-batch buffers allocated (a parent with lanes needs them for the decompress side only)
    uint64_t call_id = 0, lane_call = 0; // lane_call != parent's call_id: first batch of this call on the lane
    int device = 0;
    int base = 0, mult = 0, bs = 0;
    uint32_t G = 0;       // blocks per batch
    uint64_t S = 0;       // element stride per block
    int nbits = 0;        // bits needed for a position inside a block
    int key0_bytes = 5;   // bytes of the round-0 key (BZ2B200_KEY0_BYTES=1..5): 5 bytes = 40 key bits + 20 index bits in one element; one more
                          // radix pass than 4 bytes, but the doubling rounds then see 5, 10, 20 ... sorted bytes and ~25 % fewer rotations
    int walk_growth = 6;  // each walk chunk may be this many times the batch that hides it (the walk finds ~7 blocks while one is compressed)
    int walk_first = 48;  // blocks of the first chunk: its latency is exposed
    int prefix_walk = 1;  // BZ2B200_WALK=0: the tree-descent walk (k_rle_walk) instead of the prefix-state walk (k_rle_walk_prefix)
    uint64_t seg_first = 16ull << 20, seg_rest = 64ull << 20; // host input: bytes per H2D segment (BZ2B200_SEG_MB=first,rest)
    int pack_keys = 1;    // round-0 keys over the block's own alphabet (k_bwt_keys_packed); BZ2B200_PACK=0: the first key0_bytes bytes
    int onesweep = 1;     // BZ2B200_ONESWEEP=0: histogram + scan + scatter kernels per radix pass instead of the single-sweep pass
    int r0_bits = 8;      // BZ2B200_R0_BITS=10: four round-0 sweeps over 10-bit digits instead of five over 8-bit ones (measured SLOWER: 33.6 vs 30.6 ms of
                          // sweeps per GiB -- 1024 bins cost shared memory, a four times longer scan and look-back, and more distinct digits per MATCH)
    int r0_group = 0;     // BZ2B200_R0_GROUP=g: the five round-0 sweeps run group by group of g blocks so that a group's elements stay in L2 (measured: the
                          // small launches cost more than the HBM traffic they save, profiles/r02_*); 0 = every sweep over the whole batch
    int chain_overlap = 1; // BZ2B200_CHAIN_OVERLAP=0: a slice's batches start only when its walk is through (the walk alone on the device is ~30 % faster,
                           // but the slice's own compression then starts a walk later; measured equal at 8 GPUs, better overlapped at 2)
    int walk_smem = 190;   // BZ2B200_WALK_SMEM=KiB: dynamic shared memory of a chain's walk kernel (large: its warp has an SM to itself; 6.2 -> 5.6 ms/GiB)
    bool chain_walk = false; // the current call hands a chain on (bz2b200_compress_blocks_from with a callback)
    int bulk_copy = 1;    // BZ2B200_BULK=0: block-relative tiles are staged with byte loads instead of the TMA bulk copy (cp.async.bulk + mbarrier)
    int root_sort = 1;    // BZ2B200_ROOT=0: non-primitive blocks go through the full prefix doubling instead of their root (k_bwt_period)
    int local_sort = 1;   // sort tile-local groups in shared memory (BZ2B200_LOCALSORT=0 sends everything through the global radix sort)
    cudaStream_t stream = 0;
    std::string err;
    std::vector<DevBuf *> all;

    // per-call (grow only)
    DevBuf in, seg[RLE_MAX_LEVELS], segP[RLE_MAX_LEVELS], pstate, ltab, starts, rle_len, info, out, all_crc, all_nbits, segB, tstate;
    DevBuf dec_cand, dec_info, dec_list0, dec_len, dec_off, dec_starts, dec_crc; // decompress side (dec.cuh), allocated on first use
    // per-batch, fixed size
    DevBuf rle, L, ranks, kA, kB, loc, loc_cnt, rank, mtf, hist, tickets, digit_base, rc_partial, cnts, scalars,
        h0, root, work, lists, dcount, r2tiles, zin, obase, present, crc, orig, freq, m_arr, alpha_arr, hb, lens, tf, codes, sel, selpos, selbits, tabbits,
        databits, nbits_b, bitoff, datastart, tilebits, selbitoff, pertab;
    int ctasC = 0;
    int tilesR = 0, tilesG = 0, tilesM = 0, tilesR2 = 0, tilesD = 0;
    uint64_t sel_stride = 0;
    uint64_t *h_pinned = nullptr; // small pinned scratch (64 u64)

    // boundary walk, pipelined one batch ahead on a side stream
    cudaStream_t side = nullptr, copy = nullptr; // walk stream; output-drain stream (finished batches leave while the next one computes)
    uint64_t out_cap_bits = 0;      // bits the call's device output buffer holds (k_bit_offsets never places a block behind it)
    uint8_t *drain_dst = nullptr;   // caller's output buffer (nullptr: no overlapped drain)
    size_t drain_cap = 0, drained = 0;
    bool drain_dev = false;
    cudaEvent_t ev_tree = nullptr, ev_walk = nullptr, ev_w0 = nullptr, ev_w1 = nullptr;
    uint64_t *pin_starts = nullptr; // G + 2
    uint32_t *pin_len = nullptr;    // G + 2
    uint64_t *pin_info = nullptr;   // 4
    WalkLevels wl{};
    // input arrives in segments (host input: chunked H2D on its own stream, overlapped with the kernels); the tree of
    // summaries and the walk are extended segment by segment on the side stream
    cudaStream_t h2d = nullptr;
    std::vector<cudaEvent_t> ev_seg;
    std::vector<uint64_t> seg_end;
    size_t seg_avail = 0, seg_copied = 0;
    const uint8_t *host_src = nullptr; // non-null: segments still have to be copied from here
    bool host_pinned = false;
    uint64_t tiles_done = 0;
    bool tree_pending = false; // ev_t0/ev_t1 hold a tree update that walk_collect has not accounted yet
    int caller_final = 0;
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
    const uint8_t *w_in = nullptr;
    uint64_t w_n = 0, w_cap = 0, w_req = 0, w_found = 0;
    int w_final = 0;
    bool w_finished = true;

    // bz2b200_scan_prepare / bz2b200_scan_blocks_ex: the staged input, its summary tree and the block table stay valid for bz2b200_compress_scanned
    bool prepared = false, scan_valid = false;
    const uint8_t *prep_input = nullptr;
    size_t prep_n = 0;
    uint64_t scan_first = 0;
    DevBuf shifted; // multi-device streams: this device's part moved to its bit offset in the stream (k_shift_bits)

    // last call
    bz2b200_stats_t stats{};
    uint64_t last_b0 = 0, last_nb = 0, last_total_blocks = 0;
    std::vector<uint64_t> h_starts;
    std::vector<uint32_t> h_rle_len;
    cudaEvent_t ev[16] = {};

    // optional per-kernel timing (bz2b200_set_profiling): CUDA events around every launch on the launching stream
    bool profiling = false;
    std::vector<cudaEvent_t> prof_pool;
    std::vector<int> prof_ids;
    size_t prof_used = 0;
    double k_ms[KID_COUNT] = {};
    uint64_t k_launches[KID_COUNT] = {};
    uint64_t k_bytes[KID_COUNT] = {};   // algorithmic bytes (one read or write of each operand) of the profiled launches
};

static inline void prof_mark(bz2b200 *h, int id)
{
    if (h->prof_used == h->prof_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        h->prof_pool.push_back(e);
    }
    cudaEventRecord(h->prof_pool[h->prof_used++], h->stream);
    if (id >= 0) h->prof_ids.push_back(id);
}
// all recorded launches have completed (caller synchronised): fold the event pairs into the per-kernel totals
static void prof_collect(bz2b200 *h)
{
    for (size_t i = 0; i < h->prof_ids.size(); ++i) {
        float ms = 0;
        cudaEventElapsedTime(&ms, h->prof_pool[2 * i], h->prof_pool[2 * i + 1]);
        h->k_ms[h->prof_ids[i]] += ms;
        h->k_launches[h->prof_ids[i]] += 1;
    }
    h->prof_ids.clear();
    h->prof_used = 0;
}
#define KBYTES(id, bytes)                            \
    do {                                             \
        if (h->profiling) h->k_bytes[id] += (uint64_t)(bytes); \
    } while (0)
#define KL(id, ...)                                  \
    do {                                             \
        if (h->profiling) prof_mark(h, id);          \
        __VA_ARGS__;                                 \
        if (h->profiling) prof_mark(h, -1);          \
    } while (0)

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                          \
            return (e_ == cudaErrorMemoryAllocation) ? BZ2B200_E_NOMEM : BZ2B200_E_CUDA;           \
        }                                                                                          \
    } while (0)

static int ensure(bz2b200 *h, DevBuf &b, size_t bytes)
{
    if (b.cap >= bytes && b.p) return BZ2B200_OK;
    if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        e = cudaMalloc(&b.p, bytes + 256);
        want = bytes + 256;
    }
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        h->err = std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e);
        b.p = nullptr;
        return BZ2B200_E_NOMEM;
    }
    b.cap = want;
    if (std::find(h->all.begin(), h->all.end(), &b) == h->all.end()) h->all.push_back(&b);
    return BZ2B200_OK;
}
#define ENSURE(buf, bytes)                                   \
    do {                                                     \
        int rc_ = ensure(h, (buf), (size_t)(bytes));         \
        if (rc_ != BZ2B200_OK) return rc_;                   \
    } while (0)

template <typename T>
static T *ptr(DevBuf &b) { return reinterpret_cast<T *>(b.p); }

static int ilog2_ceil(uint64_t v)
{
    int b = 0;
    while ((1ull << b) < v) ++b;
    return b;
}

static bool g_const_ready[64] = {};

static int upload_constants(bz2b200 *h)
{
    if (h->device < 64 && g_const_ready[h->device]) return BZ2B200_OK;
    uint32_t table[256], pow8[48], lane_shift[256];
    crc_build_table(table);
    crc_build_pow8(pow8, 48);
    for (int t = 0; t < 256; ++t) lane_shift[t] = crc_xpow8_host(64ull * (255 - t), pow8);
    CK(cudaMemcpyToSymbol(c_crc_table, table, sizeof(table)));
    CK(cudaMemcpyToSymbol(c_crc_pow8, pow8, sizeof(pow8)));
    CK(cudaMemcpyToSymbol(c_crc_lane_shift, lane_shift, sizeof(lane_shift)));
    CK(cudaFuncSetAttribute(k_radix_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScatterSmem)));
    CK(cudaFuncSetAttribute(k_radix_onesweep<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScatterSmem)));
    CK(cudaFuncSetAttribute(k_radix_onesweep<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(cudaFuncSetAttribute(k_radix_onesweep<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScatterSmem)));
    CK(cudaFuncSetAttribute(k_radix_onesweep<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(cudaFuncSetAttribute(k_radix_onesweep<false, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScatterSmemT<10>)));
    CK(cudaFuncSetAttribute(k_radix_onesweep<false, 10>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(cudaFuncSetAttribute(k_radix_onesweep<true, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScatterSmemT<10>)));
    CK(cudaFuncSetAttribute(k_radix_onesweep<true, 10>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(cudaFuncSetAttribute(k_bwt_local, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(cudaFuncSetAttribute(k_rle_walk_prefix, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 << 10));
    CK(cudaFuncSetAttribute(k_dec_block, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(DB_WARPS * sizeof(DecTables))));
    // shared-memory heavy kernels: ask for the largest shared-memory carveout so several CTAs stay resident per SM
    CK(cudaFuncSetAttribute(k_mtf_ranks, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(cudaFuncSetAttribute(k_radix_scatter, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(cudaFuncSetAttribute(k_emit_data, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(cudaFuncSetAttribute(k_huff_pass, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    if (h->device < 64) g_const_ready[h->device] = true;
    return BZ2B200_OK;
}

extern "C" int bz2b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return n;
}

extern "C" const char *bz2b200_strerror(int code)
{
    switch (code) {
    case BZ2B200_OK: return "ok";
    case BZ2B200_E_BLOCKSIZE: return "Invalid block size";
    case BZ2B200_E_PARALLEL: return "Invalid parallel block count";
    case BZ2B200_E_ARG: return "invalid argument";
    case BZ2B200_E_CUDA: return "CUDA error";
    case BZ2B200_E_NOMEM: return "out of device memory";
    case BZ2B200_E_OUTPUT: return "output buffer too small";
    case BZ2B200_E_STATE: return "requested block is not resident";
    case BZ2B200_E_FORMAT: return "not a valid bzip2 stream";
    default: return "unknown error";
    }
}

extern "C" const char *bz2b200_last_error(const bz2b200_t *h) { return h ? h->err.c_str() : ""; }

extern "C" void bz2b200_destroy(bz2b200_t *h)
{
    if (!h) return;
    if (!h->children.empty()) {
        for (bz2b200 *c : h->children) bz2b200_destroy(c);
        delete h;
        return;
    }
    cudaSetDevice(h->device);
    for (std::thread &t : h->lane_thr)
        if (t.joinable()) t.join();
    for (bz2b200 *l : h->lanes) bz2b200_destroy(l);
    if (h->parent && h->stream) cudaStreamDestroy(h->stream);
    if (h->ctl) {
        for (cudaEvent_t e : h->ctl->ev_bits)
            if (e) cudaEventDestroy(e);
        if (h->ctl->ev_setup) cudaEventDestroy(h->ctl->ev_setup);
        delete h->ctl;
    }
    for (DevBuf *b : h->all)
        if (b->p) cudaFree(b->p);
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    if (h->pin_starts) cudaFreeHost(h->pin_starts);
    if (h->pin_len) cudaFreeHost(h->pin_len);
    if (h->pin_info) cudaFreeHost(h->pin_info);
    if (h->side) cudaStreamDestroy(h->side);
    if (h->copy) cudaStreamDestroy(h->copy);
    if (h->h2d) cudaStreamDestroy(h->h2d);
    for (cudaEvent_t e : h->ev_seg) cudaEventDestroy(e);
    for (cudaEvent_t e : {h->ev_t0, h->ev_t1})
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : {h->ev_tree, h->ev_walk, h->ev_w0, h->ev_w1})
        if (e) cudaEventDestroy(e);
    for (auto &e : h->ev)
        if (e) cudaEventDestroy(e);
    delete h;
}

static int alloc_batch(bz2b200 *h);
constexpr size_t R0_WORDS = 4 * 1024; // >= RC_MAXPASS * 256: digit counters of all round-0 sweeps of one block (tile)

static int create_impl(bz2b200 *h, bool with_batch)
{
    CK(cudaSetDevice(h->device));
    int rc = upload_constants(h);
    if (rc) return rc;
    const uint64_t G = h->G, S = h->S;
    h->tilesR = (int)((S + RS_TILE - 1) / RS_TILE);
    h->tilesG = (int)((S + RG_TILE - 1) / RG_TILE);
    h->tilesM = (int)((S + MTF_TILE - 1) / MTF_TILE);
    h->tilesR2 = (int)((S + R2_TILE - 1) / R2_TILE);
    h->tilesD = (int)((S + HP_SYMS - 1) / HP_SYMS);
    h->sel_stride = ((uint64_t)h->tilesD * HP_THREADS + 63) / 64 * 64;
    h->ctasC = (int)((S + RC_TILES * RS_TILE - 1) / (RC_TILES * RS_TILE));
    ENSURE(h->scalars, 64 * 8);
    if (with_batch) {
        rc = alloc_batch(h);
        if (rc) return rc;
    }
    CK(cudaMallocHost(&h->h_pinned, 64 * sizeof(uint64_t)));
    CK(cudaMallocHost(&h->pin_starts, (G + 2) * sizeof(uint64_t)));
    CK(cudaMallocHost(&h->pin_len, (G + 2) * sizeof(uint32_t)));
    CK(cudaMallocHost(&h->pin_info, 4 * sizeof(uint64_t)));
    {
        int lo_p = 0, hi_p = 0; // the walk is one warp on the critical path of every call: its stream goes first
        if (cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p) != cudaSuccess) { (void)cudaGetLastError(); hi_p = 0; }
        CK(cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, hi_p));
    }
    CK(cudaStreamCreateWithFlags(&h->copy, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h->h2d, cudaStreamNonBlocking));
    CK(cudaEventCreate(&h->ev_t0));
    CK(cudaEventCreate(&h->ev_t1));
    CK(cudaEventCreateWithFlags(&h->ev_tree, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_walk, cudaEventDisableTiming));
    CK(cudaEventCreate(&h->ev_w0));
    CK(cudaEventCreate(&h->ev_w1));
    for (auto &e : h->ev) CK(cudaEventCreate(&e));
    return BZ2B200_OK;
}

// the per-batch buffers (fixed size: G blocks of stride S)
static int alloc_batch(bz2b200 *h)
{
    const uint64_t G = h->G, S = h->S;
    ENSURE(h->rle, G * S + 64);
    ENSURE(h->L, G * S + 64);
    ENSURE(h->ranks, G * S + 64);
    ENSURE(h->kA, G * S * 8);
    ENSURE(h->kB, G * S * 8);
    ENSURE(h->loc, G * S * 8);
    ENSURE(h->loc_cnt, G * h->tilesG * 4);
    ENSURE(h->rank, G * S * 4);
    ENSURE(h->mtf, G * S * 2);
    ENSURE(h->hist, (size_t)R0_WORDS * G * h->tilesR * 4); // status words: one region per sweep of round 0 (5 x 256 or 4 x 1024 bins)
    ENSURE(h->tickets, (size_t)RC_MAXPASS * G * 4);
    ENSURE(h->digit_base, G * (size_t)R0_WORDS * 4);
    ENSURE(h->rc_partial, G * h->ctasC * (size_t)R0_WORDS * 4);
    ENSURE(h->cnts, 2 * G * 4); // bag sizes, double buffered
    ENSURE(h->h0, G * 4);
    ENSURE(h->root, G * 4);
    ENSURE(h->work, 2 * G * (size_t)h->tilesG * 4); // two lists of region ids (this round's, the next round's)
    ENSURE(h->lists, G * h->tilesM * 256);
    ENSURE(h->dcount, G * h->tilesM * 4);
    ENSURE(h->r2tiles, G * h->tilesR2 * sizeof(R2Tile));
    ENSURE(h->zin, G * h->tilesR2 * 4);
    ENSURE(h->obase, G * h->tilesR2 * 4);
    ENSURE(h->present, G * 8 * 4);
    ENSURE(h->crc, G * 4);
    ENSURE(h->orig, G * 4);
    ENSURE(h->freq, G * 258 * 4);
    ENSURE(h->m_arr, G * 4);
    ENSURE(h->alpha_arr, G * 4);
    ENSURE(h->hb, G * sizeof(HuffBlock));
    ENSURE(h->lens, G * HUFF_MAXTABLES * HUFF_LEN_STRIDE);
    ENSURE(h->tf, G * HUFF_MAXTABLES * HUFF_MAXALPHA * 4);
    ENSURE(h->codes, G * HUFF_MAXTABLES * HUFF_MAXALPHA * 4);
    ENSURE(h->sel, G * h->sel_stride);
    ENSURE(h->selpos, G * h->sel_stride);
    ENSURE(h->selbits, G * 4);
    ENSURE(h->selbitoff, G * h->sel_stride * 4);
    ENSURE(h->pertab, G * HUFF_MAXTABLES * 4);
    ENSURE(h->tabbits, G * 4);
    ENSURE(h->databits, G * 4);
    ENSURE(h->nbits_b, G * 8);
    ENSURE(h->bitoff, G * 8);
    ENSURE(h->datastart, G * 8);
    ENSURE(h->tilebits, G * h->tilesD * 4);
    h->batch_ready = true;
    return BZ2B200_OK;
}

// lanes of a fresh parent: G / n_lanes blocks each, on their own streams
static int create_lanes(bz2b200 *h)
{
    const uint32_t nl = (uint32_t)std::min<uint64_t>((uint64_t)h->n_lanes, h->G);
    if (nl < 2) return BZ2B200_OK;
    h->ctl = new LaneCtl();
    for (auto &e : h->ctl->ev_bits) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ctl->ev_setup, cudaEventDisableTiming));
    for (uint32_t i = 0; i < nl; ++i) {
        bz2b200 *l = new bz2b200();
        h->lanes.push_back(l);
        l->parent = h;
        l->device = h->device; l->base = h->base; l->mult = h->mult; l->bs = h->bs; l->S = h->S; l->nbits = h->nbits;
        l->key0_bytes = h->key0_bytes; l->pack_keys = h->pack_keys; l->onesweep = h->onesweep; l->local_sort = h->local_sort; l->root_sort = h->root_sort; l->r0_group = h->r0_group; l->r0_bits = h->r0_bits; l->bulk_copy = h->bulk_copy; l->n_lanes = 1;
        l->G = (h->G + nl - 1) / nl;
        int rc = create_impl(l, true);
        if (rc == BZ2B200_OK && cudaStreamCreateWithFlags(&l->stream, cudaStreamNonBlocking) != cudaSuccess) rc = BZ2B200_E_CUDA;
        if (rc) { h->err = l->err; return rc; }
    }
    h->lane_thr.resize(nl);
    h->lane_busy.assign(nl, 0);
    h->lane_nb.assign(nl, 0);
    return BZ2B200_OK;
}

extern "C" int bz2b200_create(bz2b200_t **out, int device_id, int base_block_size, int multiplier, int max_blocks_in_flight)
{
    if (!out) return BZ2B200_E_ARG;
    *out = nullptr;
    if (multiplier < 1 || multiplier > 9 || base_block_size < 1000 || base_block_size > 100000) return BZ2B200_E_BLOCKSIZE;
    if (max_blocks_in_flight < 1) return BZ2B200_E_PARALLEL;
    int ndev = bz2b200_device_count();
    if (ndev <= 0 || device_id < 0 || device_id >= ndev) return BZ2B200_E_CUDA;
    bz2b200 *h = new bz2b200();
    h->device = device_id;
    h->base = base_block_size;
    h->mult = multiplier;
    h->bs = base_block_size * multiplier;
    h->G = (uint32_t)max_blocks_in_flight;
    h->S = ((uint64_t)h->bs + 2 + 63) / 64 * 64;
    h->nbits = ilog2_ceil((uint64_t)h->bs + 1);
    if (const char *e = getenv("BZ2B200_KEY0_BYTES")) {
        int k = atoi(e);
        if (k >= 1 && k <= 5) h->key0_bytes = k;
    }
    if (const char *e = getenv("BZ2B200_SEG_MB")) {
        int a = 0, b = 0;
        if (sscanf(e, "%d,%d", &a, &b) == 2 && a > 0 && b > 0) { h->seg_first = (uint64_t)a << 20; h->seg_rest = (uint64_t)b << 20; }
    }
    if (const char *e = getenv("BZ2B200_PACK")) h->pack_keys = atoi(e);
    if (const char *e = getenv("BZ2B200_LOCALSORT")) h->local_sort = atoi(e);
    if (const char *e = getenv("BZ2B200_ROOT")) h->root_sort = atoi(e);
    if (const char *e = getenv("BZ2B200_BULK")) h->bulk_copy = atoi(e);
    if (const char *e = getenv("BZ2B200_CHAIN_OVERLAP")) h->chain_overlap = atoi(e);
    if (const char *e = getenv("BZ2B200_WALK_SMEM")) h->walk_smem = std::max(0, std::min(200, atoi(e)));
    if (const char *e = getenv("BZ2B200_R0_GROUP")) h->r0_group = std::max(0, atoi(e));
    if (const char *e = getenv("BZ2B200_R0_BITS")) h->r0_bits = atoi(e) == 10 ? 10 : 8;
    if (const char *e = getenv("BZ2B200_ONESWEEP")) h->onesweep = atoi(e);
    if (const char *e = getenv("BZ2B200_WALK")) h->prefix_walk = atoi(e);
    if (const char *e = getenv("BZ2B200_WALK_GROWTH")) h->walk_growth = std::max(1, atoi(e));
    if (const char *e = getenv("BZ2B200_WALK_FIRST")) h->walk_first = std::max(1, atoi(e));
    if (!h->prefix_walk) h->walk_growth = std::min(h->walk_growth, 4);
    if (const char *e = getenv("BZ2B200_TRACE")) h->trace = atoi(e);
    if (const char *e = getenv("BZ2B200_LANES")) h->n_lanes = std::max(1, std::min(4, atoi(e)));
    h->lanes_wanted = h->n_lanes > 1 && h->G > 1;
    int rc = create_impl(h, false); // per-batch buffers come with the first call that needs them: the lanes' at the first compress, the handle's own at the first decompress
    if (rc != BZ2B200_OK) {
        // keep the message reachable through a static for the failing create
        static thread_local std::string last;
        last = h->err;
        fprintf(stderr, "bz2b200_create: %s\n", last.c_str());
        bz2b200_destroy(h);
        return rc;
    }
    *out = h;
    return BZ2B200_OK;
}

extern "C" size_t bz2b200_max_output_bytes(const bz2b200_t *h, size_t n)
{
    if (!h) return 0;
    // RLE1 expands by at most 5/4, Huffman codes are at most 20 bits per symbol but a block never codes worse than
    // ~8.6 bits per RLE1 byte with 258 symbols; keep a generous bound: 1.3 n + per-block tables/selectors.
    size_t blocks = n / (size_t)std::max(1, (h->bs - 5) * 4 / 5) + 2;
    return n + n / 3 + blocks * (size_t)(2048 + h->bs / 50) + 4096;
}

struct Scalars { // layout of h->scalars (u64 each)
    enum { BITPOS = 1, BATCH_BITS = 2, WALK_INFO = 4 /* 4 entries */, ROUND = 8 /* kept rotations, max bag, sum of bags, regions with elements */, TICKET = 12 };
};

static void radix_pass(bz2b200 *h, const uint64_t *kin, uint64_t *kout, const uint32_t *counts, int shift, int bits, uint32_t nb, uint32_t max_count,
                       uint64_t total_elems)
{
    uint32_t tiles = (max_count + RS_TILE - 1) / RS_TILE;
    if (tiles == 0) return;
    dim3 grid(tiles, nb);
    uint32_t mask = (1u << bits) - 1u;
    KL(KID_k_radix_hist, k_radix_hist<uint64_t><<<grid, RH_THREADS, 0, h->stream>>>(kin, counts, nullptr, h->S, shift, mask, ptr<uint32_t>(h->hist), h->tilesR));
    KL(KID_k_radix_scan, k_radix_scan<<<nb, 256, 0, h->stream>>>(ptr<uint32_t>(h->hist), counts, h->tilesR, ptr<uint32_t>(h->digit_base)));
    KL(KID_k_radix_scatter, k_radix_scatter<<<grid, RS_THREADS, sizeof(ScatterSmem), h->stream>>>(kin, kout, counts, nullptr, h->S, shift, mask, ptr<uint32_t>(h->hist),
                                                                                                h->tilesR, ptr<uint32_t>(h->digit_base)));
    h->stats.kernel_launches += 3;
    KBYTES(KID_k_radix_hist, total_elems * 8);
    KBYTES(KID_k_radix_scatter, total_elems * 16);
}

// sorts bits [VAL_BITS, VAL_BITS + total_bits) of the elements; returns 0 if the result is in k0, 1 if in k1
static int radix_sort(bz2b200 *h, uint64_t *k0, uint64_t *k1, const uint32_t *counts, int total_bits, uint32_t nb, uint32_t max_count, uint64_t total_elems)
{
    int npass = (total_bits + 7) / 8;
    if (h->onesweep && total_bits % 8 == 0 && npass <= RC_MAXPASS && max_count > 0) {
        // digit totals of every pass in one read, then one sweep (read + write of every key) per digit
        const uint32_t nc = (max_count + RC_TILES * RS_TILE - 1) / (RC_TILES * RS_TILE), tiles = (max_count + RS_TILE - 1) / RS_TILE;
        uint32_t *ticket = reinterpret_cast<uint32_t *>(ptr<uint64_t>(h->scalars) + Scalars::TICKET);
        KL(KID_k_radix_count, k_radix_count<<<dim3(nc, nb), RH_THREADS, 0, h->stream>>>(k0, counts, h->S, npass, ptr<uint32_t>(h->rc_partial), h->ctasC));
        KL(KID_k_radix_bases, k_radix_bases<<<nb, 256, 0, h->stream>>>(ptr<uint32_t>(h->rc_partial), counts, npass, h->ctasC, nb, ptr<uint32_t>(h->digit_base), 1, nullptr));
        h->stats.kernel_launches += 2;
        KBYTES(KID_k_radix_count, total_elems * 8);
        int cur = 0;
        for (int p = 0; p < npass; ++p) {
            // status words of the tiles in use only (the bags of the later rounds fill a few tiles per block)
            cudaMemset2DAsync(h->hist.p, (size_t)h->tilesR * 1024, 0, (size_t)tiles * 1024, nb, h->stream);
            cudaMemsetAsync(ticket, 0, 4, h->stream);
            KL(KID_k_radix_onesweep, k_radix_onesweep<false><<<dim3(tiles, nb), RS_THREADS, sizeof(ScatterSmem), h->stream>>>(
                                         cur ? k1 : k0, cur ? k0 : k1, counts, h->S, VAL_BITS + 8 * p, ptr<uint32_t>(h->digit_base) + (size_t)p * nb * 256,
                                         ptr<uint32_t>(h->hist), h->tilesR, ticket, nullptr, 0, nullptr));
            h->stats.kernel_launches += 1;
            KBYTES(KID_k_radix_onesweep, total_elems * 16);
            cur ^= 1;
        }
        return cur;
    }
    int basebits = total_bits / npass, extra = total_bits % npass;
    int shift = VAL_BITS, cur = 0;
    for (int p = 0; p < npass; ++p) {
        int bits = basebits + (p < extra ? 1 : 0);
        if (cur == 0) radix_pass(h, k0, k1, counts, shift, bits, nb, max_count, total_elems);
        else radix_pass(h, k1, k0, counts, shift, bits, nb, max_count, total_elems);
        cur ^= 1;
        shift += bits;
    }
    return cur;
}

// Round 0: all rotations by their first 4 bytes.  Every pass's digit is one byte of the same cyclic text, so one byte histogram
// gives the digit totals of all four passes, and the first pass builds its elements straight from the text.  Result in kA.
static void radix_sort_round0(bz2b200 *h, const uint32_t *d_n, uint32_t nb, uint32_t max_n, uint64_t sum_n)
{
    const uint32_t nc = (max_n + RC_TILES * RS_TILE - 1) / (RC_TILES * RS_TILE), tiles = (max_n + RS_TILE - 1) / RS_TILE;
    uint32_t *ticket = reinterpret_cast<uint32_t *>(ptr<uint64_t>(h->scalars) + Scalars::TICKET);
    KL(KID_k_bwt_bytehist, k_bwt_bytehist<<<dim3(nc, nb), 256, 0, h->stream>>>(ptr<uint8_t>(h->rle), d_n, h->S, ptr<uint32_t>(h->rc_partial), h->ctasC));
    KL(KID_k_radix_bases, k_radix_bases<<<nb, 256, 0, h->stream>>>(ptr<uint32_t>(h->rc_partial), d_n, 1, h->ctasC, nb, ptr<uint32_t>(h->digit_base), h->key0_bytes, ptr<uint32_t>(h->present)));
    h->stats.kernel_launches += 2;
    KBYTES(KID_k_bwt_bytehist, sum_n);
    uint64_t *kA = ptr<uint64_t>(h->kA), *kB = ptr<uint64_t>(h->kB);
    for (int p = 0; p < h->key0_bytes; ++p) {
        cudaMemsetAsync(h->hist.p, 0, (size_t)nb * h->tilesR * 256 * 4, h->stream);
        cudaMemsetAsync(ticket, 0, 4, h->stream);
        uint32_t *db = ptr<uint32_t>(h->digit_base) + (size_t)p * nb * 256;
        if (p == 0)
            KL(KID_k_radix_onesweep, k_radix_onesweep<true><<<dim3(tiles, nb), RS_THREADS, sizeof(ScatterSmem), h->stream>>>(
                                         nullptr, kB, d_n, h->S, VAL_BITS, db, ptr<uint32_t>(h->hist), h->tilesR, ticket, ptr<uint8_t>(h->rle), h->key0_bytes, nullptr));
        else
            KL(KID_k_radix_onesweep, k_radix_onesweep<false><<<dim3(tiles, nb), RS_THREADS, sizeof(ScatterSmem), h->stream>>>(
                                         (p & 1) ? kB : kA, (p & 1) ? kA : kB, d_n, h->S, VAL_BITS + 8 * p, db, ptr<uint32_t>(h->hist), h->tilesR, ticket, nullptr, 0, nullptr));
        h->stats.kernel_launches += 1;
        KBYTES(KID_k_radix_onesweep, sum_n * (p == 0 ? 9 : 16));
    }
}

static float ev_ms(cudaEvent_t a, cudaEvent_t b)
{
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

static int is_device_ptr(const void *p)
{
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// Stage A over the whole input, part 1: the tree of summaries.  The input becomes available in segments (seg_end[]); the
// tree is extended and the sequential walk over blocks runs in chunks on a side stream (walk_extend / walk_launch /
// walk_collect), ahead of the compression kernels on the main stream.
static int walk_setup(bz2b200 *h, const uint8_t *d_in, uint64_t n, int is_final)
{
    h->scan_valid = false;
    h->scan_first = 0;
    h->h_starts.assign(1, 0);
    h->h_rle_len.clear();
    h->w_in = d_in;
    h->w_n = 0;
    h->caller_final = is_final;
    h->w_final = 0;
    h->w_found = 0;
    h->w_finished = true; // nothing available yet: walk_extend brings the first segment in
    h->seg_avail = 0;
    h->tiles_done = 0;
    if (n == 0) { h->seg_end.clear(); return BZ2B200_OK; }
    uint64_t c = (n + RLE_TILE - 1) / RLE_TILE;
    ENSURE(h->pstate, (c + 1) * sizeof(PState));
    if (h->prefix_walk) ENSURE(h->ltab, (c + 1) * 32 * 8); // lane prefix table: 256 B per 2 KiB tile
    for (int l = 0; l < RLE_MAX_LEVELS; ++l) {
        ENSURE(h->seg[l], c * sizeof(Seg));
        if (l > 0) ENSURE(h->segP[l], c * sizeof(Seg));
        if (c <= 1) break;
        c = (c + RLE_FANOUT - 1) / RLE_FANOUT;
    }
    h->w_cap = n / (uint64_t)std::max(1, (h->bs - 5) * 4 / 5) + 2;
    ENSURE(h->starts, (h->w_cap + 2) * 8);
    ENSURE(h->rle_len, (h->w_cap + 2) * 4);
    uint64_t *d_info = ptr<uint64_t>(h->scalars) + Scalars::WALK_INFO;
    CK(cudaMemsetAsync(d_info, 0, 32, h->stream));
    CK(cudaMemsetAsync(d_info + 12, 0, 16, h->stream)); // the global machine's state at the walk position (fresh)
    CK(cudaEventRecord(h->ev_tree, h->stream));
    CK(cudaStreamWaitEvent(h->side, h->ev_tree, 0));
    return BZ2B200_OK;
}

static bool walk_more_input(const bz2b200 *h) { return h->seg_avail < h->seg_end.size(); }

// the next input segment: (host input) its copy, then level-0 summaries of the new tiles and the upper tree levels
static int walk_extend(bz2b200 *h)
{
    size_t k = h->seg_avail;
    if (h->host_src) {
        // pinned memory: every copy was queued up front; pageable memory: cudaMemcpyAsync blocks the host, so copy one segment ahead only
        const size_t upto = h->host_pinned ? h->seg_end.size() : std::min(h->seg_end.size(), k + 2);
        for (; h->seg_copied < upto; ++h->seg_copied) {
            const uint64_t lo = h->seg_copied ? h->seg_end[h->seg_copied - 1] : 0, hi = h->seg_end[h->seg_copied];
            if (h->seg_copied == 0) CK(cudaEventRecord(h->ev[0], h->h2d));
            CK(cudaMemcpyAsync(ptr<uint8_t>(h->in) + lo, h->host_src + lo, hi - lo, cudaMemcpyDefault, h->h2d));
            CK(cudaEventRecord(h->ev_seg[h->seg_copied], h->h2d));
            if (h->seg_copied + 1 == h->seg_end.size()) CK(cudaEventRecord(h->ev[1], h->h2d));
        }
        while (k + 1 < h->seg_copied && cudaEventQuery(h->ev_seg[k + 1]) == cudaSuccess) ++k; // take everything that has already arrived
        (void)cudaGetLastError();
        CK(cudaStreamWaitEvent(h->side, h->ev_seg[k], 0));
    }
    const uint64_t n = h->seg_end[k];
    uint64_t cnt[RLE_MAX_LEVELS];
    int nl = 1;
    cnt[0] = (n + RLE_TILE - 1) / RLE_TILE;
    while (cnt[nl - 1] > 1 && nl < RLE_MAX_LEVELS) { cnt[nl] = (cnt[nl - 1] + RLE_FANOUT - 1) / RLE_FANOUT; ++nl; }
    const uint64_t t0 = h->tiles_done, nt = cnt[0] - t0;
    CK(cudaEventRecord(h->ev_t0, h->side));
    k_rle_tile_seg<<<(unsigned)((nt + RLE_WARPS - 1) / RLE_WARPS), RLE_WARPS * 32, 0, h->side>>>(h->w_in, n, ptr<Seg>(h->seg[0]), t0, cnt[0], h->prefix_walk ? ptr<uint64_t>(h->ltab) : nullptr);
    for (int l = 1; l < nl; ++l)
        k_rle_reduce<<<(unsigned)((cnt[l] + RLE_WARPS - 1) / RLE_WARPS), RLE_WARPS * 32, 0, h->side>>>(ptr<Seg>(h->seg[l - 1]), cnt[l - 1], ptr<Seg>(h->seg[l]), cnt[l]);
    if (h->prefix_walk) {
        // top-down: summaries of everything in front of every node; level 0 as machine states
        for (int l = nl - 2; l >= 0; --l)
            k_rle_prefix<<<(unsigned)((cnt[l + 1] + RLE_WARPS - 1) / RLE_WARPS), RLE_WARPS * 32, 0, h->side>>>(
                ptr<Seg>(h->seg[l]), cnt[l], l + 1 == nl - 1 ? nullptr : ptr<Seg>(h->segP[l + 1]), cnt[l + 1], l == 0 ? nullptr : ptr<Seg>(h->segP[l]),
                l == 0 ? ptr<PState>(h->pstate) : nullptr);
        if (nl == 1)
            k_rle_prefix<<<1, RLE_WARPS * 32, 0, h->side>>>(ptr<Seg>(h->seg[0]), cnt[0], nullptr, 1, nullptr, ptr<PState>(h->pstate));
        h->stats.kernel_launches += nl > 1 ? nl - 1 : 1;
    }
    CK(cudaEventRecord(h->ev_t1, h->side));
    h->stats.kernel_launches += nl;
    for (int l = 0; l < RLE_MAX_LEVELS; ++l) { h->wl.seg[l] = l < nl ? ptr<Seg>(h->seg[l]) : nullptr; h->wl.cnt[l] = l < nl ? cnt[l] : 0; }
    h->wl.n_levels = nl;
    h->tiles_done = n / RLE_TILE; // a partial last tile is summarised again when more bytes arrive
    h->w_n = n;
    h->seg_avail = k + 1;
    h->w_final = (h->caller_final && h->seg_avail == h->seg_end.size()) ? 1 : 0;
    h->w_finished = false;
    h->tree_pending = true;
    return BZ2B200_OK;
}

// side stream: find up to `max_new` (<= G) further blocks and bring their starts / lengths to pinned host memory
static int walk_launch(bz2b200 *h, uint64_t max_new)
{
    uint64_t *d_info = ptr<uint64_t>(h->scalars) + Scalars::WALK_INFO;
    const uint64_t b0 = h->w_found;
    h->w_req = max_new;
    CK(cudaEventRecord(h->ev_w0, h->side));
    if (h->prefix_walk)
        k_rle_walk_prefix<<<1, 32, h->chain_walk ? (size_t)h->walk_smem << 10 : 0, h->side>>>(h->w_in, h->w_n, (uint64_t)h->bs - 6, h->w_final, h->wl, ptr<PState>(h->pstate), ptr<uint64_t>(h->starts),
                                                  ptr<uint32_t>(h->rle_len), h->w_cap, d_info, max_new, ptr<uint64_t>(h->ltab));
    else
        k_rle_walk<<<1, 32, 0, h->side>>>(h->w_in, h->w_n, (uint64_t)h->bs - 6, h->w_final, h->wl, ptr<uint64_t>(h->starts), ptr<uint32_t>(h->rle_len), h->w_cap,
                                           d_info, max_new);
    CK(cudaEventRecord(h->ev_w1, h->side));
    h->stats.kernel_launches += 1;
    CK(cudaMemcpyAsync(h->pin_info, d_info, 32, cudaMemcpyDeviceToHost, h->side));
    const uint64_t room = h->w_cap + 1 - std::min(b0, h->w_cap + 1); // entries of starts[] from b0 that exist
    const uint64_t ns = std::min<uint64_t>(max_new + 1, room + 1), nl = std::min<uint64_t>(max_new, room);
    CK(cudaMemcpyAsync(h->pin_starts, ptr<uint64_t>(h->starts) + b0, ns * 8, cudaMemcpyDeviceToHost, h->side));
    if (nl) CK(cudaMemcpyAsync(h->pin_len, ptr<uint32_t>(h->rle_len) + b0, nl * 4, cudaMemcpyDeviceToHost, h->side));
    CK(cudaEventRecord(h->ev_walk, h->side));
    return BZ2B200_OK;
}

static int walk_collect(bz2b200 *h)
{
    CK(cudaEventSynchronize(h->ev_walk));
    CK(cudaGetLastError());
    const uint64_t nb = h->pin_info[0];
    if (nb > h->w_cap || nb < h->w_found || nb - h->w_found > h->w_req) { h->err = "internal: block walk out of bounds"; return BZ2B200_E_CUDA; }
    const uint64_t fresh = nb - h->w_found;
    h->h_starts.resize(nb + 1);
    h->h_rle_len.resize(nb);
    for (uint64_t i = 0; i <= fresh; ++i) h->h_starts[h->w_found + i] = h->pin_starts[i];
    for (uint64_t i = 0; i < fresh; ++i) h->h_rle_len[h->w_found + i] = h->pin_len[i];
    h->w_found = nb;
    h->w_finished = h->pin_info[2] != 0;
    if (h->trace) fprintf(stderr, "[trace %8.2f] walk: %llu blocks known (input to %llu MiB)%s\n", now_ms() - h->t_call, (unsigned long long)nb, (unsigned long long)(h->w_n >> 20), h->w_finished ? " chunk ran out of input" : "");
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev_w0, h->ev_w1);
    h->stats.ms_walk += ms;
    if (h->profiling) { h->k_ms[KID_k_rle_walk] += ms; h->k_launches[KID_k_rle_walk] += 1; }
    if (h->tree_pending) { // the walk ran behind the tree update on the same stream, so that has completed as well
        float tms = 0;
        cudaEventElapsedTime(&tms, h->ev_t0, h->ev_t1);
        h->stats.ms_walk += tms;
        if (h->profiling) { h->k_ms[KID_k_rle_tile_seg] += tms; h->k_launches[KID_k_rle_tile_seg] += 1; }
        h->tree_pending = false;
    }
    return BZ2B200_OK;
}

// resolves `input` to a 16-byte aligned device pointer and cuts it into segments.  Device input is one segment (copied once
// if misaligned); host input is copied segment by segment on the H2D stream while earlier segments are being compressed.
static int stage_input(bz2b200 *h, const uint8_t *input, size_t n, const uint8_t **d_in)
{
    h->seg_end.clear();
    h->host_src = nullptr;
    h->seg_copied = 0;
    if (n == 0) { *d_in = nullptr; return BZ2B200_OK; }
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, input);
    if (e != cudaSuccess) { (void)cudaGetLastError(); at.type = cudaMemoryTypeUnregistered; }
    const bool dev = at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
    const bool remote = at.type == cudaMemoryTypeDevice && at.device != h->device; // another GPU's memory: copied like host input, over NVLink
    if (dev && !remote) {
        h->seg_end.push_back(n);
        if ((reinterpret_cast<uintptr_t>(input) & 15u) == 0) { *d_in = input; return BZ2B200_OK; }
        ENSURE(h->in, n + 64);
        CK(cudaMemcpyAsync(h->in.p, input, n, cudaMemcpyDeviceToDevice, h->stream));
        *d_in = ptr<uint8_t>(h->in);
        return BZ2B200_OK;
    }
    ENSURE(h->in, n + 64);
    // a short first segment lets the kernels start early; the later ones are short enough that blocks become known at a steady
    // pace (a batch is cut from what the walk has found so far) and long enough to amortise a tree update
    const uint64_t first = h->seg_first, rest = h->seg_rest;
    for (uint64_t pos = 0; pos < n;) {
        pos = std::min<uint64_t>(n, pos + (pos == 0 ? first : rest));
        h->seg_end.push_back(pos);
    }
    while (h->ev_seg.size() < h->seg_end.size()) {
        cudaEvent_t ev;
        CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        h->ev_seg.push_back(ev);
    }
    h->host_src = input;
    h->host_pinned = at.type == cudaMemoryTypeHost || remote;
    *d_in = ptr<uint8_t>(h->in);
    return BZ2B200_OK;
}

// Stages `input` and builds the summary tree over all of it (asynchronously: copies and kernels are only queued).
static int scan_prepare_impl(bz2b200 *h, const uint8_t *input, size_t n_bytes)
{
    h->prepared = false;
    h->stats = bz2b200_stats_t{};
    const uint8_t *d_in = nullptr;
    int rc = stage_input(h, input, n_bytes, &d_in);
    if (rc) return rc;
    rc = walk_setup(h, d_in, n_bytes, 0);
    if (rc) return rc;
    while (walk_more_input(h)) {
        rc = walk_extend(h);
        if (rc) return rc;
    }
    h->prepared = true;
    h->prep_input = input;
    h->prep_n = n_bytes;
    return BZ2B200_OK;
}

static int multi_scan_blocks(bz2b200 *P, const uint8_t *input, size_t n_bytes, int is_final, uint64_t *block_starts, size_t capacity, uint64_t *n_blocks,
                             size_t *consumed);

// The walk over a prepared input from `first_start` on; the block table stays in the handle for bz2b200_compress_scanned.
static int scan_ex_impl(bz2b200 *h, const uint8_t *input, size_t n_bytes, int is_final, uint64_t first_start, uint64_t *block_starts, size_t capacity,
                        uint64_t *n_blocks, size_t *consumed)
{
    if (first_start > n_bytes) { h->err = "first_start behind the input"; return BZ2B200_E_ARG; }
    int rc;
    if (!(h->prepared && h->prep_input == input && h->prep_n == n_bytes)) {
        rc = scan_prepare_impl(h, input, n_bytes);
        if (rc) return rc;
    }
    h->prepared = false; // one scan per preparation: the walk's device state is consumed
    h->caller_final = is_final;
    h->w_final = is_final ? 1 : 0; // every segment has been queued, so the walk sees the whole input
    h->h_starts.assign(1, first_start);
    h->h_rle_len.clear();
    h->w_found = 0;
    h->scan_first = first_start;
    if (n_bytes) {
        uint64_t *d_info = ptr<uint64_t>(h->scalars) + Scalars::WALK_INFO;
        h->h_pinned[40] = 0; h->h_pinned[41] = first_start; h->h_pinned[42] = 0; h->h_pinned[43] = first_start ? 1 : 0;
        CK(cudaMemcpyAsync(d_info, h->h_pinned + 40, 32, cudaMemcpyHostToDevice, h->side));
        h->w_finished = false;
        while (!h->w_finished) {
            rc = walk_launch(h, h->G);
            if (rc) return rc;
            rc = walk_collect(h);
            if (rc) return rc;
        }
    }
    CK(cudaStreamSynchronize(h->h2d));
    if (h->profiling) { CK(cudaStreamSynchronize(h->stream)); prof_collect(h); }
    const uint64_t nb = h->w_found;
    h->scan_valid = true;
    *n_blocks = nb;
    if (consumed) *consumed = (size_t)h->h_starts[nb];
    for (uint64_t i = 0; i < nb && i < capacity && block_starts; ++i) block_starts[i] = h->h_starts[i];
    return (block_starts && nb > capacity) ? BZ2B200_E_OUTPUT : BZ2B200_OK;
}

extern "C" int bz2b200_scan_prepare(bz2b200_t *h, const uint8_t *input, size_t n_bytes)
{
    if (!h || (!input && n_bytes)) return BZ2B200_E_ARG;
    if (!h->children.empty()) return BZ2B200_OK; // a multi-device handle prepares its slices inside the scan
    CK(cudaSetDevice(h->device));
    h->err.clear();
    return scan_prepare_impl(h, input, n_bytes);
}

extern "C" int bz2b200_scan_blocks_ex(bz2b200_t *h, const uint8_t *input, size_t n_bytes, int is_final, uint64_t first_start, uint64_t *block_starts,
                                      size_t capacity, uint64_t *n_blocks, size_t *consumed)
{
    if (!h || (!input && n_bytes) || !n_blocks) return BZ2B200_E_ARG;
    if (!h->children.empty()) {
        if (first_start) { h->err = "first_start is not supported on a multi-device handle"; return BZ2B200_E_ARG; }
        return multi_scan_blocks(h, input, n_bytes, is_final, block_starts, capacity, n_blocks, consumed);
    }
    CK(cudaSetDevice(h->device));
    h->err.clear();
    int rc = scan_ex_impl(h, input, n_bytes, is_final, first_start, block_starts, capacity, n_blocks, consumed);
    if (rc != BZ2B200_OK && rc != BZ2B200_E_OUTPUT) {
        cudaStreamSynchronize(h->side);
        cudaStreamSynchronize(h->h2d);
        (void)cudaGetLastError();
        h->scan_valid = false;
    }
    return rc;
}

extern "C" int bz2b200_scan_blocks(bz2b200_t *h, const uint8_t *input, size_t n_bytes, uint64_t *block_starts, size_t capacity, uint64_t *n_blocks)
{
    if (!block_starts && capacity) return BZ2B200_E_ARG;
    return bz2b200_scan_blocks_ex(h, input, n_bytes, 1, 0, block_starts, capacity, n_blocks, nullptr);
}

// One batch: blocks [b0, b0 + nb) of the call, `seq` = number of the batch inside the call.  h is the handle itself or one of its
// lanes; P owns what the batches share.
static int run_batch(bz2b200 *h, const uint8_t *d_in, uint64_t b0, uint32_t nb, uint32_t *d_out_words, uint64_t seq)
{
    bz2b200 *P = h->parent ? h->parent : h;
    LaneCtl *ctl = h->parent ? P->ctl : nullptr;
    cudaStream_t st = h->stream;
    const uint64_t S = h->S;
    if (P->trace) fprintf(stderr, "[trace %8.2f] batch %llu start: blocks %llu..%llu\n", now_ms() - P->t_call, (unsigned long long)seq, (unsigned long long)b0, (unsigned long long)(b0 + nb));
    if (ctl && h->lane_call != P->call_id) { // first batch of this call on the lane: behind the parent's set-up work
        CK(cudaStreamWaitEvent(st, ctl->ev_setup, 0));
        h->lane_call = P->call_id;
    }
    uint64_t max_len = 0;
    for (uint32_t i = 0; i < nb; ++i) max_len = std::max(max_len, P->h_starts[b0 + i + 1] - P->h_starts[b0 + i]);
    const uint64_t max_tiles = (max_len + RLE_TILE - 1) / RLE_TILE;
    ENSURE(h->segB, (uint64_t)nb * max_tiles * sizeof(Seg));
    ENSURE(h->tstate, (uint64_t)nb * max_tiles * 8);
    uint64_t *d_scal = ptr<uint64_t>(h->scalars);
    const uint64_t *d_starts = ptr<uint64_t>(P->starts);
    const uint32_t *d_n = ptr<uint32_t>(P->rle_len) + b0;
    auto &stt = h->stats;

    CK(cudaMemsetAsync(h->crc.p, 0, (size_t)nb * 4, st));
    CK(cudaMemsetAsync(h->freq.p, 0, (size_t)nb * 258 * 4, st));
    CK(cudaMemsetAsync(h->tf.p, 0, (size_t)nb * HUFF_MAXTABLES * HUFF_MAXALPHA * 4, st));
    CK(cudaMemsetAsync(h->tabbits.p, 0, (size_t)nb * 4, st));
    CK(cudaMemsetAsync(h->databits.p, 0, (size_t)nb * 4, st));

    // ---- stage A: RLE1 + presence + CRC
    CK(cudaEventRecord(h->ev[2], st));
    {
        dim3 grid((unsigned)((max_tiles + RLE_WARPS - 1) / RLE_WARPS), nb);
        KL(KID_k_rle_btile_seg, k_rle_btile_seg<<<grid, RLE_WARPS * 32, 0, st>>>(d_in, d_starts, (uint32_t)b0, ptr<Seg>(h->segB), max_tiles, h->bulk_copy));
        KL(KID_k_rle_btile_scan, k_rle_btile_scan<<<(nb + RLE_WARPS - 1) / RLE_WARPS, RLE_WARPS * 32, 0, st>>>(ptr<Seg>(h->segB), d_starts, (uint32_t)b0, nb, ptr<uint64_t>(h->tstate),
                                                                                       max_tiles));
        KL(KID_k_rle_emit, k_rle_emit<<<grid, RLE_WARPS * 32, 0, st>>>(d_in, d_starts, (uint32_t)b0, ptr<uint64_t>(h->tstate), max_tiles, ptr<uint8_t>(h->rle), S, h->bulk_copy));
        dim3 cgrid((unsigned)((max_len + CRC_CHUNK - 1) / CRC_CHUNK), nb);
        KL(KID_k_crc_chunks, k_crc_chunks<<<cgrid, CRC_THREADS, 0, st>>>(d_in, d_starts, (uint32_t)b0, ptr<uint32_t>(h->crc)));
        KL(KID_k_crc_finish, k_crc_finish<<<(nb + 127) / 128, 128, 0, st>>>(d_starts, (uint32_t)b0, nb, ptr<uint32_t>(h->crc)));
        stt.kernel_launches += 5;
    }
    CK(cudaEventRecord(h->ev[3], st));

    // ---- stage B: BWT
    uint32_t max_n = 0;
    uint64_t sum_n = 0;
    for (uint32_t i = 0; i < nb; ++i) {
        max_n = std::max(max_n, P->h_rle_len[b0 + i]);
        sum_n += P->h_rle_len[b0 + i];
    }
    uint32_t *bag_cnt[2] = {ptr<uint32_t>(h->cnts), ptr<uint32_t>(h->cnts) + (size_t)h->G};
    int res = 0;
    uint32_t *d_h0 = ptr<uint32_t>(h->h0);
    // non-primitive blocks (T = w^m) are sorted through their root w: d_n is the length the SORT works on, d_nfull the block's
    const uint32_t *d_nfull = d_n;
    if (h->root_sort) {
        KL(KID_k_bwt_period, k_bwt_period<<<nb, BP_THREADS, 0, st>>>(ptr<uint8_t>(h->rle), d_nfull, S, ptr<uint32_t>(h->root)));
        stt.kernel_launches += 1;
        d_n = ptr<uint32_t>(h->root);
    }
    if (h->onesweep && h->pack_keys) {
        // keys over each block's own alphabet: byte histogram -> presence bitmap -> packed keys (+ the depth h0 they order by)
        const uint32_t nc = (max_n + RC_TILES * RS_TILE - 1) / (RC_TILES * RS_TILE);
        KL(KID_k_bwt_bytehist, k_bwt_bytehist<<<dim3(nc, nb), 256, 0, st>>>(ptr<uint8_t>(h->rle), d_n, S, ptr<uint32_t>(h->rc_partial), h->ctasC));
        KL(KID_k_radix_bases, k_radix_bases<<<nb, 256, 0, st>>>(ptr<uint32_t>(h->rc_partial), d_n, 1, h->ctasC, nb, ptr<uint32_t>(h->digit_base), 1, ptr<uint32_t>(h->present)));
        const uint32_t tiles = (max_n + RS_TILE - 1) / RS_TILE;
        const bool wide = h->r0_bits == 10;         // 4 sweeps over 10-bit digits instead of 5 over 8-bit ones
        const int r0_pass = wide ? 4 : RC_MAXPASS, r0_bins = wide ? 1024 : 256;
        if (wide) {
            KL(KID_k_bwt_keys0, k_bwt_count_packed<10, 4><<<dim3(std::max(nc, 1u), nb), 256, 0, st>>>(ptr<uint8_t>(h->rle), d_n, S, ptr<uint32_t>(h->present),
                                                                                                     ptr<uint32_t>(h->rc_partial), h->ctasC, d_h0));
            KL(KID_k_radix_bases, k_radix_bases_wide<10><<<nb, 256, 0, st>>>(ptr<uint32_t>(h->rc_partial), d_n, 4, h->ctasC, nb, ptr<uint32_t>(h->digit_base)));
        } else {
            KL(KID_k_bwt_keys0, k_bwt_count_packed<8, RC_MAXPASS><<<dim3(std::max(nc, 1u), nb), 256, 0, st>>>(ptr<uint8_t>(h->rle), d_n, S, ptr<uint32_t>(h->present),
                                                                                                               ptr<uint32_t>(h->rc_partial), h->ctasC, d_h0)); // (h0 is written by the first CTA of every block, empty or not)
            KL(KID_k_radix_bases, k_radix_bases<<<nb, 256, 0, st>>>(ptr<uint32_t>(h->rc_partial), d_n, RC_MAXPASS, h->ctasC, nb, ptr<uint32_t>(h->digit_base), 1, nullptr));
        }
        stt.kernel_launches += 4;
        KBYTES(KID_k_bwt_bytehist, sum_n);
        KBYTES(KID_k_bwt_keys0, sum_n);
        uint64_t *kA = ptr<uint64_t>(h->kA), *kB = ptr<uint64_t>(h->kB);
        // Five sweeps; the first builds its elements from the text.  They run GROUP by group of a few blocks, all five sweeps of a
        // group back to back: a block's elements (2 x 7.2 MB at 900k) then stay in the 126 MB L2 from sweep to sweep, and only the
        // text read of the first and the element write of the last sweep go to HBM.  Every (sweep, group) has its own ticket and every
        // sweep its own region of status words, zeroed once per batch.
        const uint32_t gsz = h->r0_group ? (uint32_t)h->r0_group : nb;
        const size_t pass_words = (size_t)h->G * h->tilesR * r0_bins;
        if (tiles) {
            for (int p = 0; p < r0_pass; ++p) CK(cudaMemsetAsync(ptr<uint32_t>(h->hist) + (size_t)p * pass_words, 0, (size_t)nb * h->tilesR * r0_bins * 4, st));
            CK(cudaMemsetAsync(h->tickets.p, 0, (size_t)RC_MAXPASS * h->G * 4, st));
        }
        uint32_t launch_no = 0;
        for (uint32_t g0 = 0; g0 < nb && tiles; g0 += gsz) {
            const uint32_t gc = std::min(gsz, nb - g0);
            uint32_t gmax = 0;
            for (uint32_t i = 0; i < gc; ++i) gmax = std::max(gmax, P->h_rle_len[b0 + g0 + i]);
            const uint32_t gt = (gmax + RS_TILE - 1) / RS_TILE;
            if (!gt) continue;
            const uint64_t go = (uint64_t)g0 * S;
            for (int p = 0; p < r0_pass; ++p) {
                uint32_t *db = ptr<uint32_t>(h->digit_base) + ((size_t)p * nb + g0) * r0_bins;
                uint32_t *status = ptr<uint32_t>(h->hist) + (size_t)p * pass_words + (size_t)g0 * h->tilesR * r0_bins;
                uint32_t *ticket = ptr<uint32_t>(h->tickets) + launch_no++;
                // sweep p reads the buffer sweep p-1 wrote; the LAST sweep must end in kB (the regroup reads it there)
                const bool to_b = ((r0_pass - 1 - p) & 1) == 0;
                uint64_t *src = (to_b ? kA : kB) + go, *dst = (to_b ? kB : kA) + go;
                if (wide) {
                    if (p == 0)
                        KL(KID_k_radix_onesweep, k_radix_onesweep<true, 10><<<dim3(gt, gc), RS_THREADS, sizeof(ScatterSmemT<10>), st>>>(
                                                     nullptr, dst, d_n + g0, S, VAL_BITS, db, status, h->tilesR, ticket, ptr<uint8_t>(h->rle) + go, 0,
                                                     ptr<uint32_t>(h->present) + (size_t)g0 * 8));
                    else
                        KL(KID_k_radix_onesweep, k_radix_onesweep<false, 10><<<dim3(gt, gc), RS_THREADS, sizeof(ScatterSmemT<10>), st>>>(
                                                     src, dst, d_n + g0, S, VAL_BITS + 10 * p, db, status, h->tilesR, ticket, nullptr, 0, nullptr));
                } else {
                    if (p == 0)
                        KL(KID_k_radix_onesweep, k_radix_onesweep<true><<<dim3(gt, gc), RS_THREADS, sizeof(ScatterSmem), st>>>(
                                                     nullptr, dst, d_n + g0, S, VAL_BITS, db, status, h->tilesR, ticket, ptr<uint8_t>(h->rle) + go, 0,
                                                     ptr<uint32_t>(h->present) + (size_t)g0 * 8));
                    else
                        KL(KID_k_radix_onesweep, k_radix_onesweep<false><<<dim3(gt, gc), RS_THREADS, sizeof(ScatterSmem), st>>>(
                                                     src, dst, d_n + g0, S, VAL_BITS + 8 * p, db, status, h->tilesR, ticket, nullptr, 0, nullptr));
                }
                stt.kernel_launches += 1;
            }
        }
        KBYTES(KID_k_radix_onesweep, sum_n * (9 + (uint64_t)(r0_pass - 1) * 16));
        res = 1; // the last sweep ends in kB
    } else if (h->onesweep && h->key0_bytes >= 4) {
        k_fill_u32<<<(nb + 255) / 256, 256, 0, st>>>(d_h0, nb, (uint32_t)h->key0_bytes);
        radix_sort_round0(h, d_n, nb, max_n, sum_n);
        res = h->key0_bytes & 1; // pass p writes kB for even p: four passes end in kA (res 0), five in kB (res 1)
    } else {
        // (development paths) the byte histogram is still what gives the presence bitmap
        const uint32_t nc = (max_n + RC_TILES * RS_TILE - 1) / (RC_TILES * RS_TILE);
        KL(KID_k_bwt_bytehist, k_bwt_bytehist<<<dim3(nc, nb), 256, 0, st>>>(ptr<uint8_t>(h->rle), d_n, S, ptr<uint32_t>(h->rc_partial), h->ctasC));
        KL(KID_k_radix_bases, k_radix_bases<<<nb, 256, 0, st>>>(ptr<uint32_t>(h->rc_partial), d_n, 1, h->ctasC, nb, ptr<uint32_t>(h->digit_base), 1, ptr<uint32_t>(h->present)));
        k_fill_u32<<<(nb + 255) / 256, 256, 0, st>>>(d_h0, nb, (uint32_t)h->key0_bytes);
        dim3 g((max_n + 255) / 256, nb);
        KL(KID_k_bwt_keys0, k_bwt_keys0<<<g, 256, 0, st>>>(ptr<uint8_t>(h->rle), d_n, S, h->key0_bytes, ptr<uint64_t>(h->kA)));
        stt.kernel_launches += 1;
        res = radix_sort(h, ptr<uint64_t>(h->kA), ptr<uint64_t>(h->kB), d_n, 8 * h->key0_bytes, nb, max_n, sum_n);
    }
    CK(cudaEventRecord(h->ev[4], st));

    RoundArgs A;
    A.n_arr = d_n;
    A.stride = S;
    A.T = ptr<uint8_t>(h->rle);
    A.L = ptr<uint8_t>(h->L);
    A.rank = ptr<uint32_t>(h->rank);
    A.loc = ptr<uint64_t>(h->loc);
    A.loc_cnt = ptr<uint32_t>(h->loc_cnt);
    A.tiles_per_block = h->tilesG;
    A.totals = reinterpret_cast<unsigned long long *>(d_scal + Scalars::ROUND);
    A.split = h->local_sort != 0;
    uint32_t *work_list[2] = {ptr<uint32_t>(h->work), ptr<uint32_t>(h->work) + (size_t)h->G * h->tilesG};
    A.work_out_cnt = reinterpret_cast<uint32_t *>(d_scal + Scalars::ROUND + 3);
    int wcur = 0; // work_list[wcur]: the regions that hold elements for the coming round

    const int round_bits = 2 * VAL_BITS;
    const dim3 gtiles((max_n + RG_TILE - 1) / RG_TILE, nb);
    A.h0 = d_h0;
    uint32_t round_no = 0; // rotations of block b are ordered by h0[b] << round_no bytes
    uint64_t rounds = 0;
    CK(cudaMemsetAsync(h->loc_cnt.p, 0, (size_t)nb * h->tilesG * 4, st));
    // round 0: regroup the whole block, sorted by its first key
    {
        uint64_t *sorted = res ? ptr<uint64_t>(h->kB) : ptr<uint64_t>(h->kA);
        uint64_t *bag = res ? ptr<uint64_t>(h->kA) : ptr<uint64_t>(h->kB);
        A.round = round_no;
        A.work_in = work_list[0];
        A.work_out = work_list[0];
        CK(cudaMemsetAsync(d_scal + Scalars::ROUND, 0, 32, st));
        CK(cudaMemsetAsync(bag_cnt[0], 0, (size_t)nb * 4, st));
        KL(KID_k_bwt_regroup, k_bwt_regroup<true><<<gtiles, RG_THREADS, 0, st>>>(sorted, d_n, A, bag, bag_cnt[0]));
        KL(KID_k_bwt_roundstats, k_bwt_roundstats<<<(nb + 255) / 256, 256, 0, st>>>(bag_cnt[0], nb, A.totals));
        stt.kernel_launches += 2;
        res ^= 1; // the bag is in the other buffer
    }
    int cs = 0; // bag_cnt[cs] = sizes of the current bags
    // algorithmic bytes of the regroup kernels: element in (8) + rank out (4) + T in / L out for final rotations (2) or the
    // kept element out (8); known once the next round's counts are back
    uint64_t prev_regroup_in = sum_n, prev_local_in = 0, prev_active = sum_n;
    for (;;) {
        CK(cudaMemcpyAsync(h->h_pinned, d_scal + Scalars::ROUND, 32, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (h->profiling) prof_collect(h);
        const uint64_t active = h->h_pinned[0], total_b = h->h_pinned[2];
        const uint32_t max_b = (uint32_t)h->h_pinned[1], n_regions = (uint32_t)(h->h_pinned[3] & 0xffffffffull);
        if (h->profiling) {
            const uint64_t in = prev_regroup_in + prev_local_in, fin = prev_active > active ? prev_active - active : 0; // final this round
            const uint64_t kept_r = in ? (uint64_t)((double)active * prev_regroup_in / in) : 0;
            KBYTES(KID_k_bwt_regroup, 12 * prev_regroup_in + 2 * (in ? (uint64_t)((double)fin * prev_regroup_in / in) : 0) + 8 * kept_r);
            KBYTES(KID_k_bwt_local, 12 * prev_local_in + 2 * (in ? (uint64_t)((double)fin * prev_local_in / in) : 0) + 8 * (active - kept_r));
        }
        prev_active = active;
        prev_regroup_in = total_b;
        prev_local_in = active - total_b;
        if (active == 0) break;
        stt.bwt_active_volume += active;
        ++rounds;
        // next round: key = (group position, rank of the rotation sorted_len further on)
        uint64_t *bag = res ? ptr<uint64_t>(h->kB) : ptr<uint64_t>(h->kA);
        uint64_t *other = res ? ptr<uint64_t>(h->kA) : ptr<uint64_t>(h->kB);
        const bool have_local = n_regions > 0;
        CK(cudaMemsetAsync(d_scal + Scalars::ROUND, 0, 32, st));
        A.work_in = work_list[wcur];
        A.work_out = work_list[wcur ^ 1];
        wcur ^= 1;
        if (have_local) {
            KL(KID_k_bwt_gather_local, k_bwt_gather_local<<<n_regions, 256, 0, st>>>(A));
            stt.kernel_launches += 1;
            KBYTES(KID_k_bwt_gather_local, (active - total_b) * 20);
        }
        if (max_b) {
            dim3 gg((max_b + 255) / 256, nb);
            KL(KID_k_bwt_gather, k_bwt_gather<<<gg, 256, 0, st>>>(bag, bag_cnt[cs], d_n, S, ptr<uint32_t>(h->rank), d_h0, round_no));
            stt.kernel_launches += 1;
        }
        A.round = round_no + 1;
        if (have_local) {
            KL(KID_k_bwt_local, k_bwt_local<<<n_regions, LS_THREADS, 0, st>>>(A));
            stt.kernel_launches += 1;
        }
        CK(cudaMemsetAsync(bag_cnt[cs ^ 1], 0, (size_t)nb * 4, st));
        if (max_b) {
            const int r = radix_sort(h, bag, other, bag_cnt[cs], round_bits, nb, max_b, total_b);
            uint64_t *sorted = r ? other : bag, *newbag = r ? bag : other;
            dim3 g((max_b + RG_TILE - 1) / RG_TILE, nb);
            KL(KID_k_bwt_regroup, k_bwt_regroup<false><<<g, RG_THREADS, 0, st>>>(sorted, bag_cnt[cs], A, newbag, bag_cnt[cs ^ 1]));
            stt.kernel_launches += 1;
            res = (newbag == ptr<uint64_t>(h->kB)) ? 1 : 0;
        }
        KL(KID_k_bwt_roundstats, k_bwt_roundstats<<<(nb + 255) / 256, 256, 0, st>>>(bag_cnt[cs ^ 1], nb, A.totals));
        stt.kernel_launches += 1;
        cs ^= 1;
        ++round_no;
    }
    stt.bwt_rounds = std::max<uint64_t>(stt.bwt_rounds, rounds);
    CK(cudaEventRecord(h->ev[5], st));
    {
        KL(KID_k_bwt_finish, k_bwt_finish<<<(nb + 127) / 128, 128, 0, st>>>(ptr<uint32_t>(h->rank), d_nfull, d_n, S, nb, ptr<uint32_t>(h->orig)));
        stt.kernel_launches += 1;
        if (h->root_sort) {
            KL(KID_k_bwt_expand, k_bwt_expand_copy<<<dim3((BP_MAX + 255) / 256, nb), 256, 0, st>>>(ptr<uint8_t>(h->L), d_nfull, d_n, S, ptr<uint8_t>(h->ranks)));
            KL(KID_k_bwt_expand, k_bwt_expand<<<dim3(64, nb), 256, 0, st>>>(ptr<uint8_t>(h->L), d_nfull, d_n, S, ptr<uint8_t>(h->ranks)));
            stt.kernel_launches += 2;
        }
        d_n = d_nfull;
    }
    CK(cudaEventRecord(h->ev[6], st));

    // ---- stage C: MTF + RLE2
    {
        dim3 gm((max_n + MW_WARPS * MTF_TILE - 1) / (MW_WARPS * MTF_TILE), nb);
        dim3 gq((max_n + MR_WARPS * MTF_TILE - 1) / (MR_WARPS * MTF_TILE), nb);
        KL(KID_k_mtf_recency, k_mtf_recency<<<gq, MR_WARPS * 32, 0, st>>>(ptr<uint8_t>(h->L), d_n, S, ptr<uint32_t>(h->present), ptr<uint64_t>(h->lists),
                                                                          ptr<uint32_t>(h->dcount), h->tilesM));
        KL(KID_k_mtf_compose, k_mtf_compose<<<nb, MC_WARPS * 32, 0, st>>>(d_n, ptr<uint64_t>(h->lists), ptr<uint32_t>(h->dcount), h->tilesM, nb));
        KL(KID_k_mtf_ranks, k_mtf_ranks<<<gm, MW_WARPS * 32, 0, st>>>(ptr<uint8_t>(h->L), d_n, S, ptr<uint32_t>(h->present), ptr<uint64_t>(h->lists), h->tilesM,
                                                                   ptr<uint8_t>(h->ranks), ptr<R2Tile>(h->r2tiles)));
        dim3 gr((max_n + R2_TILE - 1) / R2_TILE, nb);
        KL(KID_k_rle2_scan, k_rle2_scan<<<(nb + 127) / 128, 128, 0, st>>>(ptr<R2Tile>(h->r2tiles), d_n, h->tilesR2, nb, ptr<uint32_t>(h->zin), ptr<uint32_t>(h->obase),
                                                       ptr<uint32_t>(h->present), ptr<uint16_t>(h->mtf), S, ptr<uint32_t>(h->freq), ptr<uint32_t>(h->m_arr),
                                                       ptr<uint32_t>(h->alpha_arr)));
        KL(KID_k_rle2_emit, k_rle2_emit<<<gr, R2_THREADS, 0, st>>>(ptr<uint8_t>(h->ranks), d_n, S, ptr<uint32_t>(h->zin), ptr<uint32_t>(h->obase), h->tilesR2, ptr<uint16_t>(h->mtf), S,
                                                ptr<uint32_t>(h->freq)));
        stt.kernel_launches += 5;
        KBYTES(KID_k_mtf_recency, sum_n);
        KBYTES(KID_k_mtf_ranks, 2 * sum_n);
    }
    CK(cudaEventRecord(h->ev[7], st));

    // ---- stage D: Huffman tables
    const uint32_t max_m = max_n + 1;
    dim3 gh((max_m + HP_SYMS - 1) / HP_SYMS, nb);
    {
        KL(KID_k_huff_seeds, k_huff_seeds<<<(nb + 127) / 128, 128, 0, st>>>(ptr<uint32_t>(h->m_arr), ptr<uint32_t>(h->alpha_arr), ptr<uint32_t>(h->freq), nb, ptr<HuffBlock>(h->hb),
                                                        ptr<uint8_t>(h->lens)));
        for (int it = 0; it < 4; ++it) {
            KL(KID_k_huff_pass, k_huff_pass<<<gh, HP_THREADS, 0, st>>>(ptr<uint16_t>(h->mtf), S, ptr<HuffBlock>(h->hb), ptr<uint8_t>(h->lens), ptr<uint32_t>(h->tf), ptr<uint8_t>(h->sel),
                                                    h->sel_stride));
            KL(KID_k_huff_lengths, k_huff_lengths<<<dim3(HUFF_MAXTABLES, nb), HL_THREADS, 0, st>>>(ptr<HuffBlock>(h->hb), ptr<uint32_t>(h->tf), ptr<uint8_t>(h->lens), ptr<uint32_t>(h->databits),
                                                                     it == 3));
        }
        KL(KID_k_huff_codes, k_huff_codes<<<(nb * HUFF_MAXTABLES + 63) / 64, 64, 0, st>>>(ptr<HuffBlock>(h->hb), ptr<uint8_t>(h->lens), ptr<uint32_t>(h->codes), ptr<uint32_t>(h->tabbits),
                                                                      ptr<uint32_t>(h->pertab), nb));
        KL(KID_k_huff_selmtf, k_huff_selmtf<<<nb, SM_THREADS, 0, st>>>(ptr<HuffBlock>(h->hb), ptr<uint8_t>(h->sel), h->sel_stride, ptr<uint8_t>(h->selpos),
                                                                       ptr<uint32_t>(h->selbitoff), ptr<uint32_t>(h->selbits), nb));
        stt.kernel_launches += 11;
    }
    CK(cudaEventRecord(h->ev[8], st));

    // ---- stage E: bit emission at the final stream offsets
    {
        KL(KID_k_block_bits, k_block_bits<<<(nb + 127) / 128, 128, 0, st>>>(ptr<HuffBlock>(h->hb), ptr<uint32_t>(h->present), ptr<uint32_t>(h->selbits), ptr<uint32_t>(h->tabbits),
                                                        ptr<uint32_t>(h->databits), nb, ptr<uint64_t>(h->nbits_b)));
        if (ctl) { // the stream's bit position passes from batch to batch in order
            std::unique_lock<std::mutex> lk(ctl->mu);
            ctl->cv.wait(lk, [&] { return ctl->bits_issued == seq || ctl->rc != 0; });
            if (ctl->rc) return ctl->rc;
            if (seq) CK(cudaStreamWaitEvent(st, ctl->ev_bits[(seq - 1) & 3], 0));
        }
        KL(KID_k_bit_offsets, k_bit_offsets<<<1, 1024, 0, st>>>(ptr<uint64_t>(h->nbits_b), nb, ptr<uint64_t>(P->scalars) + Scalars::BITPOS, ptr<uint64_t>(h->bitoff),
                                                                d_scal + Scalars::BATCH_BITS, P->out_cap_bits));
        if (ctl) {
            CK(cudaEventRecord(ctl->ev_bits[seq & 3], st));
            std::lock_guard<std::mutex> lk(ctl->mu);
            ctl->bits_issued = seq + 1;
            ctl->cv.notify_all();
        }
        KL(KID_k_emit_header, k_emit_header<<<nb, EH_THREADS, 0, st>>>(ptr<HuffBlock>(h->hb), ptr<uint32_t>(h->present), ptr<uint32_t>(h->crc), ptr<uint32_t>(h->orig),
                                                                       ptr<uint8_t>(h->selpos), ptr<uint32_t>(h->selbitoff), h->sel_stride, ptr<uint32_t>(h->selbits),
                                                                       ptr<uint32_t>(h->pertab), ptr<uint8_t>(h->lens), ptr<uint64_t>(h->bitoff), nb, d_out_words,
                                                                       ptr<uint64_t>(h->datastart)));
        KL(KID_k_data_tilebits, k_data_tilebits<<<gh, HP_THREADS, 0, st>>>(ptr<uint16_t>(h->mtf), S, ptr<HuffBlock>(h->hb), ptr<uint8_t>(h->lens), ptr<uint8_t>(h->sel), h->sel_stride,
                                                    ptr<uint32_t>(h->tilebits), h->tilesD));
        KL(KID_k_data_scan, k_data_scan<<<(nb + 127) / 128, 128, 0, st>>>(ptr<HuffBlock>(h->hb), ptr<uint32_t>(h->tilebits), h->tilesD, nb));
        KL(KID_k_emit_data, k_emit_data<<<gh, HP_THREADS, 0, st>>>(ptr<uint16_t>(h->mtf), S, ptr<HuffBlock>(h->hb), ptr<uint32_t>(h->codes), ptr<uint8_t>(h->sel), h->sel_stride,
                                                ptr<uint32_t>(h->tilebits), h->tilesD, ptr<uint64_t>(h->datastart), d_out_words));
        stt.kernel_launches += 6;
        CK(cudaMemcpyAsync(ptr<uint32_t>(P->all_crc) + b0, h->crc.p, (size_t)nb * 4, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(ptr<uint64_t>(P->all_nbits) + b0, h->nbits_b.p, (size_t)nb * 8, cudaMemcpyDeviceToDevice, st));
    }
    CK(cudaEventRecord(h->ev[9], st));
    CK(cudaMemcpyAsync(h->h_pinned + 16, d_scal + Scalars::BATCH_BITS, 8, cudaMemcpyDeviceToHost, st)); // bit position behind this batch
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    if (h->profiling) prof_collect(h);
    if (P->trace) fprintf(stderr, "[trace %8.2f] batch %llu done\n", now_ms() - P->t_call, (unsigned long long)seq);
    if (ctl) { // batches finish in order: everything in front of this one has been written (and drained)
        std::unique_lock<std::mutex> lk(ctl->mu);
        ctl->cv.wait(lk, [&] { return ctl->done == seq || ctl->rc != 0; });
        if (ctl->rc) return ctl->rc;
    }
    if (P->drain_dst) {
        // every byte below the current bit position is final: ship it now, overlapped with the next batch
        const size_t done = (size_t)(h->h_pinned[16] >> 3);
        if (done > P->drain_cap) { h->err = "output buffer too small"; return BZ2B200_E_OUTPUT; }
        if (done > P->drained) {
            CK(cudaMemcpyAsync(P->drain_dst + P->drained, ptr<uint8_t>(P->out) + P->drained, done - P->drained,
                               P->drain_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, P->copy));
            P->drained = done;
        }
    }
    if (ctl) {
        std::lock_guard<std::mutex> lk(ctl->mu);
        ctl->done = seq + 1;
        ctl->cv.notify_all();
    }
    stt.ms_rle += ev_ms(h->ev[2], h->ev[3]);
    stt.ms_bwt_sort0 += ev_ms(h->ev[3], h->ev[4]);
    stt.ms_bwt_rounds += ev_ms(h->ev[4], h->ev[5]);
    stt.ms_bwt_finish += ev_ms(h->ev[5], h->ev[6]);
    stt.ms_mtf += ev_ms(h->ev[6], h->ev[7]);
    stt.ms_huffman += ev_ms(h->ev[7], h->ev[8]);
    stt.ms_emit += ev_ms(h->ev[8], h->ev[9]);
    h->last_b0 = b0;
    h->last_nb = nb;
    return BZ2B200_OK;
}

// scanned: compress the blocks found by the last bz2b200_scan_blocks_ex (input staged, block table in the handle) instead of staging and
// walking `input`; keep_dev: leave the packed bits in the handle's own device buffer (h->out) -- multi-device streams move them from there
// from: the first block starts at input[from->first_start] (a slice of a longer stream); from->on_walked is called with the offset of
// the first unfinished block as soon as the walk has been through the slice -- long before the blocks are compressed
struct FromArgs {
    uint64_t first_start = 0;
    void (*on_walked)(uint64_t next_start, void *user) = nullptr;
    void *user = nullptr;
};
static int compress_blocks_impl(bz2b200_t *h, const uint8_t *input, size_t n_bytes, int is_final, unsigned start_bit, uint8_t *out,
                                size_t out_capacity, uint64_t *out_nbits, size_t *consumed, uint32_t *block_crcs, uint64_t *block_nbits,
                                size_t block_capacity, uint64_t *n_blocks, bool scanned, bool keep_dev, const FromArgs *from = nullptr);
static int multi_compress_blocks(bz2b200 *P, const uint8_t *input, size_t n_bytes, int is_final, unsigned start_bit, uint8_t *out, size_t out_capacity,
                                 uint64_t *out_nbits, size_t *consumed, uint32_t *block_crcs, uint64_t *block_nbits, size_t block_capacity, uint64_t *n_blocks);

// ---- lanes: one host thread per batch in flight
static void lane_main(bz2b200 *l, int idx, const uint8_t *d_in, uint64_t b0, uint32_t nb, uint32_t *d_out_words, uint64_t seq)
{
    bz2b200 *P = l->parent;
    int rc = cudaSetDevice(l->device) == cudaSuccess ? BZ2B200_OK : BZ2B200_E_CUDA;
    if (rc == BZ2B200_OK) rc = run_batch(l, d_in, b0, nb, d_out_words, seq);
    std::lock_guard<std::mutex> lk(P->ctl->mu);
    if (rc && !P->ctl->rc) { P->ctl->rc = rc; P->ctl->err = l->err; }
    P->lane_busy[idx] = 0;
    P->ctl->cv.notify_all();
}

// hands the batch to a free lane (waits for one); with profiling on, batches run one at a time so that per-kernel times are clean
// (wait == false: returns LANES_BUSY instead of waiting when no lane is free)
enum { LANES_BUSY = 1 };
static int lane_dispatch(bz2b200 *h, const uint8_t *d_in, uint64_t b0, uint32_t nb, uint32_t *d_out_words, uint64_t seq, bool wait = true)
{
    LaneCtl *ctl = h->ctl;
    int idx = -1;
    {
        std::unique_lock<std::mutex> lk(ctl->mu);
        auto free_lane = [&] {
            if (ctl->rc) return true;
            size_t busy = 0;
            for (char b : h->lane_busy) busy += b ? 1 : 0;
            return h->profiling ? busy == 0 : busy < h->lane_busy.size();
        };
        if (!wait && !free_lane()) return LANES_BUSY;
        ctl->cv.wait(lk, free_lane);
        if (ctl->rc) { h->err = ctl->err; return ctl->rc; }
        for (size_t i = 0; i < h->lane_busy.size(); ++i) {
            const size_t k = (seq + i) % h->lane_busy.size(); // alternate, so that a profiled pass uses every lane too
            if (!h->lane_busy[k]) { idx = (int)k; break; }
        }
        h->lane_busy[idx] = 1;
        h->lane_nb[idx] = nb;
    }
    if (h->lane_thr[idx].joinable()) h->lane_thr[idx].join();
    bz2b200 *l = h->lanes[idx];
    l->profiling = h->profiling;
    h->last_lane = idx;
    ++h->call_batches;
    try {
        h->lane_thr[idx] = std::thread(lane_main, l, idx, d_in, b0, nb, d_out_words, seq);
    } catch (...) { // no thread to be had: the batch runs here (its predecessors are on their way, so the ordered hand-overs still work)
        lane_main(l, idx, d_in, b0, nb, d_out_words, seq);
    }
    return BZ2B200_OK;
}

// waits for every batch in flight, folds the lanes' counters into the parent's
static int lanes_finish(bz2b200 *h)
{
    if (h->lanes.empty()) return BZ2B200_OK;
    for (std::thread &t : h->lane_thr)
        if (t.joinable()) t.join();
    auto &s = h->stats;
    for (bz2b200 *l : h->lanes) {
        auto &q = l->stats;
        s.ms_rle += q.ms_rle; s.ms_bwt_sort0 += q.ms_bwt_sort0; s.ms_bwt_rounds += q.ms_bwt_rounds; s.ms_bwt_finish += q.ms_bwt_finish;
        s.ms_mtf += q.ms_mtf; s.ms_huffman += q.ms_huffman; s.ms_emit += q.ms_emit;
        s.bwt_rounds = std::max(s.bwt_rounds, q.bwt_rounds);
        s.bwt_active_volume += q.bwt_active_volume;
        s.kernel_launches += q.kernel_launches;
        q = bz2b200_stats_t{};
        for (int i = 0; i < KID_COUNT; ++i) {
            h->k_ms[i] += l->k_ms[i]; h->k_launches[i] += l->k_launches[i]; h->k_bytes[i] += l->k_bytes[i];
            l->k_ms[i] = 0; l->k_launches[i] = 0; l->k_bytes[i] = 0;
        }
    }
    if (h->call_batches) { // the batch that stays resident for bz2b200_stage_dump
        bz2b200 *last = h->lanes[h->last_lane];
        h->last_b0 = last->last_b0;
        h->last_nb = last->last_nb;
    }
    if (h->ctl->rc) { h->err = h->ctl->err; return h->ctl->rc; }
    return BZ2B200_OK;
}

static int compress_blocks_guarded(bz2b200_t *h, const uint8_t *input, size_t n_bytes, int is_final, unsigned start_bit, uint8_t *out,
                                   size_t out_capacity, uint64_t *out_nbits, size_t *consumed, uint32_t *block_crcs, uint64_t *block_nbits,
                                   size_t block_capacity, uint64_t *n_blocks, bool scanned, bool keep_dev, const FromArgs *from = nullptr)
{
    int rc = compress_blocks_impl(h, input, n_bytes, is_final, start_bit, out, out_capacity, out_nbits, consumed, block_crcs, block_nbits,
                                  block_capacity, n_blocks, scanned, keep_dev, from);
    if (h) {
        // never leave batches, copies into the caller's buffer or a walk in flight, whatever the outcome
        (void)lanes_finish(h);
        cudaStreamSynchronize(h->copy);
        cudaStreamSynchronize(h->side);
        cudaStreamSynchronize(h->h2d);
        h->drain_dst = nullptr;
        h->host_src = nullptr;
        h->scan_valid = false; // a block table is good for one compress call
        h->chain_walk = false;
        if (rc != BZ2B200_OK) (void)cudaGetLastError();
    }
    return rc;
}

extern "C" int bz2b200_compress_blocks(bz2b200_t *h, const uint8_t *input, size_t n_bytes, int is_final, unsigned start_bit, uint8_t *out,
                                       size_t out_capacity, uint64_t *out_nbits, size_t *consumed, uint32_t *block_crcs, uint64_t *block_nbits,
                                       size_t block_capacity, uint64_t *n_blocks)
{
    if (h && !h->children.empty())
        return multi_compress_blocks(h, input, n_bytes, is_final, start_bit, out, out_capacity, out_nbits, consumed, block_crcs, block_nbits, block_capacity,
                                     n_blocks);
    return compress_blocks_guarded(h, input, n_bytes, is_final, start_bit, out, out_capacity, out_nbits, consumed, block_crcs, block_nbits, block_capacity,
                                   n_blocks, false, false);
}

extern "C" int bz2b200_compress_blocks_from(bz2b200_t *h, const uint8_t *input, size_t n_bytes, int is_final, uint64_t first_start, unsigned start_bit,
                                            uint8_t *out, size_t out_capacity, uint64_t *out_nbits, size_t *consumed, uint32_t *block_crcs,
                                            uint64_t *block_nbits, size_t block_capacity, uint64_t *n_blocks, bz2b200_walked_fn on_walked, void *user)
{
    if (!h) return BZ2B200_E_ARG;
    if (!h->children.empty()) { h->err = "compress_blocks_from works on a single-device handle"; return BZ2B200_E_ARG; }
    FromArgs fa;
    fa.first_start = first_start;
    fa.on_walked = on_walked;
    fa.user = user;
    return compress_blocks_guarded(h, input, n_bytes, is_final, start_bit, out, out_capacity, out_nbits, consumed, block_crcs, block_nbits, block_capacity,
                                   n_blocks, false, false, &fa);
}

extern "C" int bz2b200_compress_scanned(bz2b200_t *h, unsigned start_bit, uint8_t *out, size_t out_capacity, uint64_t *out_nbits, uint32_t *block_crcs,
                                        uint64_t *block_nbits, size_t block_capacity, uint64_t *n_blocks)
{
    if (!h) return BZ2B200_E_ARG;
    if (!h->children.empty()) { h->err = "compress_scanned works on a single-device handle"; return BZ2B200_E_ARG; }
    size_t consumed = 0;
    return compress_blocks_guarded(h, nullptr, 0, 0, start_bit, out, out_capacity, out_nbits, &consumed, block_crcs, block_nbits, block_capacity, n_blocks,
                                   true, false);
}

static int compress_blocks_impl(bz2b200_t *h, const uint8_t *input, size_t n_bytes, int is_final, unsigned start_bit, uint8_t *out,
                                size_t out_capacity, uint64_t *out_nbits, size_t *consumed, uint32_t *block_crcs, uint64_t *block_nbits,
                                size_t block_capacity, uint64_t *n_blocks, bool scanned, bool keep_dev, const FromArgs *from)
{
    if (!h) return BZ2B200_E_ARG;
    if ((!input && n_bytes) || !out_nbits || !consumed || !n_blocks || start_bit > 7 || (!out && out_capacity)) {
        h->err = "bad argument";
        return BZ2B200_E_ARG;
    }
    CK(cudaSetDevice(h->device));
    h->err.clear();
    if (scanned && !h->scan_valid) { h->err = "no block table: call bz2b200_scan_blocks_ex first"; return BZ2B200_E_STATE; }
    int rc;
    if (h->lanes_wanted && h->lanes.empty()) {
        rc = create_lanes(h);
        if (rc) return rc;
    }
    if (h->lanes.empty() && !h->batch_ready) {
        rc = alloc_batch(h);
        if (rc) return rc;
    }
    h->stats = bz2b200_stats_t{};
    h->last_nb = 0;
    h->call_batches = 0;
    h->t_call = now_ms();
    *out_nbits = 0; *consumed = 0; *n_blocks = 0;
    cudaStream_t st = h->stream;
    const uint8_t *d_in = nullptr;
    if (scanned) {
        d_in = h->w_in;
        n_bytes = (size_t)(h->h_starts[h->w_found] - h->scan_first); // the bytes the scanned blocks cover
        CK(cudaEventRecord(h->ev[10], st));
    } else if (h->prepared && h->prep_input == input && h->prep_n == n_bytes) {
        // bz2b200_scan_prepare has staged this input and queued its summary tree: only the walk is left
        h->prepared = false;
        d_in = h->w_in;
        h->caller_final = is_final;
        h->w_final = is_final ? 1 : 0;
        h->w_finished = n_bytes == 0;
        CK(cudaEventRecord(h->ev[10], st));
    } else {
        h->prepared = false;
        rc = stage_input(h, input, n_bytes, &d_in);
        if (rc) return rc;
        CK(cudaEventRecord(h->ev[10], st));
        rc = walk_setup(h, d_in, n_bytes, is_final);
        if (rc) return rc;
    }
    if (from && from->first_start > n_bytes) { h->err = "first_start behind the input"; return BZ2B200_E_ARG; }
    if (from && !scanned && n_bytes) { // the walk starts inside the buffer (k_rle_walk_prefix derives the machine's state there)
        uint64_t *d_info = ptr<uint64_t>(h->scalars) + Scalars::WALK_INFO;
        h->h_pinned[40] = 0; h->h_pinned[41] = from->first_start; h->h_pinned[42] = 0; h->h_pinned[43] = from->first_start ? 1 : 0;
        CK(cudaMemcpyAsync(d_info, h->h_pinned + 40, 32, cudaMemcpyHostToDevice, h->side));
        h->h_starts.assign(1, from->first_start);
        h->scan_first = from->first_start;
    }
    const uint64_t cap_blocks = n_bytes / (uint64_t)std::max(1, (h->bs - 5) * 4 / 5) + 2;
    const size_t need = bz2b200_max_output_bytes(h, n_bytes) + 8;
    const size_t dcap = (need + 3) / 4 * 4;
    ENSURE(h->out, dcap + 16);
    ENSURE(h->all_crc, (cap_blocks + 1) * 4);
    ENSURE(h->all_nbits, (cap_blocks + 1) * 8);
    uint64_t *d_scal = ptr<uint64_t>(h->scalars);
    h->h_pinned[32] = start_bit;
    CK(cudaMemcpyAsync(d_scal + Scalars::BITPOS, h->h_pinned + 32, 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(h->out.p, 0, dcap, st));
    h->out_cap_bits = (uint64_t)dcap * 8;
    h->drain_dst = keep_dev ? nullptr : out;
    h->drain_cap = out_capacity;
    h->drained = 0;
    h->drain_dev = out ? is_device_ptr(out) != 0 : false;
    const bool laned = !h->lanes.empty();
    uint64_t seq = 0;
    if (laned) {
        h->h_starts.reserve(h->w_cap + 2); // the lanes read the block table while the walk appends to it: no reallocation
        h->h_rle_len.reserve(h->w_cap + 2);
        ++h->call_id;
        h->ctl->bits_issued = 0; h->ctl->done = 0; h->ctl->rc = 0; h->ctl->err.clear();
        for (bz2b200 *l : h->lanes) l->stats = bz2b200_stats_t{};
        CK(cudaEventRecord(h->ctl->ev_setup, st));
    }
    const uint64_t Gb = laned ? h->lanes[0]->G : h->G; // blocks per batch
    {
        // the walk runs one chunk ahead on the side stream; a small first chunk keeps its exposed latency short
        bool pending = false;
        uint64_t b0 = 0, chunk = std::max<uint64_t>(1, std::min<uint64_t>(h->G, (uint64_t)h->walk_first) / (uint64_t)h->walk_growth);
        // A slice of a longer stream (from->on_walked): the device behind this one waits for the walk's end, so the walk never waits
        // for a lane -- batches are handed out only while a lane is free until the walk is through and its end has been passed on.
        bool walk_first = from && from->on_walked && !scanned;
        h->chain_walk = walk_first;
        if (walk_first && !h->host_src && !h->chain_overlap) chunk = h->G; // resident input: the walk goes through in as few launches as possible
        if (scanned && from && from->on_walked) from->on_walked(h->h_starts[h->w_found], from->user);
        while (pending || b0 < h->w_found || !h->w_finished || walk_more_input(h)) {
            if (pending) {
                rc = walk_collect(h);
                if (rc) return rc;
                pending = false;
            }
            if (h->w_finished && walk_more_input(h)) { // the walk ran out of bytes: bring the next input segment in
                rc = walk_extend(h);
                if (rc) return rc;
            }
            if (!h->w_finished) {
                chunk = std::min<uint64_t>(h->G, chunk * (uint64_t)h->walk_growth);
                rc = walk_launch(h, chunk);
                if (rc) return rc;
                pending = true;
            }
            if (walk_first && !pending && h->w_finished && !walk_more_input(h)) { // the walk has seen the whole slice
                walk_first = false;
                from->on_walked(h->h_starts[h->w_found], from->user);
            }
            bool go = b0 < h->w_found;
            // (until the chain has been handed on, batches start only while input is still on its way from the host -- the walk waits
            // for those bytes anyway; with the whole slice resident the walk is a few milliseconds alone on the device, and several
            // times that next to the compression kernels)
            if (go && walk_first && (!laned || !(h->chain_overlap || (h->host_src && walk_more_input(h))))) go = false;
            if (go && laned) {
                // A batch costs about the same wall time whatever its size (its one-warp-per-block kernels), so a lane is not spent
                // on a few blocks while the others have real work: wait until half a batch is known -- unless nothing more will
                // come, or what is in flight is small itself (the first blocks of a host buffer: the GPU is mostly idle then).
                const uint64_t undis = h->w_found - b0;
                const bool nothing_more = !pending && h->w_finished && !walk_more_input(h);
                if (undis < (Gb + 1) / 2 && !nothing_more) {
                    std::lock_guard<std::mutex> lk(h->ctl->mu);
                    uint64_t inflight = 0;
                    for (size_t i = 0; i < h->lane_busy.size(); ++i) inflight += h->lane_busy[i] ? h->lane_nb[i] : 0u;
                    go = inflight < (Gb + 3) / 4;
                }
            }
            if (go) {
                uint32_t cnt = (uint32_t)std::min<uint64_t>(Gb, h->w_found - b0);
                rc = laned ? lane_dispatch(h, d_in, b0, cnt, ptr<uint32_t>(h->out), seq, !walk_first) : run_batch(h, d_in, b0, cnt, ptr<uint32_t>(h->out), seq);
                if (rc == LANES_BUSY) continue; // (the walk goes on; the loop comes back when its next chunk has been collected)
                if (rc) return rc;
                b0 += cnt;
                ++seq;
            }
        }
        if (walk_first) from->on_walked(h->h_starts[h->w_found], from->user); // (nothing to walk)
    }
    rc = lanes_finish(h);
    if (rc) return rc;
    const uint64_t nb = h->w_found, used = h->h_starts[nb];
    h->last_total_blocks = nb;
    *n_blocks = nb;
    *consumed = (size_t)used;
    h->stats.n_input = used - h->scan_first;
    h->stats.n_blocks = nb;
    if ((block_crcs || block_nbits) && nb > block_capacity) { h->err = "block_capacity too small"; return BZ2B200_E_OUTPUT; }
    CK(cudaEventRecord(h->ev[11], st));
    CK(cudaMemcpyAsync(h->h_pinned, d_scal + Scalars::BITPOS, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const uint64_t endbit = h->h_pinned[0];
    const uint64_t bits = endbit - start_bit;
    const size_t out_bytes = (size_t)((endbit + 7) / 8);
    h->stats.ms_total = ev_ms(h->ev[10], h->ev[11]);
    h->stats.n_out_bits = bits;
    if (out_bytes > dcap) { h->err = "internal: output bound exceeded"; return BZ2B200_E_CUDA; } // (nothing was written past the buffer: k_bit_offsets)
    if (!keep_dev && out_bytes > out_capacity) { h->err = "output buffer too small"; return BZ2B200_E_OUTPUT; }
    *out_nbits = bits;
    if (out_bytes && !keep_dev) {
        // the tail that was not drained while the batches ran (at least the last partial byte)
        CK(cudaStreamSynchronize(h->copy));
        CK(cudaEventRecord(h->ev[12], st));
        if (out_bytes > h->drained)
            CK(cudaMemcpyAsync(out + h->drained, ptr<uint8_t>(h->out) + h->drained, out_bytes - h->drained,
                               h->drain_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(h->ev[13], st));
        CK(cudaStreamSynchronize(st));
        h->stats.ms_d2h = ev_ms(h->ev[12], h->ev[13]);
    }
    h->drain_dst = nullptr;
    if (h->host_src && h->seg_copied == h->seg_end.size() && !h->seg_end.empty()) {
        CK(cudaStreamSynchronize(h->h2d));
        h->stats.ms_h2d = ev_ms(h->ev[0], h->ev[1]);
    }
    if (block_crcs && nb) CK(cudaMemcpy(block_crcs, h->all_crc.p, nb * 4, cudaMemcpyDeviceToHost));
    if (block_nbits && nb) CK(cudaMemcpy(block_nbits, h->all_nbits.p, nb * 8, cudaMemcpyDeviceToHost));
    {
        uint64_t sum = 0;
        for (uint32_t v : h->h_rle_len) sum += v;
        h->stats.n_rle = sum;
    }
    return BZ2B200_OK;
}

// ---------------------------------------------------------------------------------------------------------------- several GPUs, one stream
// blocks share no state, so a stream is cut into contiguous byte slices, one per device, and every device compresses
// the blocks that START in its slice.  What is sequential is only the question where those blocks start: the machine of
// BlockCompressor::write (BlockCompressor.hpp:69-96) runs through the whole stream.  Every device therefore stages its slice plus the
// HEAD bytes in front of it (no unfinished block is longer than HEAD raw bytes) and summarises them at once, all devices in parallel;
// the walks then run as a chain: device d starts at the position where device d-1's last complete block ended, finds its blocks and
// hands the start of its unfinished last block on to device d+1 (a number), and compresses while the chain moves on.  Slices shrink
// geometrically so that devices further down the chain, which start later, finish together with the first ones.  Every device packs
// its blocks from bit 0 of its own buffer; when the bit lengths are known each part is moved to its bit offset on its own device
// (k_shift_bits) and copied to its place in the caller's buffer.  The reference does all of this on one device picked in
// OutputStream.hpp:61 and batches in :190-240.
struct MultiChain {
    std::mutex mu;
    std::condition_variable cv;
    std::vector<uint64_t> start;  // absolute stream offset where device d's first block starts
    std::vector<char> known;
    int rc = 0;
    std::string err;
};

struct MultiPart {
    uint64_t base = 0, lo = 0, hi = 0; // staged bytes [base, hi), blocks starting in [.., hi) are this device's
    uint64_t nbits = 0, nblocks = 0, first = 0;
    std::vector<uint32_t> crcs;
    std::vector<uint64_t> nbits_b;
    std::vector<uint64_t> starts; // multi_scan_blocks only
    uint8_t tail[2] = {0, 0};
    bz2b200_stats_t stats{};
};

static uint64_t multi_head_bytes(const bz2b200 *c)
{
    // an unfinished block holds at most (limit/5 + 1) full 255-byte chunks of a run plus what short runs add
    const uint64_t limit = (uint64_t)c->bs - 6;
    return ((limit / 5 + 2) * 255 + 2 * (uint64_t)c->bs + 4095) / 4096 * 4096;
}

static int multi_plan(bz2b200 *P, size_t n_bytes, std::vector<MultiPart> &parts)
{
    const bz2b200 *c0 = P->children[0];
    size_t k = P->children.size();
    if (P->usable_children > 0) k = std::min<size_t>(k, (size_t)P->usable_children);
    const uint64_t min_slice = 16ull * (uint64_t)c0->bs; // a device is not woken for less
    k = (size_t)std::max<uint64_t>(1, std::min<uint64_t>(k, n_bytes / std::max<uint64_t>(1, min_slice)));
    // device d starts about d walks later than device 0: sizes fall by the ratio walk time / compress time per byte
    double r = 0.95;
    if (const char *e = getenv("BZ2B200_MULTI_RATIO")) { const double v = atof(e); if (v > 0.1 && v <= 1.0) r = v; }
    std::vector<double> w(k);
    double tot = 0;
    for (size_t d = 0; d < k; ++d) { w[d] = d ? w[d - 1] * r : 1.0; tot += w[d]; }
    parts.assign(k, MultiPart());
    const uint64_t head = multi_head_bytes(c0);
    double acc = 0;
    uint64_t lo = 0;
    for (size_t d = 0; d < k; ++d) {
        acc += w[d];
        uint64_t hi = d + 1 == k ? (uint64_t)n_bytes : std::min<uint64_t>(n_bytes, (uint64_t)((double)n_bytes * (acc / tot)) / 4096 * 4096);
        if (hi < lo) hi = lo;
        parts[d].lo = lo;
        parts[d].hi = hi;
        parts[d].base = lo > head ? (lo - head) / 4096 * 4096 : 0;
        lo = hi;
    }
    return BZ2B200_OK;
}

// runs fn(d) on one host thread per part (the calling thread takes part 0)
template <typename F>
static void multi_threads(size_t k, F fn)
{
    std::vector<std::thread> thr;
    for (size_t d = 1; d < k; ++d) thr.emplace_back([&fn, d] { fn(d); });
    fn(0);
    for (std::thread &t : thr) t.join();
}

static void chain_publish(MultiChain &ch, size_t d, uint64_t start, int rc, const std::string &err)
{
    std::lock_guard<std::mutex> lk(ch.mu);
    if (rc && !ch.rc) { ch.rc = rc; ch.err = err; }
    if (d < ch.start.size()) { ch.start[d] = start; ch.known[d] = 1; }
    ch.cv.notify_all();
}

// stage + summarise the part's bytes, wait for the chain, walk; returns the child's rc (chain already informed)
static int multi_scan_part(bz2b200 *P, size_t d, const uint8_t *input, size_t n_bytes, int is_final, std::vector<MultiPart> &parts, MultiChain &ch)
{
    bz2b200 *c = P->children[d];
    MultiPart &pt = parts[d];
    const bool last = d + 1 == parts.size();
    int rc = cudaSetDevice(c->device) == cudaSuccess ? BZ2B200_OK : BZ2B200_E_CUDA;
    c->err.clear();
    c->profiling = P->profiling;
    if (rc == BZ2B200_OK && pt.hi > pt.base) rc = scan_prepare_impl(c, input + pt.base, (size_t)(pt.hi - pt.base));
    uint64_t s = 0;
    {
        std::unique_lock<std::mutex> lk(ch.mu);
        ch.cv.wait(lk, [&] { return ch.known[d] || ch.rc; });
        if (ch.rc && !rc) rc = ch.rc;
        s = ch.start[d];
    }
    uint64_t next = s;
    if (rc == BZ2B200_OK && s < pt.hi) {
        if (s < pt.base) { c->err = "internal: an unfinished block is longer than the staged head"; rc = BZ2B200_E_STATE; }
        uint64_t nb = 0;
        size_t used = 0;
        if (rc == BZ2B200_OK)
            rc = scan_ex_impl(c, input + pt.base, (size_t)(pt.hi - pt.base), (last && is_final) ? 1 : 0, s - pt.base, nullptr, 0, &nb, &used);
        if (rc == BZ2B200_OK) {
            pt.nblocks = nb;
            pt.first = s;
            next = pt.base + used;
        }
    }
    chain_publish(ch, d + 1, next, rc, c->err);
    (void)n_bytes;
    return rc;
}

static int multi_scan_blocks(bz2b200 *P, const uint8_t *input, size_t n_bytes, int is_final, uint64_t *block_starts, size_t capacity, uint64_t *n_blocks,
                             size_t *consumed)
{
    P->err.clear();
    std::vector<MultiPart> parts;
    multi_plan(P, n_bytes, parts);
    MultiChain ch;
    ch.start.assign(parts.size() + 1, 0);
    ch.known.assign(parts.size() + 1, 0);
    ch.known[0] = 1;
    multi_threads(parts.size(), [&](size_t d) {
        int rc = multi_scan_part(P, d, input, n_bytes, is_final, parts, ch);
        if (rc == BZ2B200_OK && parts[d].nblocks) {
            bz2b200 *c = P->children[d];
            parts[d].starts.assign(c->h_starts.begin(), c->h_starts.begin() + (ptrdiff_t)parts[d].nblocks);
            for (uint64_t &v : parts[d].starts) v += parts[d].base;
        }
        cudaStreamSynchronize(P->children[d]->side);
        cudaStreamSynchronize(P->children[d]->h2d);
        P->children[d]->scan_valid = false;
    });
    if (ch.rc) { P->err = ch.err; return ch.rc; }
    uint64_t nb = 0;
    for (const MultiPart &pt : parts)
        for (uint64_t v : pt.starts) {
            if (block_starts && nb < capacity) block_starts[nb] = v;
            ++nb;
        }
    *n_blocks = nb;
    if (consumed) *consumed = (size_t)ch.start[parts.size()];
    return (block_starts && nb > capacity) ? BZ2B200_E_OUTPUT : BZ2B200_OK;
}

static int multi_compress_blocks(bz2b200 *P, const uint8_t *input, size_t n_bytes, int is_final, unsigned start_bit, uint8_t *out, size_t out_capacity,
                                 uint64_t *out_nbits, size_t *consumed, uint32_t *block_crcs, uint64_t *block_nbits, size_t block_capacity, uint64_t *n_blocks)
{
    if ((!input && n_bytes) || !out_nbits || !consumed || !n_blocks || start_bit > 7 || (!out && out_capacity)) {
        P->err = "bad argument";
        return BZ2B200_E_ARG;
    }
    P->err.clear();
    P->stats = bz2b200_stats_t{};
    *out_nbits = 0; *consumed = 0; *n_blocks = 0;
    const double t0 = now_ms();
    std::vector<MultiPart> parts;
    multi_plan(P, n_bytes, parts);
    const size_t k = parts.size();
    if (k == 1) { // too little for a second device: the first one takes the call as it is (walk, copies and kernels overlapped)
        bz2b200 *c = P->children[0];
        c->profiling = P->profiling;
        int rc = compress_blocks_guarded(c, input, n_bytes, is_final, start_bit, out, out_capacity, out_nbits, consumed, block_crcs, block_nbits, block_capacity,
                                         n_blocks, false, false);
        P->stats = c->stats;
        if (rc) P->err = c->err;
        return rc;
    }
    MultiChain ch;
    ch.start.assign(k + 1, 0);
    ch.known.assign(k + 1, 0);
    ch.known[0] = 1;
    // ---- per device: stage + summarise at once; when the chain arrives, walk (handing the chain on as soon as the walk is through) and
    // compress into the device's own buffer
    const uint64_t cap_blocks = n_bytes / (uint64_t)std::max(1, (P->bs - 5) * 4 / 5) + 4;
    multi_threads(k, [&](size_t d) {
        bz2b200 *c = P->children[d];
        MultiPart &pt = parts[d];
        const bool last = d + 1 == k;
        int rc = cudaSetDevice(c->device) == cudaSuccess ? BZ2B200_OK : BZ2B200_E_CUDA;
        c->err.clear();
        c->profiling = P->profiling;
        const uint8_t *src = input + pt.base;
        const size_t len = (size_t)(pt.hi - pt.base);
        if (rc == BZ2B200_OK && len && d > 0) rc = scan_prepare_impl(c, src, len); // (device 0 starts at once: its copies run under its kernels)
        uint64_t s = 0;
        {
            std::unique_lock<std::mutex> lk(ch.mu);
            ch.cv.wait(lk, [&] { return ch.known[d] || ch.rc; });
            if (ch.rc && !rc) rc = ch.rc;
            s = ch.start[d];
        }
        bool published = false;
        struct Hand { MultiChain *ch; size_t d; uint64_t base; bool *published; } hand{&ch, d + 1, pt.base, &published};
        if (rc == BZ2B200_OK && s < pt.hi) {
            if (s < pt.base) { c->err = "internal: an unfinished block is longer than the staged head"; rc = BZ2B200_E_STATE; }
            if (rc == BZ2B200_OK) {
                pt.crcs.resize(cap_blocks);
                pt.nbits_b.resize(cap_blocks);
                FromArgs fa;
                fa.first_start = s - pt.base;
                fa.user = &hand;
                fa.on_walked = [](uint64_t next, void *u) {
                    Hand *hd = static_cast<Hand *>(u);
                    *hd->published = true;
                    chain_publish(*hd->ch, hd->d, hd->base + next, BZ2B200_OK, std::string());
                };
                size_t used = 0;
                uint64_t nb = 0;
                rc = compress_blocks_guarded(c, src, len, (last && is_final) ? 1 : 0, 0, nullptr, 0, &pt.nbits, &used, pt.crcs.data(), pt.nbits_b.data(), cap_blocks, &nb,
                                             false, true, &fa);
                if (rc == BZ2B200_OK) {
                    pt.nblocks = nb;
                    pt.first = s;
                    if (pt.nbits) { // the part's last two bytes: the bits the next part's first byte starts with
                        const uint64_t e = (pt.nbits - 1) >> 3;
                        if (cudaMemcpy(pt.tail, ptr<uint8_t>(c->out) + (e ? e - 1 : 0), e ? 2 : 1, cudaMemcpyDeviceToHost) != cudaSuccess) { c->err = "tail read failed"; rc = BZ2B200_E_CUDA; }
                        if (!e) { pt.tail[1] = pt.tail[0]; pt.tail[0] = 0; }
                    }
                    pt.stats = c->stats;
                }
            }
        } else { // nothing of this slice starts a block, or the chain failed upstream: the preparation is dropped
            cudaStreamSynchronize(c->side);
            cudaStreamSynchronize(c->h2d);
            c->prepared = false;
            c->host_src = nullptr;
        }
        if (!published) chain_publish(ch, d + 1, s, rc, c->err); // nothing walked (or a failure): the chain passes through
        if (rc) chain_publish(ch, k + 1, 0, rc, c->err);
    });
    if (ch.rc) { P->err = ch.err; return ch.rc; }
    // ---- the stream: bit offsets, per-block tables
    uint64_t total_blocks = 0, bitpos = start_bit;
    for (const MultiPart &pt : parts) total_blocks += pt.nblocks;
    *n_blocks = total_blocks;
    *consumed = (size_t)ch.start[k];
    if ((block_crcs || block_nbits) && total_blocks > block_capacity) { P->err = "block_capacity too small"; return BZ2B200_E_OUTPUT; }
    {
        uint64_t b = 0;
        for (const MultiPart &pt : parts) {
            for (uint64_t i = 0; i < pt.nblocks; ++i, ++b) {
                if (block_crcs) block_crcs[b] = pt.crcs[i];
                if (block_nbits) block_nbits[b] = pt.nbits_b[i];
            }
        }
    }
    std::vector<uint64_t> off(k, 0);
    for (size_t d = 0; d < k; ++d) { off[d] = bitpos; bitpos += parts[d].nbits; }
    const uint64_t bits = bitpos - start_bit;
    const size_t out_bytes = (size_t)((bitpos + 7) / 8);
    if (bits && out_bytes > out_capacity) { P->err = "output buffer too small"; return BZ2B200_E_OUTPUT; }
    // ---- every part to its place: shifted on its own device, copied by its own device
    int last_part = -1;
    for (size_t d = 0; d < k; ++d)
        if (parts[d].nbits) last_part = (int)d;
    int rc_all = BZ2B200_OK;
    std::mutex emu;
    const bool out_dev = out && is_device_ptr(out);
    multi_threads(k, [&](size_t d) {
        bz2b200 *c = P->children[d];
        const MultiPart &pt = parts[d];
        if (!pt.nbits) return;
        auto fail = [&](const char *what) {
            std::lock_guard<std::mutex> lk(emu);
            if (!rc_all) { rc_all = BZ2B200_E_CUDA; P->err = std::string(what) + ": " + cudaGetErrorString(cudaGetLastError()); }
        };
        if (cudaSetDevice(c->device) != cudaSuccess) return fail("cudaSetDevice");
        const uint32_t shift = (uint32_t)(off[d] & 7);
        uint32_t carry = 0;
        if (shift) { // the last `shift` bits of the part in front (every non-empty part holds at least a block header)
            int p = (int)d - 1;
            while (p >= 0 && !parts[p].nbits) --p;
            if (p >= 0) {
                const MultiPart &q = parts[p];
                const uint64_t e = (q.nbits - 1) >> 3;
                const uint32_t v16 = ((uint32_t)q.tail[0] << 8) | q.tail[1];
                const uint32_t have = (uint32_t)(q.nbits - 8 * (e ? e - 1 : 0)) + (e ? 0u : 8u); // bits of v16 (from its top) that belong to the part
                carry = ((v16 >> (16 - have)) & ((1u << shift) - 1u)) << (8 - shift);
            }
        }
        const uint64_t n_src = (pt.nbits + 7) / 8, n_dst = (shift + pt.nbits + 7) / 8;
        const uint8_t *src = ptr<uint8_t>(c->out);
        const uint8_t *from = src;
        cudaStream_t st = c->stream;
        if (shift) {
            if (ensure(c, c->shifted, n_dst + 16) != BZ2B200_OK) return fail("cudaMalloc");
            k_shift_bits<<<(unsigned)((n_dst + 16 * 256 - 1) / (16 * 256)), 256, 0, st>>>(src, n_src, ptr<uint8_t>(c->shifted), n_dst, shift, carry);
            from = ptr<uint8_t>(c->shifted);
        }
        // the byte a later part starts in belongs to that part (it carries these bits in)
        const bool shares_last = (int)d != last_part && ((off[d] + pt.nbits) & 7) != 0;
        const uint64_t n_copy = n_dst - (shares_last ? 1 : 0);
        if (n_copy && cudaMemcpyAsync(out + (off[d] >> 3), from, n_copy, out_dev ? cudaMemcpyDefault : cudaMemcpyDeviceToHost, st) != cudaSuccess) return fail("cudaMemcpyAsync");
        if (cudaStreamSynchronize(st) != cudaSuccess) return fail("cudaStreamSynchronize");
    });
    if (rc_all) return rc_all;
    *out_nbits = bits;
    // ---- counters of the call
    auto &s = P->stats;
    for (const MultiPart &pt : parts) {
        const auto &q = pt.stats;
        s.ms_rle += q.ms_rle; s.ms_bwt_sort0 += q.ms_bwt_sort0; s.ms_bwt_rounds += q.ms_bwt_rounds; s.ms_bwt_finish += q.ms_bwt_finish;
        s.ms_mtf += q.ms_mtf; s.ms_huffman += q.ms_huffman; s.ms_emit += q.ms_emit; s.ms_walk += q.ms_walk;
        s.ms_h2d = std::max(s.ms_h2d, q.ms_h2d);
        s.n_rle += q.n_rle;
        s.bwt_rounds = std::max(s.bwt_rounds, q.bwt_rounds);
        s.bwt_active_volume += q.bwt_active_volume;
        s.kernel_launches += q.kernel_launches + (pt.nbits ? 1 : 0);
    }
    s.ms_total = (float)(now_ms() - t0); // host clock: the devices have no common event timeline
    s.n_input = ch.start[k];
    s.n_blocks = total_blocks;
    s.n_out_bits = bits;
    return BZ2B200_OK;
}

extern "C" int bz2b200_create_multi(bz2b200_t **out, const int *device_ids, int n_devices, int base_block_size, int multiplier, int max_blocks_in_flight)
{
    if (!out) return BZ2B200_E_ARG;
    *out = nullptr;
    if (!device_ids || n_devices < 1 || n_devices > 64) return BZ2B200_E_ARG;
    bz2b200 *P = new bz2b200();
    P->device = device_ids[0];
    P->base = base_block_size; P->mult = multiplier; P->bs = base_block_size * multiplier; P->G = (uint32_t)std::max(1, max_blocks_in_flight);
    for (int i = 0; i < n_devices; ++i) {
        bz2b200 *c = nullptr;
        int rc = bz2b200_create(&c, device_ids[i], base_block_size, multiplier, max_blocks_in_flight);
        if (rc != BZ2B200_OK) { bz2b200_destroy(P); return rc; }
        P->children.push_back(c);
    }
    // parts leave over NVLink when the caller's buffers live on one of the devices
    for (int i = 0; i < n_devices; ++i)
        for (int j = 0; j < n_devices; ++j)
            if (device_ids[i] != device_ids[j]) {
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, device_ids[i], device_ids[j]) == cudaSuccess && can) {
                    cudaSetDevice(device_ids[i]);
                    if (cudaDeviceEnablePeerAccess(device_ids[j], 0) != cudaSuccess) (void)cudaGetLastError();
                }
            }
    (void)cudaGetLastError();
    if (const char *e = getenv("BZ2B200_MULTI_MAX")) P->usable_children = std::max(0, atoi(e));
    *out = P;
    return BZ2B200_OK;
}

extern "C" int bz2b200_device_count_of(const bz2b200_t *h) { return !h ? 0 : h->children.empty() ? 1 : (int)h->children.size(); }

// One part of a multi-device / multi-process stream to its bit offset: dst = src shifted right by `shift` bits, `carry`
// (already positioned in the top `shift` bits) in front.  src and dst are device pointers on the handle's device.
extern "C" int bz2b200_shift_bits(bz2b200_t *h, const uint8_t *src, uint64_t nbits, unsigned shift, unsigned carry, uint8_t *dst, size_t dst_capacity)
{
    if (!h || shift > 7 || (nbits && (!src || !dst))) return BZ2B200_E_ARG;
    if (!h->children.empty()) h = h->children[0];
    CK(cudaSetDevice(h->device));
    const uint64_t n_src = (nbits + 7) / 8, n_dst = (shift + nbits + 7) / 8;
    if (n_dst > dst_capacity) { h->err = "output buffer too small"; return BZ2B200_E_OUTPUT; }
    if (!n_dst) return BZ2B200_OK;
    k_shift_bits<<<(unsigned)((n_dst + 16 * 256 - 1) / (16 * 256)), 256, 0, h->stream>>>(src, n_src, dst, n_dst, shift, carry & 0xffu);
    CK(cudaStreamSynchronize(h->stream));
    return BZ2B200_OK;
}

// pinned host memory for the caller's staging buffers (asynchronous copies need it; OutputStream double-buffers with it)
extern "C" void *bz2b200_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void bz2b200_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

// blocks per device batch that keep the device busy without asking for more than a third of its free memory
extern "C" int bz2b200_recommended_blocks(int device_id, int base_block_size, int multiplier)
{
    if (multiplier < 1 || multiplier > 9 || base_block_size < 1000 || base_block_size > 100000) return 0;
    int ndev = bz2b200_device_count();
    if (device_id < 0 || device_id >= ndev) return 0;
    size_t fr = 0, tot = 0;
    if (cudaSetDevice(device_id) != cudaSuccess || cudaMemGetInfo(&fr, &tot) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    const uint64_t per_block = 50ull * ((uint64_t)base_block_size * multiplier + 64); // ~47 B per block byte (data layout)
    const uint64_t fit = (uint64_t)(fr / 3) / per_block;
    const uint64_t want = std::max<uint64_t>(64, 600ull * 900000 / ((uint64_t)base_block_size * multiplier)); // ~540 MB of input per batch pair
    return (int)std::max<uint64_t>(1, std::min<uint64_t>(std::min<uint64_t>(fit, want), 4096));
}

extern "C" int bz2b200_compress_stream(bz2b200_t *h, const uint8_t *input, size_t n_bytes, uint8_t *out, size_t out_capacity, size_t *out_bytes)
{
    if (!h || !out || !out_bytes) return BZ2B200_E_ARG;
    if (out_capacity < 14) { h->err = "output buffer too small"; return BZ2B200_E_OUTPUT; }
    if (is_device_ptr(out)) { h->err = "compress_stream needs a host output buffer"; return BZ2B200_E_ARG; }
    // stream header (OutputStream.hpp:126-128)
    out[0] = 'B'; out[1] = 'Z'; out[2] = 'h'; out[3] = (uint8_t)('0' + h->mult);
    uint64_t nbits = 0, nblocks = 0;
    size_t consumed = 0;
    std::vector<uint32_t> crcs(n_bytes / (size_t)std::max(1, (h->bs - 5) * 4 / 5) + 4);
    int rc = bz2b200_compress_blocks(h, input, n_bytes, 1, 0, out + 4, out_capacity - 4, &nbits, &consumed, crcs.data(), nullptr, crcs.size(), &nblocks);
    if (rc) return rc;
    // trailer (OutputStream.hpp:163-176): end magic, combined CRC (fold :202), zero padding
    uint32_t scrc = 0;
    for (uint64_t i = 0; i < nblocks; ++i) scrc = ((scrc << 1) | (scrc >> 31)) ^ crcs[i];
    uint64_t bitpos = 32 + nbits;
    if ((bitpos + 80 + 7) / 8 > out_capacity) { h->err = "output buffer too small"; return BZ2B200_E_OUTPUT; }
    if ((bitpos & 7) == 0) out[bitpos >> 3] = 0;
    auto put = [&](int count, uint64_t value) {
        for (int i = count - 1; i >= 0; --i) {
            if ((bitpos & 7) == 0) out[bitpos >> 3] = 0;
            if ((value >> i) & 1) out[bitpos >> 3] |= (uint8_t)(0x80u >> (bitpos & 7));
            ++bitpos;
        }
    };
    put(24, 0x177245); put(24, 0x385090); put(32, scrc);
    *out_bytes = (size_t)((bitpos + 7) / 8);
    return BZ2B200_OK;
}

// ---------------------------------------------------------------------------------------------------------------- decompress
static const char *dec_err_text(uint32_t e)
{
    switch (e) { // the reference's exception messages (BlockDecompressor.hpp, HuffmanStageDecoder.hpp, BitInputStream.hpp)
    case DEC_E_TABLES: return "block Huffman tables invalid";
    case DEC_E_BLOCKSIZE: return "BZip2 block exceeds declared block size";
    case DEC_E_DECODE: return "Error decoding  block";
    case DEC_E_STARTPTR: return "BZip2 start pointer invalid";
    case DEC_E_DATA: return "Insufficient data";
    case DEC_E_RANDOMISED: return "BZip2 randomised blocks not implemented";
    default: return "BZip2 stream format error";
    }
}

// multi: after a complete stream keep decoding while another "BZh1".."BZh9" header follows at the next byte (concatenated .bz2
// files, what bzip2 -d and libbzip2 front ends do); anything else behind a complete stream is ignored, as the
// reference's InputStream ignores everything behind its first stream (InputStream.hpp:146-153).
static int decompress_impl(bz2b200_t *h, const uint8_t *input, size_t n_in, uint8_t *out, size_t out_capacity, size_t *out_bytes, bool multi, uint64_t *n_streams)
{
    cudaStream_t st = h->stream;
    h->stats = bz2b200_stats_t{};
    *out_bytes = 0;
    if (!h->batch_ready) { // a handle whose compress batches run on lanes gets its own per-batch buffers at the first decompress call
        int rc = alloc_batch(h);
        if (rc) return rc;
    }
    auto fail = [&](const char *msg) { h->err = msg; return BZ2B200_E_FORMAT; };
    // ---- stream header (InputStream::initializeStream, InputStream.hpp:104-128)
    if (n_in < 4) return fail(n_in == 0 ? "Insufficient data" : "Invalid BZip2 header");
    uint8_t hdr[4];
    CK(cudaMemcpy(hdr, input, 4, cudaMemcpyDefault));
    if (hdr[0] != 'B' || hdr[1] != 'Z' || hdr[2] != 'h' || hdr[3] < '1' || hdr[3] > '9') return fail("Invalid BZip2 header");
    uint32_t stream_bs = (uint32_t)(hdr[3] - '0') * (uint32_t)h->base;
    if (stream_bs > (uint32_t)h->bs) { h->err = "stream block size exceeds the handle's block size"; return BZ2B200_E_BLOCKSIZE; }
    const uint32_t max_selectors = (uint32_t)h->base * 9 / 50 + 1; // HUFFMAN_MAXIMUM_SELECTORS, Config.hpp:41
    // ---- input to the device, zero padded (the bit reader fetches whole 32-bit windows)
    ENSURE(h->in, n_in + 64);
    CK(cudaMemsetAsync(ptr<uint8_t>(h->in) + n_in, 0, 64, st));
    CK(cudaEventRecord(h->ev[0], st));
    CK(cudaMemcpyAsync(h->in.p, input, n_in, cudaMemcpyDefault, st));
    CK(cudaEventRecord(h->ev[1], st));
    const uint8_t *d_in = ptr<uint8_t>(h->in);
    const uint64_t n_bits = (uint64_t)n_in * 8;
    const bool out_dev = out && is_device_ptr(out);
    uint8_t *d_out = out;
    if (!out_dev) { ENSURE(h->out, out_capacity + 64); d_out = ptr<uint8_t>(h->out); }
    // ---- block candidates: both magics at every bit position
    const uint32_t cap_c = (uint32_t)std::min<uint64_t>(n_in / 16 + 1024, 1u << 28);
    const uint32_t G = h->G;
    ENSURE(h->dec_cand, (size_t)cap_c * 8);
    ENSURE(h->dec_info, (size_t)G * sizeof(DecInfo));
    ENSURE(h->dec_list0, (size_t)G * 256);
    ENSURE(h->dec_len, (size_t)G * 8);
    ENSURE(h->dec_off, (size_t)G * 8);
    ENSURE(h->dec_starts, (size_t)(G + 1) * 8);
    ENSURE(h->dec_crc, (size_t)G * 4);
    uint32_t *d_count = reinterpret_cast<uint32_t *>(ptr<uint64_t>(h->scalars) + Scalars::TICKET) + 1;
    CK(cudaEventRecord(h->ev[10], st));
    CK(cudaMemsetAsync(d_count, 0, 4, st));
    KL(KID_k_dec_scan, k_dec_scan<<<(unsigned)((n_in + 255) / 256), 256, 0, st>>>(d_in, n_in, ptr<uint64_t>(h->dec_cand), cap_c, d_count));
    h->stats.kernel_launches += 1;
    uint32_t n_cand = 0;
    CK(cudaMemcpyAsync(&n_cand, d_count, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    h->stats.ms_h2d = ev_ms(h->ev[0], h->ev[1]);
    if (n_cand > cap_c) return fail("BZip2 stream format error");
    std::vector<uint64_t> cand(n_cand);
    if (n_cand) CK(cudaMemcpy(cand.data(), h->dec_cand.p, (size_t)n_cand * 8, cudaMemcpyDeviceToHost));
    std::sort(cand.begin(), cand.end());
    if (n_cand) CK(cudaMemcpy(h->dec_cand.p, cand.data(), (size_t)n_cand * 8, cudaMemcpyHostToDevice));
    // ---- batches of candidates; the chain walk (InputStream::initializeNextBlock, :130-158) accepts a candidate only where the
    // previous block ended, so a magic that occurs inside coded data is never used
    std::vector<DecInfo> info(G);
    std::vector<uint64_t> len(G), off(G), starts(G + 1);
    std::vector<uint32_t> crc(G);
    uint64_t expected = 32, total_out = 0;
    uint32_t stream_crc = 0;
    bool ended = false, sizing = false;
    size_t ci = 0;
    uint64_t nblocks = 0, nstreams = 0;
    for (;;) { // one iteration per stream
    ended = false;
    stream_crc = 0;
    while (!ended) {
        // next batch: block candidates at or behind `expected`; stop at an end marker sitting exactly there
        while (ci < cand.size() && (cand[ci] >> 1) < expected) ++ci;
        if (ci >= cand.size()) return fail((expected + 48 > n_bits) ? "Insufficient data" : "BZip2 stream format error");
        if ((cand[ci] >> 1) != expected) return fail("BZip2 stream format error");
        if (cand[ci] & 1) { ended = true; break; }
        const size_t c0 = ci;
        uint32_t nb = 0;
        while (ci < cand.size() && nb < G && !(cand[ci] & 1)) { ++ci; ++nb; } // block candidates only; an end marker closes the batch
        // (sorted candidates: block candidates between c0 and ci are contiguous in the list, false positives included)
        CK(cudaMemsetAsync(h->ranks.p, 0, (size_t)nb * h->S, st)); // rank 0 everywhere: runs are not stored
        KL(KID_k_dec_block, k_dec_block<<<(nb + DB_WARPS - 1) / DB_WARPS, DB_WARPS * 32, DB_WARPS * sizeof(DecTables), st>>>(
                                d_in, n_bits, ptr<uint64_t>(h->dec_cand), (uint32_t)c0, nb, stream_bs, ptr<uint8_t>(h->ranks), h->S, ptr<uint8_t>(h->sel), h->sel_stride,
                                max_selectors, (uint32_t)std::min<uint64_t>(h->sel_stride, 0xffffffffull), ptr<uint64_t>(h->dec_list0), ptr<DecInfo>(h->dec_info)));
        uint32_t *d_counts = ptr<uint32_t>(h->cnts);
        KL(KID_k_dec_bases, k_dec_counts<<<(nb + 255) / 256, 256, 0, st>>>(ptr<DecInfo>(h->dec_info), nb, d_counts));
        CK(cudaMemcpyAsync(info.data(), h->dec_info.p, (size_t)nb * sizeof(DecInfo), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (h->profiling) prof_collect(h);
        uint32_t max_n = 0;
        for (uint32_t i = 0; i < nb; ++i) max_n = std::max(max_n, info[i].n);
        if (max_n) {
            // inverse BWT: stable counting sort of L by byte value (one single-sweep radix pass), then the pointer chase
            uint32_t *ticket = reinterpret_cast<uint32_t *>(ptr<uint64_t>(h->scalars) + Scalars::TICKET);
            const uint32_t tiles = (max_n + RS_TILE - 1) / RS_TILE;
            // inverse move-to-front, tile-parallel: ranks -> L
            {
                dim3 gi((max_n + IM_WARPS * IM_TILE - 1) / (IM_WARPS * IM_TILE), nb);
                KL(KID_k_imtf_tiles, k_imtf_tiles<false><<<gi, IM_WARPS * 32, 0, st>>>(ptr<uint8_t>(h->ranks), d_counts, h->S, ptr<uint64_t>(h->lists), h->tilesM, ptr<uint8_t>(h->L)));
                KL(KID_k_imtf_compose, k_imtf_compose<<<(nb + IM_WARPS - 1) / IM_WARPS, IM_WARPS * 32, 0, st>>>(d_counts, ptr<uint64_t>(h->lists), h->tilesM,
                                                                                                          ptr<uint64_t>(h->dec_list0), nb));
                KL(KID_k_imtf_tiles, k_imtf_tiles<true><<<gi, IM_WARPS * 32, 0, st>>>(ptr<uint8_t>(h->ranks), d_counts, h->S, ptr<uint64_t>(h->lists), h->tilesM, ptr<uint8_t>(h->L)));
                h->stats.kernel_launches += 3;
            }
            const uint32_t nc = (max_n + RC_TILES * RS_TILE - 1) / (RC_TILES * RS_TILE);
            KL(KID_k_bwt_bytehist, k_bwt_bytehist<<<dim3(nc, nb), 256, 0, st>>>(ptr<uint8_t>(h->L), d_counts, h->S, ptr<uint32_t>(h->rc_partial), h->ctasC));
            KL(KID_k_radix_bases, k_radix_bases<<<nb, 256, 0, st>>>(ptr<uint32_t>(h->rc_partial), d_counts, 1, h->ctasC, nb, ptr<uint32_t>(h->digit_base), 1, nullptr));
            CK(cudaMemsetAsync(h->hist.p, 0, (size_t)nb * h->tilesR * 256 * 4, st));
            CK(cudaMemsetAsync(ticket, 0, 4, st));
            KL(KID_k_radix_onesweep, k_radix_onesweep<true><<<dim3(tiles, nb), RS_THREADS, sizeof(ScatterSmem), st>>>(
                                         nullptr, ptr<uint64_t>(h->kA), d_counts, h->S, VAL_BITS, ptr<uint32_t>(h->digit_base), ptr<uint32_t>(h->hist), h->tilesR,
                                         ticket, ptr<uint8_t>(h->L), 1, nullptr));
            dim3 gd((max_n + 255) / 256, nb);
            KL(KID_k_dec_double, k_dec_double<<<gd, 256, 0, st>>>(ptr<uint64_t>(h->kA), ptr<uint64_t>(h->kB), d_counts, h->S, 1));
            KL(KID_k_dec_double, k_dec_double<<<gd, 256, 0, st>>>(ptr<uint64_t>(h->kB), ptr<uint64_t>(h->kA), d_counts, h->S, 2));
            KL(KID_k_dec_double, k_dec_double8<<<gd, 256, 0, st>>>(ptr<uint64_t>(h->kA), ptr<uint64_t>(h->loc), ptr<uint32_t>(h->rank), d_counts, h->S));
            KL(KID_k_dec_double, k_dec_double16<<<gd, 256, 0, st>>>(ptr<uint64_t>(h->loc), ptr<uint32_t>(h->rank), ptr<uint64_t>(h->kB), ptr<uint32_t>(h->kA), d_counts, h->S));
            KL(KID_k_dec_chase, k_dec_chase<<<(nb + 3) / 4, 128, 0, st>>>(ptr<uint64_t>(h->loc), ptr<uint64_t>(h->kB), ptr<uint32_t>(h->kA), h->S, ptr<DecInfo>(h->dec_info), nb,
                                                                      ptr<uint8_t>(h->rle), ptr<uint64_t>(h->dec_len)));
            h->stats.kernel_launches += 6;
        } else {
            CK(cudaMemsetAsync(h->dec_len.p, 0, (size_t)nb * 8, st));
        }
        h->stats.kernel_launches += 2;
        CK(cudaMemcpyAsync(len.data(), h->dec_len.p, (size_t)nb * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (h->profiling) prof_collect(h);
        // chain walk over the batch
        uint32_t n_acc = 0;
        uint64_t written_end = total_out;
        const uint64_t batch_out0 = total_out;
        bool stop_batch = false;
        for (uint32_t i = 0; i < nb; ++i) {
            off[i] = ~0ull;
            if (stop_batch) continue;
            const uint64_t pos = cand[c0 + i] >> 1;
            if (pos < expected) continue;                 // a magic inside the previous block's bits
            if (pos > expected) { stop_batch = true; continue; } // resolved by the outer loop (end marker or format error)
            if (info[i].err != DEC_OK) return fail(dec_err_text(info[i].err));
            // the output does not fit: go on WITHOUT writing (lengths only), so that the caller learns the size the stream needs and
            // retries once instead of doubling its buffer and decoding again and again
            if (total_out + len[i] > out_capacity) sizing = true;
            if (!sizing) {
                off[i] = total_out;
                starts[n_acc] = total_out;
                crc[n_acc] = info[i].crc;
                ++n_acc;
                written_end = total_out + len[i];
            }
            total_out += len[i];
            expected = info[i].end_bit;
            ++nblocks;
        }
        if (n_acc == 0 && !sizing) return fail("BZip2 stream format error");
        starts[n_acc] = written_end;
        CK(cudaMemcpyAsync(h->dec_off.p, off.data(), (size_t)nb * 8, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(h->dec_starts.p, starts.data(), (size_t)(n_acc + 1) * 8, cudaMemcpyHostToDevice, st));
        KL(KID_k_dec_expand, k_dec_expand<<<(nb + DX_WARPS - 1) / DX_WARPS, DX_WARPS * 32, 0, st>>>(ptr<uint8_t>(h->rle), h->S, ptr<DecInfo>(h->dec_info), nb, ptr<uint64_t>(h->dec_off), d_out));
        // block CRCs over the decoded bytes (BlockDecompressor::checkCRC, :99-106)
        uint64_t max_len = 0;
        for (uint32_t i = 0; i < n_acc; ++i) max_len = std::max(max_len, starts[i + 1] - starts[i]);
        if (n_acc) CK(cudaMemsetAsync(h->dec_crc.p, 0, (size_t)n_acc * 4, st));
        if (max_len) {
            dim3 cgrid((unsigned)((max_len + CRC_CHUNK - 1) / CRC_CHUNK), n_acc);
            KL(KID_k_crc_chunks, k_crc_chunks<<<cgrid, CRC_THREADS, 0, st>>>(d_out, ptr<uint64_t>(h->dec_starts), 0u, ptr<uint32_t>(h->dec_crc)));
        }
        if (n_acc) KL(KID_k_crc_finish, k_crc_finish<<<(n_acc + 127) / 128, 128, 0, st>>>(ptr<uint64_t>(h->dec_starts), 0u, n_acc, ptr<uint32_t>(h->dec_crc)));
        h->stats.kernel_launches += 3;
        std::vector<uint32_t> got(n_acc);
        if (n_acc) CK(cudaMemcpyAsync(got.data(), h->dec_crc.p, (size_t)n_acc * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (h->profiling) prof_collect(h);
        for (uint32_t i = 0; i < n_acc; ++i) {
            if (got[i] != crc[i]) return fail("BZip2 block CRC error");
            stream_crc = ((stream_crc << 1) | (stream_crc >> 31)) ^ crc[i]; // InputStream.hpp:136-137
        }
        (void)batch_out0;
        // candidates of this batch that were not reached stay for the next one
        size_t back = c0;
        while (back < ci && (cand[back] >> 1) < expected) ++back;
        ci = back;
    }
    // ---- end of stream: stored combined CRC (InputStream.hpp:146-153)
    if (expected + 48 + 32 > n_bits) return fail("Insufficient data");
    uint8_t tail[8] = {0};
    const uint64_t tb = (expected + 48) >> 3;
    CK(cudaMemcpy(tail, d_in + tb, std::min<uint64_t>(8, n_in + 64 - tb), cudaMemcpyDeviceToHost));
    uint64_t w = 0;
    for (int k = 0; k < 8; ++k) w = (w << 8) | tail[k];
    const uint32_t stored = (uint32_t)((w << ((expected + 48) & 7)) >> 32);
    if (stored != stream_crc && !sizing) return fail("BZip2 stream CRC error");
    ++nstreams;
    if (!multi) break;
    const uint64_t nxt = (expected + 48 + 32 + 7) >> 3; // the padding ends the stream on a byte boundary (OutputStream.hpp:172)
    if (nxt + 4 > n_in) break;
    CK(cudaMemcpy(hdr, d_in + nxt, 4, cudaMemcpyDeviceToHost));
    if (hdr[0] != 'B' || hdr[1] != 'Z' || hdr[2] != 'h' || hdr[3] < '1' || hdr[3] > '9') break; // trailing bytes that are no stream
    stream_bs = (uint32_t)(hdr[3] - '0') * (uint32_t)h->base;
    if (stream_bs > (uint32_t)h->bs) { h->err = "stream block size exceeds the handle's block size"; return BZ2B200_E_BLOCKSIZE; }
    expected = nxt * 8 + 32;
    } // streams
    if (n_streams) *n_streams = nstreams;
    if (sizing) { // every block has been decoded far enough to know its length: this is what the caller's buffer must hold
        CK(cudaStreamSynchronize(st));
        *out_bytes = (size_t)total_out;
        h->err = "output buffer too small";
        return BZ2B200_E_OUTPUT;
    }
    CK(cudaEventRecord(h->ev[11], st));
    if (!out_dev && total_out) {
        CK(cudaEventRecord(h->ev[12], st));
        CK(cudaMemcpyAsync(out, d_out, total_out, cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(h->ev[13], st));
    }
    CK(cudaStreamSynchronize(st));
    h->stats.ms_total = ev_ms(h->ev[10], h->ev[11]);
    if (!out_dev && total_out) h->stats.ms_d2h = ev_ms(h->ev[12], h->ev[13]);
    h->stats.n_input = n_in;
    h->stats.n_blocks = nblocks;
    *out_bytes = (size_t)total_out;
    return BZ2B200_OK;
}

extern "C" int bz2b200_decompress_stream(bz2b200_t *h, const uint8_t *input, size_t n_bytes, uint8_t *out, size_t out_capacity, size_t *out_bytes)
{
    if (!h) return BZ2B200_E_ARG;
    if (!h->children.empty()) { // the decompress side runs on the first device
        int rc = bz2b200_decompress_stream(h->children[0], input, n_bytes, out, out_capacity, out_bytes);
        h->stats = h->children[0]->stats; h->err = h->children[0]->err;
        return rc;
    }
    if ((!input && n_bytes) || !out_bytes || (!out && out_capacity)) { h->err = "bad argument"; return BZ2B200_E_ARG; }
    CK(cudaSetDevice(h->device));
    h->err.clear();
    int rc = decompress_impl(h, input, n_bytes, out, out_capacity, out_bytes, false, nullptr);
    cudaStreamSynchronize(h->stream);
    if (rc != BZ2B200_OK) (void)cudaGetLastError();
    return rc;
}

extern "C" int bz2b200_decompress_concat(bz2b200_t *h, const uint8_t *input, size_t n_bytes, uint8_t *out, size_t out_capacity, size_t *out_bytes,
                                         uint64_t *n_streams)
{
    if (!h) return BZ2B200_E_ARG;
    if (!h->children.empty()) {
        int rc = bz2b200_decompress_concat(h->children[0], input, n_bytes, out, out_capacity, out_bytes, n_streams);
        h->stats = h->children[0]->stats; h->err = h->children[0]->err;
        return rc;
    }
    if ((!input && n_bytes) || !out_bytes || (!out && out_capacity)) { h->err = "bad argument"; return BZ2B200_E_ARG; }
    CK(cudaSetDevice(h->device));
    h->err.clear();
    int rc = decompress_impl(h, input, n_bytes, out, out_capacity, out_bytes, true, n_streams);
    cudaStreamSynchronize(h->stream);
    if (rc != BZ2B200_OK) (void)cudaGetLastError();
    return rc;
}

extern "C" int bz2b200_stage_dump(bz2b200_t *h, uint64_t block_index, int stage, void *dst, size_t cap, size_t *len_bytes)
{
    if (!h || !dst || !len_bytes) return BZ2B200_E_ARG;
    if (!h->children.empty()) return BZ2B200_E_STATE; // per-device batches: dump through a single-device handle
    if (block_index < h->last_b0 || block_index >= h->last_b0 + h->last_nb) return BZ2B200_E_STATE;
    CK(cudaSetDevice(h->device));
    const uint64_t bl = block_index - h->last_b0, S = h->S;
    uint32_t n = 0;
    CK(cudaMemcpy(&n, ptr<uint32_t>(h->rle_len) + block_index, 4, cudaMemcpyDeviceToHost));
    if (!h->lanes.empty()) h = h->lanes[h->last_lane]; // the batch's buffers belong to the lane that ran it
    HuffBlock hb;
    CK(cudaMemcpy(&hb, ptr<HuffBlock>(h->hb) + bl, sizeof(hb), cudaMemcpyDeviceToHost));
    const void *src = nullptr;
    size_t len = 0;
    switch (stage) {
    case BZ2B200_STAGE_RLE1: src = ptr<uint8_t>(h->rle) + bl * S; len = n; break;
    case BZ2B200_STAGE_BWT: src = ptr<uint8_t>(h->L) + bl * S; len = n; break;
    case BZ2B200_STAGE_ORIGPTR: src = ptr<uint32_t>(h->orig) + bl; len = 4; break;
    case BZ2B200_STAGE_MTF: src = ptr<uint16_t>(h->mtf) + bl * S; len = (size_t)hb.m * 2; break;
    case BZ2B200_STAGE_FREQ: src = ptr<uint32_t>(h->freq) + bl * 258; len = 258 * 4; break;
    case BZ2B200_STAGE_SELECTORS: src = ptr<uint8_t>(h->sel) + bl * h->sel_stride; len = hb.nsel; break;
    case BZ2B200_STAGE_RANK: src = ptr<uint32_t>(h->rank) + bl * S; len = (size_t)n * 4; break;
    case BZ2B200_STAGE_PRESENT: {
        uint32_t bm[8];
        CK(cudaMemcpy(bm, ptr<uint32_t>(h->present) + bl * 8, 32, cudaMemcpyDeviceToHost));
        if (cap < 256) return BZ2B200_E_OUTPUT;
        for (int v = 0; v < 256; ++v) ((uint8_t *)dst)[v] = (bm[v >> 5] >> (v & 31)) & 1;
        *len_bytes = 256;
        return BZ2B200_OK;
    }
    case BZ2B200_STAGE_LENGTHS: {
        std::vector<uint8_t> tmp((size_t)HUFF_MAXTABLES * HUFF_LEN_STRIDE);
        CK(cudaMemcpy(tmp.data(), ptr<uint8_t>(h->lens) + bl * HUFF_MAXTABLES * HUFF_LEN_STRIDE, tmp.size(), cudaMemcpyDeviceToHost));
        len = (size_t)hb.ntab * 258;
        if (cap < len) return BZ2B200_E_OUTPUT;
        memset(dst, 0, len);
        for (uint32_t t = 0; t < hb.ntab; ++t) memcpy((uint8_t *)dst + t * 258, tmp.data() + t * HUFF_LEN_STRIDE, hb.alpha);
        *len_bytes = len;
        return BZ2B200_OK;
    }
    default: return BZ2B200_E_ARG;
    }
    if (cap < len) return BZ2B200_E_OUTPUT;
    if (len) CK(cudaMemcpy(dst, src, len, cudaMemcpyDeviceToHost));
    *len_bytes = len;
    return BZ2B200_OK;
}

extern "C" int bz2b200_set_profiling(bz2b200_t *h, int on)
{
    if (!h) return BZ2B200_E_ARG;
    h->profiling = on != 0;
    for (int i = 0; i < KID_COUNT; ++i) { h->k_ms[i] = 0; h->k_launches[i] = 0; h->k_bytes[i] = 0; }
    for (bz2b200 *c : h->children) bz2b200_set_profiling(c, on);
    return BZ2B200_OK;
}

extern "C" int bz2b200_kernel_times(const bz2b200_t *h, int index, const char **name, double *ms, uint64_t *launches, uint64_t *bytes)
{
    if (!h || !name || !ms || !launches || !bytes) return BZ2B200_E_ARG;
    if (index < 0 || index >= KID_COUNT) return BZ2B200_E_STATE;
    *name = kKernelNames[index];
    *ms = h->k_ms[index];
    *launches = h->k_launches[index];
    *bytes = h->k_bytes[index];
    for (const bz2b200 *c : h->children) { *ms += c->k_ms[index]; *launches += c->k_launches[index]; *bytes += c->k_bytes[index]; }
    return BZ2B200_OK;
}

extern "C" int bz2b200_get_stats(const bz2b200_t *h, bz2b200_stats_t *s)
{
    if (!h || !s) return BZ2B200_E_ARG;
    *s = h->stats;
    return BZ2B200_OK;
}
