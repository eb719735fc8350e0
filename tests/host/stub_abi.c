/* stub_abi.c -- TEST INFRASTRUCTURE, never shipped: the few entry points of include/bz2b200.h the OutputStream facade calls,
 * answered on the CPU by the oracle (oracle/bz2_oracle.c, compiled into the same test binary).  It exists so that the facade's
 * HOST logic -- batching, the bytes carried from batch to batch, the < 8 carried bits, the worker and writer threads, readFrom() --
 * runs under ThreadSanitizer / AddressSanitizer in the CPU suite (tests/test_host_logic.py); the product library has no such
 * path and fails when there is no CUDA device.
 *
 * compress_blocks semantics restated from include/bz2b200.h: with is_final == 0 only blocks whose refusing byte was seen are
 * produced and *consumed is the offset of the first byte of the unfinished block.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "bz2b200.h"

int orc_compress(const uint8_t *in, size_t n, int base, int mult, uint8_t *out, size_t cap, size_t *out_len);
long orc_rle1_blocks(const uint8_t *in, size_t n, int base, int mult, uint64_t *starts, int32_t *lengths, uint32_t *crcs, long max_blocks);

struct bz2b200 {
    int base, mult, blocks, devices;
    char err[128];
    unsigned long calls;
};

static int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return e && atoi(e) > 0 ? atoi(e) : dflt;
}

int bz2b200_device_count(void) { return env_int("STUB_DEVICES", 1); }
int bz2b200_recommended_blocks(int device, int base, int mult) { (void)device; (void)base; (void)mult; return env_int("STUB_RECOMMENDED", 4); }
void *bz2b200_host_alloc(size_t n) { return malloc(n ? n : 1); }
void bz2b200_host_free(void *p) { free(p); }
const char *bz2b200_strerror(int rc) { return rc == BZ2B200_OK ? "ok" : "stub error"; }
const char *bz2b200_last_error(const bz2b200_t *h) { return h ? h->err : ""; }

static int make(bz2b200_t **out, int base, int mult, int blocks, int devices)
{
    if (mult < 1 || mult > 9) return BZ2B200_E_BLOCKSIZE;
    if (blocks < 1) return BZ2B200_E_PARALLEL;
    bz2b200_t *h = (bz2b200_t *)calloc(1, sizeof(*h));
    if (!h) return BZ2B200_E_NOMEM;
    h->base = base; h->mult = mult; h->blocks = blocks; h->devices = devices;
    *out = h;
    return BZ2B200_OK;
}
int bz2b200_create(bz2b200_t **out, int device, int base, int mult, int blocks) { (void)device; return make(out, base, mult, blocks, 1); }
int bz2b200_create_multi(bz2b200_t **out, const int *ids, int n, int base, int mult, int blocks) { (void)ids; return make(out, base, mult, blocks, n); }
void bz2b200_destroy(bz2b200_t *h)
{
    if (h && getenv("STUB_TRACE")) fprintf(stderr, "stub: %lu compress_blocks calls, %d device(s)\n", h->calls, h->devices);
    free(h);
}
size_t bz2b200_max_output_bytes(const bz2b200_t *h, size_t n) { (void)h; return n + n / 2 + 4096; }

static int bit_at(const uint8_t *p, uint64_t i) { return (p[i >> 3] >> (7 - (i & 7))) & 1; }

int bz2b200_compress_blocks(bz2b200_t *h, const uint8_t *input, size_t n, int is_final, unsigned start_bit, uint8_t *out, size_t out_cap,
                            uint64_t *out_nbits, size_t *consumed, uint32_t *crcs, uint64_t *nbits_b, size_t cap_b, uint64_t *n_blocks)
{
    if (!h || start_bit > 7 || (n && !input)) return BZ2B200_E_ARG;
    h->calls++;
    *out_nbits = 0; *consumed = 0; *n_blocks = 0;
    if ((unsigned long)env_int("STUB_FAIL_CALL", 0) == h->calls) { /* injected failure: the facade must report it and still shut down */
        snprintf(h->err, sizeof h->err, "stub: injected failure in call %lu", h->calls);
        return BZ2B200_E_CUDA;
    }
    if (!n) return BZ2B200_OK;
    long max = (long)(n / (size_t)((h->base * h->mult) * 4 / 5 - 4)) + 4;
    uint64_t *starts = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)max);
    int32_t *lens = (int32_t *)malloc(sizeof(int32_t) * (size_t)max);
    uint32_t *bc = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)max);
    long nb = orc_rle1_blocks(input, n, h->base, h->mult, starts, lens, bc, max);
    long emit = is_final ? nb : (nb > 0 ? nb - 1 : 0);
    size_t used = is_final ? n : (nb > 0 ? (size_t)starts[nb - 1] : 0);
    int rc = BZ2B200_OK;
    if (emit > 0) {
        if ((size_t)emit > cap_b && (crcs || nbits_b)) rc = BZ2B200_E_OUTPUT;
        size_t cap = used + used / 2 + 4096, len = 0;
        uint8_t *tmp = (uint8_t *)malloc(cap);
        if (rc == BZ2B200_OK && orc_compress(input, used, h->base, h->mult, tmp, cap, &len) != 0) rc = BZ2B200_E_OUTPUT;
        if (rc == BZ2B200_OK) {
            /* stream = 32 header bits | blocks | 48-bit end marker | 32-bit CRC | 0..7 zero bits: find the marker from the back */
            uint32_t folded = 0;
            for (long i = 0; i < emit; ++i) folded = ((folded << 1) | (folded >> 31)) ^ bc[i];
            uint64_t end = 0;
            for (unsigned pad = 0; pad < 8 && !end; ++pad) {
                const uint64_t at = 8ull * len - pad - 80;
                uint64_t magic = 0; uint32_t crc = 0;
                for (int i = 0; i < 48; ++i) magic = (magic << 1) | (uint64_t)bit_at(tmp, at + (uint64_t)i);
                for (int i = 0; i < 32; ++i) crc = (crc << 1) | (uint32_t)bit_at(tmp, at + 48 + (uint64_t)i);
                if (magic == 0x177245385090ull && crc == folded) end = at;
            }
            const uint64_t nbits = end - 32;
            if (!end) { snprintf(h->err, sizeof h->err, "stub: end marker not found"); rc = BZ2B200_E_ARG; }
            else if ((start_bit + nbits + 7) / 8 > out_cap) rc = BZ2B200_E_OUTPUT;
            else {
                memset(out, 0, (size_t)((start_bit + nbits + 7) / 8));
                for (uint64_t i = 0; i < nbits; ++i)
                    if (bit_at(tmp, 32 + i)) out[(start_bit + i) >> 3] |= (uint8_t)(0x80u >> ((start_bit + i) & 7));
                *out_nbits = nbits;
                *n_blocks = (uint64_t)emit;
                *consumed = used;
                if (crcs) for (long i = 0; i < emit; ++i) crcs[i] = bc[i];
                if (nbits_b) for (long i = 0; i < emit; ++i) nbits_b[i] = 0; /* (not needed by the facade) */
            }
        }
        free(tmp);
    } else if (is_final) *consumed = n;
    free(starts); free(lens); free(bc);
    return rc;
}
