// Plain C99 caller of include/bz2b200.h: the boundary is a C ABI (no C++ types, no torch types); run by tests/test_host_logic.py without a GPU.
#include "bz2b200.h"
static void cb(uint64_t next, void *u) { (void)next; (void)u; }
int main(void) {
    bz2b200_t *h = 0; int ids[2] = {0, 0}; uint64_t nb = 0, nbits = 0; size_t used = 0;
    int rc = bz2b200_create_multi(&h, ids, 2, 100000, 9, 4);
    if (rc == BZ2B200_OK) {
        rc = bz2b200_compress_blocks_from(h, 0, 0, 1, 0, 0, 0, 0, &nbits, &used, 0, 0, 0, &nb, cb, 0);
        bz2b200_destroy(h);
    }
    return bz2b200_device_count_of(0) + (bz2b200_host_alloc(0) != 0);
}
