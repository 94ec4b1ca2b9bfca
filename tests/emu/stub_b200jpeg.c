/* TEST STUB - NOT the product.  A stand-in for libb200jpeg.so that implements only the entry points the <dpu>
 * facade (pim_jpeg_decoder_b200/host/compat/dpu) calls, with bj_exec_mcus answered by the oracle restatement
 * (oracle/restate.c).  It exists so the facade's gather/scatter logic and the reference host's use of it can be
 * checked on a machine without a GPU (tests/test_host_compat.py); it is built into tests/emu/_stub/ and nothing
 * under pim_jpeg_decoder_b200/ ever links it. */
#include <stdlib.h>
#include <string.h>
#include "b200jpeg.h"
#include "restate.h"

struct bj_ctx { int device; };

int bj_create(bj_ctx **ctx, int device) { *ctx = calloc(1, sizeof(bj_ctx)); (*ctx)->device = device; return BJ_OK; }
void bj_destroy(bj_ctx *ctx) { free(ctx); }
const char *bj_status_string(int s) { return s == BJ_OK ? "ok" : "error"; }
const char *bj_last_error(const bj_ctx *ctx) { (void)ctx; return ""; }
void *bj_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void bj_host_free(void *p) { free(p); }
int bj_exec_mcus(bj_ctx *ctx, const uint32_t *metadata, int16_t *mcus, int nchunk) {
    (void)ctx;
    rs_exec_mcus(metadata, mcus, nchunk);
    return BJ_OK;
}
int bj_get_stat(const bj_ctx *ctx, const char *name, double *value) { (void)ctx; (void)name; *value = 0.0; return BJ_OK; }
