/* TEST INFRASTRUCTURE (oracle build only) - fake UPMEM runtime, written from scratch.
 * Stands in for the UPMEM SDK's <mram.h> so that /root/reference/src/decoder_dpu.c
 * compiles unmodified for x86.  MRAM is ordinary memory here. */
#ifndef ORACLE_SHIM_MRAM_H
#define ORACLE_SHIM_MRAM_H
#include <string.h>
#include <stdint.h>
#define __mram
#define __mram_noinit
#define __mram_ptr
#define __host
#define __dma_aligned __attribute__((aligned(8)))
static inline void mram_read(const void *from_mram, void *to_wram, unsigned int nbytes) {
    memcpy(to_wram, from_mram, nbytes);
}
static inline void mram_write(const void *from_wram, void *to_mram, unsigned int nbytes) {
    memcpy(to_mram, from_wram, nbytes);
}
#endif
