/* TEST INFRASTRUCTURE - fake <barrier.h>.  Tasklets are run one after the other (0..N-1) by the
 * glue, tasklet 0 first, so the single barrier in decoder_dpu.c:92 is trivially satisfied. */
#ifndef ORACLE_SHIM_BARRIER_H
#define ORACLE_SHIM_BARRIER_H
typedef struct { int unused; } oracle_fake_barrier_t;
#define BARRIER_INIT(name, count) oracle_fake_barrier_t name = { count }
static inline void barrier_wait(oracle_fake_barrier_t *b) { (void)b; }
#endif
