/* TEST INFRASTRUCTURE - fake <perfcounter.h>: a monotonically increasing fake cycle counter. */
#ifndef ORACLE_SHIM_PERFCOUNTER_H
#define ORACLE_SHIM_PERFCOUNTER_H
#include <stdint.h>
#include <stdbool.h>
#define COUNT_CYCLES 1
extern uint32_t oracle_fake_cycles;
static inline void perfcounter_config(int what, bool reset) { (void)what; if (reset) oracle_fake_cycles = 0; }
static inline uint32_t perfcounter_get(void) { return ++oracle_fake_cycles; }
#endif
