/* TEST INFRASTRUCTURE - fake <defs.h>: me() returns the tasklet the glue is currently running. */
#ifndef ORACLE_SHIM_DEFS_H
#define ORACLE_SHIM_DEFS_H
extern int oracle_fake_tasklet_id;
static inline int me(void) { return oracle_fake_tasklet_id; }
#endif
