/* TEST INFRASTRUCTURE (oracle build only).
 * Compiles the reference's DPU program (/root/reference/src/decoder_dpu.c) verbatim, from where it lies,
 * as an ordinary C translation unit: its main() is renamed, its MRAM/host symbols are ordinary globals,
 * and the NR_TASKLETS tasklets are run one after the other (valid: tasklet 0 initialises before the only
 * barrier, decoder_dpu.c:85-92, and tasklets own disjoint blocks, :159-163).  Nothing of the reference is
 * copied into this repository; build products go to oracle/_ref/ (git-ignored). */
#include <stdint.h>
#include <string.h>

int oracle_fake_tasklet_id = 0;
uint32_t oracle_fake_cycles = 0;

#define main oracle_dpu_program_main
#include "decoder_dpu.c"
#undef main

int oracle_dpu_mcus_len(void) { return 64 * MAX_MCU_PER_DPU * 3; }

void oracle_dpu_run(const uint32_t *metadata276, short *mcus_inout) {
    memcpy(metadata_buffer, metadata276, sizeof(metadata_buffer));
    memcpy(mcus, mcus_inout, sizeof(mcus));
    accumulated_cycles = 0;
    for (int t = 0; t < NR_TASKLETS; t++) {
        oracle_fake_tasklet_id = t;
        oracle_dpu_program_main();
    }
    memcpy(mcus_inout, mcus, sizeof(mcus));
}

void oracle_dpu_counters(uint32_t out4[4]) {
    out4[0] = initialization; out4[1] = dequantization; out4[2] = inverse_dct; out4[3] = color_space_conversion;
}
