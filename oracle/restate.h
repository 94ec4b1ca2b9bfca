/* TEST INFRASTRUCTURE - CPU restatement ("port") of the reference's decode algorithm.
 *
 * This is the oracle of SURVEY.md section 8c.  It is NOT part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product path (pim_jpeg_decoder_b200/) never links, imports or calls anything in oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against the real reference code
 * (oracle/_ref/libref.so = /root/reference/src compiled verbatim) on every fixture in tests/golden/,
 * stage by stage (post-Huffman buffer, post-exec buffer, BMP bytes), and against the committed SHA-256
 * values in tests/golden/golden.json that the reference itself produced.
 *
 * Every function cites the reference file:line it restates (paths relative to /root/reference/).
 */
#ifndef ORACLE_RESTATE_H
#define ORACLE_RESTATE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint8_t offsets[17];   /* cumulative code counts, jpeg.h:129-134 */
    uint8_t symbols[162];
    uint8_t set;
} rs_huff;

typedef struct {
    uint32_t width, height;
    uint32_t ncomp;
    uint32_t hs, vs;                 /* luma sampling = MCU size in 8x8 positions (jpeg_scanner.cpp:250-264) */
    uint32_t mcu_w, mcu_h;           /* ceil(w/8), ceil(h/8)            (jpeg_scanner.cpp:210-211) */
    uint32_t mcu_w_real, mcu_h_real; /* padded to the MCU size          (jpeg_scanner.cpp:257-262) */
    uint32_t restart_interval;
    uint8_t qt_id[3], dc_id[3], ac_id[3], comp_h[3], comp_v[3];
    uint16_t qt_zz[4][64];           /* as stored in the file: zig-zag order */
    uint8_t qt_set[4];
    rs_huff dc[4], ac[4];
    size_t scan_off;                 /* first entropy-coded byte in the file */
    size_t scan_len;                 /* bytes up to (not including) the EOI marker */
    int frame_type;                  /* 0xC0 baseline, 0xC2 progressive */
    int valid;
} rs_header;

enum { RS_RESTART_CORRECT = 0, RS_RESTART_REFQUIRK = 1 };

/* read_JPEG marker walk, jpeg_scanner.cpp:345-403 (+ the segment readers :6-343). 0 = ok. */
int rs_parse(const uint8_t *file, size_t len, rs_header *h);

/* scan-byte loop, jpeg_scanner.cpp:405-433: un-stuff FF00, drop RSTn and fill bytes.  seg_starts (optional)
 * receives the un-stuffed byte offset that follows each RST marker.  Returns un-stuffed length or -1. */
long rs_unstuff(const uint8_t *scan, size_t len, uint8_t *out, uint32_t *seg_starts, int max_seg, int *nseg);

uint32_t rs_num_mcus(const rs_header *h);
uint32_t rs_blocks_per_mcu(const rs_header *h);
uint32_t rs_num_chunks(const rs_header *h, int M);   /* decoder_host.cpp:125-128 */

/* Baseline entropy decode, jpeg_scanner.cpp:438-520,707-756.  Output: coefficients in ZIG-ZAG order,
 * one 64-entry data unit after the other in decode (MCU-interleaved) order; DC already un-differenced.
 * Returns 0 if the whole scan decoded, 1 if the reference would have stopped early (rest left zero). */
int rs_huffman_zz(const rs_header *h, const uint8_t *file, int16_t *coef_zz, int restart_mode);

/* Scatter into the reference's MCU_buffer layout [chunk][blk][comp][pos][64] through the reference's
 * zigzag_map including its idx-48 quirk (common.h:9-18, jpeg_scanner.cpp:517,733-741). */
void rs_coef_to_ref_mcus(const rs_header *h, const int16_t *coef_zz, int16_t *mcus, int M);

/* metadata record of decoder_host.cpp:156-178 (quantisation tables through the quirky map, :306,311). */
void rs_metadata(const rs_header *h, uint32_t md[276], int M);

/* The DPU program on the reference layout, in place: decoder_dpu.c:82-390. */
void rs_exec_mcus(const uint32_t *metadata, int16_t *mcus, int nchunk);

/* write_BMP, bmp_writer.cpp:19-67, into memory.  Returns bytes written (= rs_bmp_size). */
size_t rs_bmp_size(uint32_t width, uint32_t height);
size_t rs_mcus_to_bmp(const uint32_t md[276], const int16_t *mcus, uint8_t *bmp);

/* Convenience: whole pipeline on in-memory file bytes.  rgb (optional) = top-down packed RGB8,
 * bmp (optional) = the exact BMP file bytes. Returns 0 ok, <0 invalid file. */
int rs_decode(const uint8_t *file, size_t len, int restart_mode, uint8_t *rgb, uint8_t *bmp);

/* 1-D IDCT pass and colour maths exposed for unit tests (decoder_dpu.c:219-267, :376-382). */
void rs_idct8(const int32_t in[8], int32_t out[8]);
void rs_ycc_to_rgb(int y, int cb, int cr, int *r, int *g, int *b);

#ifdef __cplusplus
}
#endif
#endif
