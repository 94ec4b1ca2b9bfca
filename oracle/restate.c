/* TEST INFRASTRUCTURE - CPU restatement of the reference decode path.  See restate.h for the rules on who
 * may use this file.  Written from the behaviour of /root/reference/src (cited per function); no code copied.
 * All arithmetic is integer; tolerance against the reference is 0. */
#include "restate.h"
#include <stdlib.h>
#include <string.h>

/* The reference's zig-zag table INCLUDING its non-standard entry: index 48 maps to 38 (a standard table has 58
 * there), so natural position 58 is never written and 38 receives index 48 and then index 52
 * (headers/common.h:9-18). */
static const uint8_t QMAP[64] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
    30, 37, 44, 51, 38, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

/* ------------------------------------------------------------------ container parse */

typedef struct {
    const uint8_t *p;
    size_t len, pos;
    int eof;
} rd_t;

/* std::ifstream::get() stored into a `byte`: EOF reads as 0xFF and sets the fail state. */
static unsigned rd_get(rd_t *r) {
    if (r->pos >= r->len) { r->eof = 1; return 0xFF; }
    return r->p[r->pos++];
}

/* jpeg_scanner.cpp:187-285 */
static void parse_sof(rd_t *r, rs_header *h, uint8_t used_in_frame[3], int *zero_based) {
    if (h->ncomp != 0) { h->valid = 0; return; }
    unsigned length = (rd_get(r) << 8) + rd_get(r);
    if (rd_get(r) != 8) { h->valid = 0; return; }
    h->height = (rd_get(r) << 8) + rd_get(r);
    h->width = (rd_get(r) << 8) + rd_get(r);
    if (h->height == 0 || h->width == 0) { h->valid = 0; return; }
    h->mcu_h = (h->height + 7) / 8;
    h->mcu_w = (h->width + 7) / 8;
    h->mcu_h_real = h->mcu_h;
    h->mcu_w_real = h->mcu_w;
    unsigned nc = rd_get(r);
    if (nc == 4 || nc == 0) { h->valid = 0; return; }
    h->ncomp = nc;
    for (unsigned i = 0; i < nc; i++) {
        unsigned id = rd_get(r);
        if (id == 0 && i == 0) *zero_based = 1;
        if (*zero_based) id = (id + 1) & 0xFF;
        if (id == 4 || id == 5 || id == 0 || id > nc) { h->valid = 0; return; }
        if (used_in_frame[id - 1]) { h->valid = 0; return; }
        used_in_frame[id - 1] = 1;
        unsigned sf = rd_get(r);
        unsigned ch = sf >> 4, cv = sf & 15;
        h->comp_h[id - 1] = (uint8_t)ch;
        h->comp_v[id - 1] = (uint8_t)cv;
        if (id == 1) {
            if ((ch != 1 && ch != 2) || (cv != 1 && cv != 2)) { h->valid = 0; return; }
            if (ch == 2 && h->mcu_w % 2 == 1) h->mcu_w_real += 1;
            if (cv == 2 && h->mcu_h % 2 == 1) h->mcu_h_real += 1;
            h->hs = ch;
            h->vs = cv;
        } else if (ch != 1 || cv != 1) { h->valid = 0; return; }
        unsigned q = rd_get(r);
        h->qt_id[id - 1] = (uint8_t)q;
        if (q > 3) { h->valid = 0; return; }
    }
    if (length - 8 - 3 * nc != 0) h->valid = 0;
}

/* jpeg_scanner.cpp:287-321 (values kept in file order; the quirky map is applied by the consumers) */
static void parse_dqt(rd_t *r, rs_header *h) {
    int length = (int)((rd_get(r) << 8) + rd_get(r)) - 2;
    while (length > 0) {
        unsigned info = rd_get(r);
        length -= 1;
        unsigned id = info & 15;
        if (id > 3) { h->valid = 0; return; }
        h->qt_set[id] = 1;
        if (info >> 4) {
            for (int i = 0; i < 64; i++) { unsigned hi = rd_get(r); h->qt_zz[id][i] = (uint16_t)((hi << 8) + rd_get(r)); }
            length -= 128;
        } else {
            for (int i = 0; i < 64; i++) h->qt_zz[id][i] = (uint16_t)rd_get(r);
            length -= 64;
        }
        if (r->eof) { h->valid = 0; return; }
    }
    if (length != 0) h->valid = 0;
}

/* jpeg_scanner.cpp:140-185 */
static void parse_dht(rd_t *r, rs_header *h) {
    int length = (int)((rd_get(r) << 8) + rd_get(r)) - 2;
    while (length > 0) {
        unsigned info = rd_get(r);
        unsigned id = info & 15;
        if (id > 3) { h->valid = 0; return; }
        rs_huff *t = (info >> 4) ? &h->ac[id] : &h->dc[id];
        t->set = 1;
        t->offsets[0] = 0;
        unsigned all = 0;
        for (int i = 1; i <= 16; i++) { all += rd_get(r); t->offsets[i] = (uint8_t)all; }
        if (all > 162) { h->valid = 0; return; }
        for (unsigned i = 0; i < all; i++) t->symbols[i] = (uint8_t)rd_get(r);
        length -= 17 + (int)all;
        if (r->eof) { h->valid = 0; return; }
    }
    if (length != 0) h->valid = 0;
}

/* jpeg_scanner.cpp:6-138 */
static void parse_sos(rd_t *r, rs_header *h, const uint8_t used_in_frame[3], int zero_based) {
    if (h->ncomp == 0) { h->valid = 0; return; }
    unsigned length = (rd_get(r) << 8) + rd_get(r);
    uint8_t in_scan[3] = {0, 0, 0};
    unsigned ns = rd_get(r);
    if (ns == 0) { h->valid = 0; return; }
    for (unsigned i = 0; i < ns; i++) {
        unsigned id = rd_get(r);
        if (zero_based) id = (id + 1) & 0xFF;
        if (id == 0 || id > h->ncomp) { h->valid = 0; return; }
        if (!used_in_frame[id - 1] || in_scan[id - 1]) { h->valid = 0; return; }
        in_scan[id - 1] = 1;
        unsigned t = rd_get(r);
        h->dc_id[id - 1] = (uint8_t)(t >> 4);
        h->ac_id[id - 1] = (uint8_t)(t & 15);
        if ((t >> 4) > 3 || (t & 15) > 3) { h->valid = 0; return; }
    }
    unsigned ss = rd_get(r), se = rd_get(r), a = rd_get(r);
    if (h->frame_type == 0xC0) {
        if (ss != 0 || se != 63 || a != 0) { h->valid = 0; return; }
    } else if (h->frame_type == 0xC2) {                  /* :79-106 (a progressive file is never decoded, but its scan header is checked) */
        unsigned ah = a >> 4, al = a & 15;
        if (ss > se || se > 63 || (ss == 0 && se != 0) || (ss != 0 && ns != 1) || (ah != 0 && al != ah - 1)) { h->valid = 0; return; }
    }
    for (unsigned i = 0; i < h->ncomp; i++) {
        if (!in_scan[i]) continue;
        if (!h->qt_set[h->qt_id[i]]) { h->valid = 0; return; }
        if (ss == 0 && !h->dc[h->dc_id[i]].set) { h->valid = 0; return; }
        if (se > 0 && !h->ac[h->ac_id[i]].set) { h->valid = 0; return; }
    }
    if (length - 6 - 2 * ns != 0) h->valid = 0;
}

static void skip_segment(rd_t *r) {
    unsigned length = (rd_get(r) << 8) + rd_get(r);
    for (unsigned i = 0; i + 2 < length + 0u && !r->eof; i++) rd_get(r);
}

int rs_parse(const uint8_t *file, size_t len, rs_header *h) {
    memset(h, 0, sizeof(*h));
    h->valid = 1;
    h->hs = h->vs = 1;
    for (int i = 0; i < 3; i++) h->comp_h[i] = h->comp_v[i] = 1;
    rd_t r = {file, len, 0, 0};
    uint8_t used_in_frame[3] = {0, 0, 0};
    int zero_based = 0;

    unsigned last = rd_get(&r), cur = rd_get(&r);
    if (last != 0xFF || cur != 0xD8) { h->valid = 0; return -1; }
    last = rd_get(&r);
    cur = rd_get(&r);
    int got_sos = 0;
    while (h->valid) {                                   /* jpeg_scanner.cpp:371-403 */
        if (r.eof || last != 0xFF) { h->valid = 0; return -1; }
        if (cur == 0xC0 || cur == 0xC2) { h->frame_type = (int)cur; parse_sof(&r, h, used_in_frame, &zero_based); }
        else if (cur == 0xDB) parse_dqt(&r, h);
        else if (cur == 0xC4) parse_dht(&r, h);
        else if (cur == 0xDA) { parse_sos(&r, h, used_in_frame, zero_based); got_sos = 1; break; }
        else if (cur == 0xDD) {                          /* :323-331 */
            unsigned length = (rd_get(&r) << 8) + rd_get(&r);
            h->restart_interval = (rd_get(&r) << 8) + rd_get(&r);
            if (length != 4) h->valid = 0;
        }
        else if ((cur >= 0xE0 && cur <= 0xEF) || cur == 0xFE || (cur >= 0xF0 && cur <= 0xFD) || cur == 0xDC || cur == 0xDE || cur == 0xDF) skip_segment(&r);
        else if (cur == 0x01) {}
        else if (cur == 0xFF) { cur = rd_get(&r); continue; }
        /* any other marker: the reference only prints a message (:399) and goes on */
        last = rd_get(&r);
        cur = rd_get(&r);
    }
    if (!h->valid || !got_sos || r.eof) { h->valid = 0; return -1; }

    /* scan-byte walk, jpeg_scanner.cpp:405-433: find the EOI, reject any marker other than RSTn */
    h->scan_off = r.pos;
    size_t i = r.pos;
    for (;;) {
        if (i >= len) { h->valid = 0; return -1; }       /* "File ended prematurely" */
        if (file[i] != 0xFF) { i++; continue; }
        if (i + 1 >= len) { h->valid = 0; return -1; }
        unsigned n = file[i + 1];
        if (n == 0xD9) break;
        if (n == 0x00 || (n >= 0xD0 && n <= 0xD7)) { i += 2; if (n == 0x00 && i > len) { h->valid = 0; return -1; } continue; }
        if (n == 0xFF) { i++; continue; }
        h->valid = 0;                                    /* "Invalid marker during compressed data scan" */
        return -1;
    }
    h->scan_len = i - h->scan_off;
    return 0;
}

long rs_unstuff(const uint8_t *scan, size_t len, uint8_t *out, uint32_t *seg_starts, int max_seg, int *nseg) {
    size_t o = 0;
    int ns = 0;
    for (size_t i = 0; i < len;) {
        unsigned b = scan[i];
        if (b != 0xFF) { out[o++] = (uint8_t)b; i++; continue; }
        unsigned n = (i + 1 < len) ? scan[i + 1] : 0xFF;
        if (n == 0x00) { out[o++] = 0xFF; i += 2; }
        else if (n >= 0xD0 && n <= 0xD7) { if (seg_starts && ns < max_seg) seg_starts[ns] = (uint32_t)o; ns++; i += 2; }
        else if (n == 0xFF) i++;
        else return -1;
    }
    if (nseg) *nseg = ns;
    return (long)o;
}

uint32_t rs_blocks_per_mcu(const rs_header *h) {
    uint32_t n = 0;
    for (uint32_t j = 0; j < h->ncomp; j++) n += (uint32_t)h->comp_h[j] * h->comp_v[j];
    return n;
}

uint32_t rs_num_mcus(const rs_header *h) {
    return ((h->mcu_h + h->vs - 1) / h->vs) * ((h->mcu_w + h->hs - 1) / h->hs);
}

uint32_t rs_num_chunks(const rs_header *h, int M) {
    uint32_t pw = (h->mcu_w_real + 1) / 2 * 2, ph = (h->mcu_h_real + 1) / 2 * 2;
    return (pw * ph + (uint32_t)M - 1) / (uint32_t)M;
}

/* ------------------------------------------------------------------ entropy decode */

typedef struct {
    const uint8_t *d;
    size_t n, byte;
    unsigned bit;
} bits_t;

/* BitReader::read_bit, jpeg.h:91-100 */
static int get_bit(bits_t *b) {
    if (b->byte >= b->n) return -1;
    int v = (b->d[b->byte] >> (7 - b->bit)) & 1;
    if (++b->bit == 8) { b->bit = 0; b->byte++; }
    return v;
}
/* BitReader::read_bits, jpeg.h:102-113 */
static int get_bits(bits_t *b, unsigned n) {
    int v = 0;
    for (unsigned i = 0; i < n; i++) {
        int t = get_bit(b);
        if (t < 0) return -1;
        v = (v << 1) | t;
    }
    return v;
}
/* BitReader::align, jpeg.h:115-121 */
static void bit_align(bits_t *b) {
    if (b->byte >= b->n) return;
    if (b->bit) { b->bit = 0; b->byte++; }
}

typedef struct { uint32_t code[162]; } codes_t;

/* generate_codes, jpeg_scanner.cpp:438-448 */
static void make_codes(const rs_huff *t, codes_t *c) {
    uint32_t code = 0;
    for (int l = 0; l < 16; l++) {
        for (unsigned j = t->offsets[l]; j < t->offsets[l + 1]; j++) c->code[j] = code++;
        code <<= 1;
    }
}

/* get_next_symbol, jpeg_scanner.cpp:450-465 (0xFF doubles as the error value, as in the reference) */
static unsigned next_symbol(bits_t *b, const rs_huff *t, const codes_t *c) {
    uint32_t cw = 0;
    for (int l = 0; l < 16; l++) {
        int bit = get_bit(b);
        if (bit < 0) return 0xFF;
        cw = (cw << 1) | (uint32_t)bit;
        for (unsigned j = t->offsets[l]; j < t->offsets[l + 1]; j++)
            if (cw == c->code[j]) return t->symbols[j];
    }
    return 0xFF;
}

/* decode_MCU_component, baseline branch, jpeg_scanner.cpp:468-520.  zz[] is in zig-zag order. */
static int decode_unit(bits_t *b, int16_t *zz, int *pred, const rs_huff *dt, const codes_t *dcod, const rs_huff *at, const codes_t *acod) {
    unsigned len = next_symbol(b, dt, dcod);
    if (len == 0xFF || len > 11) return 0;
    int v = get_bits(b, len);
    if (v < 0) return 0;
    if (len != 0 && v < (1 << (len - 1))) v -= (1 << len) - 1;
    zz[0] = (int16_t)(v + *pred);
    *pred = zz[0];
    for (unsigned i = 1; i < 64; i++) {
        unsigned s = next_symbol(b, at, acod);
        if (s == 0xFF) return 0;
        if (s == 0) return 1;
        unsigned run = s >> 4, sz = s & 15;
        if (i + run >= 64) return 0;
        i += run;
        if (sz > 10) return 0;
        v = get_bits(b, sz);
        if (v < 0) return 0;
        if (sz == 0) v = 0; else if (v < (1 << (sz - 1))) v -= (1 << sz) - 1;
        zz[i] = (int16_t)v;
    }
    return 1;
}

int rs_huffman_zz(const rs_header *h, const uint8_t *file, int16_t *coef_zz, int restart_mode) {
    uint32_t bpm = rs_blocks_per_mcu(h), nmcu = rs_num_mcus(h);
    memset(coef_zz, 0, (size_t)nmcu * bpm * 64 * sizeof(int16_t));
    uint8_t *data = (uint8_t *)malloc(h->scan_len + 1);
    long n = rs_unstuff(file + h->scan_off, h->scan_len, data, NULL, 0, NULL);
    if (n < 0) { free(data); return 1; }
    codes_t dcod[4], acod[4];
    for (int i = 0; i < 4; i++) { make_codes(&h->dc[i], &dcod[i]); make_codes(&h->ac[i], &acod[i]); }
    bits_t b = {data, (size_t)n, 0, 0};
    int pred[3] = {0, 0, 0};
    int16_t *out = coef_zz;
    uint32_t mcu = 0;
    int ok = 1;
    /* decode_Huffman_data loop nest, jpeg_scanner.cpp:721-753 */
    for (uint32_t y = 0; y < h->mcu_h && ok; y += h->vs) {
        for (uint32_t x = 0; x < h->mcu_w && ok; x += h->hs, mcu++) {
            if (h->restart_interval != 0) {
                int hit = (restart_mode == RS_RESTART_REFQUIRK)
                              ? ((y * h->mcu_w_real + x) % h->restart_interval == 0)   /* :723, wrong for subsampled files */
                              : (mcu % h->restart_interval == 0);                       /* what T.81 says */
                if (hit) { pred[0] = pred[1] = pred[2] = 0; bit_align(&b); }
            }
            for (uint32_t j = 0; j < h->ncomp && ok; j++)
                for (uint32_t k = 0; k < (uint32_t)h->comp_v[j] * h->comp_h[j] && ok; k++, out += 64)
                    ok = decode_unit(&b, out, &pred[j], &h->dc[h->dc_id[j]], &dcod[h->dc_id[j]], &h->ac[h->ac_id[j]], &acod[h->ac_id[j]]);
        }
    }
    free(data);
    return ok ? 0 : 1;
}

/* destination of the data unit at 8x8 position (py,px), jpeg_scanner.cpp:733-741 / bmp_writer.cpp:51-56 */
static size_t ref_slot(uint32_t W, uint32_t py, uint32_t px, uint32_t comp, int M) {
    uint32_t idx = py * W + px;
    uint32_t blk = (idx / (W * 2)) * ((W + 1) / 2) + ((idx % W) / 2);
    uint32_t pos = ((idx / W) % 2) * 2 + ((idx % W) % 2);
    uint32_t per = (uint32_t)M / 4;
    return (size_t)(blk / per) * (64u * M * 3) + (size_t)(blk % per) * 768 + comp * 256 + pos * 64;
}

void rs_coef_to_ref_mcus(const rs_header *h, const int16_t *coef_zz, int16_t *mcus, int M) {
    memset(mcus, 0, (size_t)rs_num_chunks(h, M) * 64 * M * 3 * sizeof(int16_t));
    const int16_t *in = coef_zz;
    for (uint32_t y = 0; y < h->mcu_h; y += h->vs)
        for (uint32_t x = 0; x < h->mcu_w; x += h->hs)
            for (uint32_t j = 0; j < h->ncomp; j++)
                for (uint32_t v = 0; v < h->comp_v[j]; v++)
                    for (uint32_t hh = 0; hh < h->comp_h[j]; hh++, in += 64) {
                        int16_t *dst = mcus + ref_slot(h->mcu_w_real, y + v, x + hh, j, M);
                        dst[0] = in[0];
                        for (int i = 1; i < 64; i++)           /* ascending order: the later index wins at 38 */
                            if (in[i] != 0) dst[QMAP[i]] = in[i];
                    }
}

void rs_metadata(const rs_header *h, uint32_t md[276], int M) {
    memset(md, 0, 276 * sizeof(uint32_t));
    md[0] = h->mcu_h; md[1] = h->mcu_w; md[2] = h->mcu_h_real; md[3] = h->mcu_w_real;
    md[4] = h->ncomp; md[5] = h->vs; md[6] = h->hs;
    for (uint32_t j = 0; j < h->ncomp; j++) {
        md[7 + j] = h->qt_id[j];
        md[7 + h->ncomp + j] = h->comp_h[j];
        md[7 + 2 * h->ncomp + j] = h->comp_v[j];
    }
    md[17] = h->height; md[18] = h->width; md[19] = (uint32_t)M;
    for (int t = 0; t < 4; t++) {
        if (!h->qt_set[t]) break;                              /* decoder_host.cpp:174 stops at the first unset table */
        for (int i = 0; i < 64; i++) md[20 + t * 64 + QMAP[i]] = h->qt_zz[t][i];
    }
}

/* ------------------------------------------------------------------ the DPU program */

/* One 1-D pass of idct_component, decoder_dpu.c:219-267 (rows) == :271-319 (columns). */
void rs_idct8(const int32_t in[8], int32_t out[8]) {
    int32_t g0 = (in[0] * 181) >> 5, g1 = (in[4] * 181) >> 5, g2 = (in[2] * 59) >> 3, g3 = (in[6] * 49) >> 4;
    int32_t g4 = (in[5] * 71) >> 4, g5 = (in[1] * 251) >> 5, g6 = (in[7] * 25) >> 4, g7 = (in[3] * 213) >> 5;
    int32_t f4 = g4 - g7, f5 = g5 + g6, f6 = g5 - g6, f7 = g4 + g7;
    int32_t e2 = g2 - g3, e3 = g2 + g3, e5 = f5 - f7, e7 = f5 + f7, e8 = f4 + f6;
    int32_t d2 = (e2 * 181) >> 7, d4 = (f4 * 277) >> 8, d5 = (e5 * 181) >> 7, d6 = (f6 * 669) >> 8, d8 = (e8 * 49) >> 6;
    int32_t c0 = g0 + g1, c1 = g0 - g1, c2 = d2 - e3, c4 = d4 + d8, c5 = d5 + e7, c6 = d6 - d8, c8 = c5 - c6;
    int32_t b0 = c0 + e3, b1 = c1 + c2, b2 = c1 - c2, b3 = c0 - e3, b4 = c4 - c8, b6 = c6 - e7;
    out[0] = (b0 + e7) >> 4; out[1] = (b1 + b6) >> 4; out[2] = (b2 + c8) >> 4; out[3] = (b3 + b4) >> 4;
    out[4] = (b3 - b4) >> 4; out[5] = (b2 - c8) >> 4; out[6] = (b1 - b6) >> 4; out[7] = (b0 - e7) >> 4;
}

static void idct_tile(int16_t *t) {
    int32_t in[8], out[8];
    for (int r = 0; r < 8; r++) {                              /* rows, results stored back as short (:260-267) */
        for (int k = 0; k < 8; k++) in[k] = t[r * 8 + k];
        rs_idct8(in, out);
        for (int k = 0; k < 8; k++) t[r * 8 + k] = (int16_t)out[k];
    }
    for (int c = 0; c < 8; c++) {                              /* columns (:312-319) */
        for (int k = 0; k < 8; k++) in[k] = t[k * 8 + c];
        rs_idct8(in, out);
        for (int k = 0; k < 8; k++) t[k * 8 + c] = (int16_t)out[k];
    }
}

/* decoder_dpu.c:376-382; the products are 32-bit and wrap. */
void rs_ycc_to_rgb(int y, int cb, int cr, int *r, int *g, int *b) {
    int32_t tr = (int32_t)(5880414u * (uint32_t)cr) >> 22;
    int32_t tg1 = (int32_t)(1442840u * (uint32_t)cb) >> 22;
    int32_t tg2 = (int32_t)(2994733u * (uint32_t)cr) >> 22;
    int32_t tb = (int32_t)(7432306u * (uint32_t)cb) >> 22;
    int R = y + tr + 128, G = y - tg1 - tg2 + 128, B = y + tb + 128;
    *r = R < 0 ? 0 : (R > 255 ? 255 : R);
    *g = G < 0 ? 0 : (G > 255 ? 255 : G);
    *b = B < 0 ? 0 : (B > 255 ? 255 : B);
}

/* convert_colorspace_component, decoder_dpu.c:361-390, done out of place into rgb[3][64]. */
static void colour_tile(const int16_t *blk, int ypos, int cpos, int vpos, int hpos, int vs, int hs, int16_t rgb[3][64]) {
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) {
            int p = y * 8 + x, cp = ((y / vs) + 4 * vpos) * 8 + (x / hs) + 4 * hpos;
            int r, g, b;
            rs_ycc_to_rgb(blk[0 * 256 + ypos * 64 + p], blk[1 * 256 + cpos * 64 + cp], blk[2 * 256 + cpos * 64 + cp], &r, &g, &b);
            rgb[0][p] = (int16_t)r; rgb[1][p] = (int16_t)g; rgb[2][p] = (int16_t)b;
        }
}

void rs_exec_mcus(const uint32_t *metadata, int16_t *mcus, int nchunk) {
    for (int c = 0; c < nchunk; c++) {
        const uint32_t *md = metadata + (size_t)c * 276;
        uint32_t M = md[19], ncomp = md[4], vs = md[5], hs = md[6];
        if (M == 0) M = metadata[19];                          /* idle DPU: buffer length is a build constant */
        int16_t *base = mcus + (size_t)c * 64 * M * 3;
        uint32_t nblk = md[19] / 4;                            /* decoder_dpu.c:130 */
        for (uint32_t bi = 0; bi < nblk; bi++) {
            int16_t *blk = base + (size_t)bi * 768;            /* [comp][pos][64] */
            for (uint32_t j = 0; j < ncomp && j < 3; j++) {    /* dequantize, :158-177 */
                const uint32_t *q = md + 20 + 64 * md[7 + j];
                for (int p = 0; p < 4; p++)
                    for (int k = 0; k < 64; k++) {
                        int16_t *v = &blk[j * 256 + p * 64 + k];
                        *v = (int16_t)((uint32_t)(int32_t)*v * q[k]);
                    }
            }
            for (int p = 0; p < 4; p++)                        /* idct, :179-207 */
                for (int j = 0; j < 3; j++) idct_tile(&blk[j * 256 + p * 64]);
            /* convert_colorspace, :323-359: which chroma tile / quadrant feeds which luma tile */
            int16_t rgb[4][3][64];
            int done[4] = {0, 0, 0, 0};
            if (vs == 1 && hs == 1) for (int p = 0; p < 4; p++) { colour_tile(blk, p, p, 0, 0, 1, 1, rgb[p]); done[p] = 1; }
            if (vs == 2 && hs == 1) for (int p = 0; p < 4; p++) { colour_tile(blk, p, p & 1, p >> 1, 0, 2, 1, rgb[p]); done[p] = 1; }
            if (vs == 1 && hs == 2) for (int p = 0; p < 4; p++) { colour_tile(blk, p, p & 2, 0, p & 1, 1, 2, rgb[p]); done[p] = 1; }
            if (vs == 2 && hs == 2) for (int p = 0; p < 4; p++) { colour_tile(blk, p, 0, p >> 1, p & 1, 2, 2, rgb[p]); done[p] = 1; }
            for (int p = 0; p < 4; p++)
                if (done[p])
                    for (int j = 0; j < 3; j++) memcpy(&blk[j * 256 + p * 64], rgb[p][j], 128);
        }
    }
}

/* ------------------------------------------------------------------ BMP */

size_t rs_bmp_size(uint32_t width, uint32_t height) { return 26 + (size_t)height * width * 3 + (size_t)(width % 4) * height; }

static uint8_t *put_le(uint8_t *p, uint32_t v, int n) { for (int i = 0; i < n; i++) *p++ = (uint8_t)(v >> (8 * i)); return p; }

size_t rs_mcus_to_bmp(const uint32_t md[276], const int16_t *mcus, uint8_t *bmp) {
    uint32_t width = md[18], height = md[17], W = md[3], M = md[19], pad = width % 4;
    uint8_t *p = bmp;
    *p++ = 'B'; *p++ = 'M';
    p = put_le(p, (uint32_t)rs_bmp_size(width, height), 4);
    p = put_le(p, 0, 4); p = put_le(p, 0x1A, 4); p = put_le(p, 12, 4);
    p = put_le(p, width, 2); p = put_le(p, height, 2); p = put_le(p, 1, 2); p = put_le(p, 24, 2);
    for (uint32_t y = height; y-- > 0;) {
        for (uint32_t x = 0; x < width; x++) {
            const int16_t *t = mcus + ref_slot(W, y / 8, x / 8, 0, (int)M) + (y % 8) * 8 + (x % 8);
            *p++ = (uint8_t)t[512]; *p++ = (uint8_t)t[256]; *p++ = (uint8_t)t[0];
        }
        for (uint32_t i = 0; i < pad; i++) *p++ = 0;
    }
    return (size_t)(p - bmp);
}

int rs_decode(const uint8_t *file, size_t len, int restart_mode, uint8_t *rgb, uint8_t *bmp) {
    rs_header h;
    if (rs_parse(file, len, &h) != 0) return -1;
    if (h.frame_type != 0xC0) return -2;
    const int M = 100;
    uint32_t nchunk = rs_num_chunks(&h, M);
    int16_t *zz = (int16_t *)malloc((size_t)rs_num_mcus(&h) * rs_blocks_per_mcu(&h) * 128);
    int16_t *mcus = (int16_t *)malloc((size_t)nchunk * 64 * M * 3 * 2);
    uint32_t *md = (uint32_t *)malloc((size_t)nchunk * 276 * 4);
    rs_huffman_zz(&h, file, zz, restart_mode);
    rs_coef_to_ref_mcus(&h, zz, mcus, M);
    rs_metadata(&h, md, M);
    for (uint32_t c = 1; c < nchunk; c++) memcpy(md + (size_t)c * 276, md, 276 * 4);
    rs_exec_mcus(md, mcus, (int)nchunk);
    if (bmp) rs_mcus_to_bmp(md, mcus, bmp);
    if (rgb)
        for (uint32_t y = 0; y < h.height; y++)
            for (uint32_t x = 0; x < h.width; x++) {
                const int16_t *t = mcus + ref_slot(h.mcu_w_real, y / 8, x / 8, 0, M) + (y % 8) * 8 + (x % 8);
                uint8_t *o = rgb + ((size_t)y * h.width + x) * 3;
                o[0] = (uint8_t)t[0]; o[1] = (uint8_t)t[256]; o[2] = (uint8_t)t[512];
            }
    free(zz); free(mcus); free(md);
    return 0;
}
