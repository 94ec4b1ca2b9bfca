// Entropy-stage core shared by the CUDA kernels (kernels_huff.cuh) and by the host-side emulation used in the
// CPU tests (tests/emu/huff_emu.cpp).  Everything here is plain integer code that compiles both as device and
// as host code; there is no CPU decode path in the library itself - the host build of these functions exists
// only so the algorithm (speculative sub-sequence decode + fix-up + ownership rule) can be checked without a GPU.
//
// What it restates (semantics, not code) - reference paths relative to the reference tree:
//   canonical code assignment        src/jpeg_scanner.cpp:438-448   generate_codes
//   symbol lookup                    src/jpeg_scanner.cpp:450-465   get_next_symbol (first match, shortest length)
//   baseline DC/AC unit decode       src/jpeg_scanner.cpp:467-520   decode_MCU_component
//   scan-byte filtering              src/jpeg_scanner.cpp:405-433   FF00 un-stuffing, RSTn and fill-byte removal
#pragma once
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define BJ_HD __host__ __device__ __forceinline__
#else
#define BJ_HD inline
#endif

namespace bj {

// ------------------------------------------------------------------------------------------------ lookup tables
// One table per Huffman table: kRootBits-bit root + second-level tables for longer codes, 16-bit entries.
//   leaf  : bit15 = 0, bits 12..8 = code length (1..16), bits 7..0 = symbol;  0 = no code with this prefix
//   link  : bit15 = 1, bits 14..12 = k (second-level index width, 1..16-kRootBits), bits 11..0 = offset of the
//           second-level table (2^k leaves) from the start of this table
constexpr int kRootBits = 10;
constexpr int kLutCap = 2048;              // entries per table (4 KB)
constexpr uint32_t kLutLink = 0x8000u;

// Host: build one table.  Returns the number of entries used, or -1 if the second-level tables do not fit.
// Over-subscribed (invalid) tables keep the reference's behaviour: the shortest matching code wins and codes
// that do not fit their length never match.
inline int build_lut(const uint8_t offsets[17], const uint8_t symbols[162], uint16_t *lut) {
    const int R = kRootBits;
    memset(lut, 0, sizeof(uint16_t) * kLutCap);
    uint8_t maxlen[1 << kRootBits];
    memset(maxlen, 0, sizeof(maxlen));
    uint32_t code = 0;
    for (int l = 1; l <= 16; l++) {                       // codes that fit the root
        for (unsigned j = offsets[l - 1]; j < offsets[l]; j++) {
            const uint32_t cv = code++;
            if (l > R || (cv >> l)) continue;
            const uint32_t first = cv << (R - l), cnt = 1u << (R - l);
            for (uint32_t i = 0; i < cnt; i++)
                if (lut[first + i] == 0) lut[first + i] = (uint16_t)((l << 8) | symbols[j]);
        }
        code <<= 1;
    }
    code = 0;
    for (int l = 1; l <= 16; l++) {                       // size the second-level tables
        for (unsigned j = offsets[l - 1]; j < offsets[l]; j++) {
            const uint32_t cv = code++;
            if (l <= R || (cv >> l)) continue;
            const uint32_t pre = cv >> (l - R);
            if (lut[pre] != 0 && !(lut[pre] & kLutLink)) continue;      // a shorter code owns this prefix
            if (l > maxlen[pre]) maxlen[pre] = (uint8_t)l;
            lut[pre] = (uint16_t)kLutLink;
        }
        code <<= 1;
    }
    int next = 1 << R;
    for (int pre = 0; pre < (1 << R); pre++) {
        if (!maxlen[pre]) continue;
        const int k = maxlen[pre] - R;
        if (next + (1 << k) > kLutCap) return -1;
        lut[pre] = (uint16_t)(kLutLink | (k << 12) | next);
        next += 1 << k;
    }
    code = 0;
    for (int l = 1; l <= 16; l++) {                       // fill them
        for (unsigned j = offsets[l - 1]; j < offsets[l]; j++) {
            const uint32_t cv = code++;
            if (l <= R || (cv >> l)) continue;
            const uint32_t pre = cv >> (l - R);
            if (!(lut[pre] & kLutLink)) continue;
            const int k = (lut[pre] >> 12) & 7, rem = l - R;
            const uint32_t base = lut[pre] & 0xFFFu;
            const uint32_t first = (cv & ((1u << rem) - 1)) << (k - rem), cnt = 1u << (k - rem);
            for (uint32_t i = 0; i < cnt; i++)
                if (lut[base + first + i] == 0) lut[base + first + i] = (uint16_t)((l << 8) | symbols[j]);
        }
        code <<= 1;
    }
    return next;
}

// win = the next 32 bits of the stream, MSB first.  Returns (length << 8) | symbol, 0 if no code matches.
BJ_HD uint32_t lut_lookup(const uint16_t *lut, uint32_t win) {
    uint32_t e = lut[win >> (32 - kRootBits)];
    if (e & kLutLink) {
        const uint32_t k = (e >> 12) & 7u;
        e = lut[(e & 0xFFFu) + ((win << kRootBits) >> (32 - k))];
    }
    return e;
}

// ------------------------------------------------------------------------------------------------ scan bytes
// Classification of one raw scan byte from its neighbours (see the scan-byte walk cited above):
//   FF xx : FF kept iff xx == 00 (stuffed data byte); FF before FF is a fill byte; FF before RSTn is dropped
//   xx after FF : dropped (the stuffing zero or the marker code).  RSTn marks a segment boundary.
BJ_HD bool scan_keep(unsigned prev, unsigned b, unsigned next) { return b == 0xFFu ? next == 0x00u : prev != 0xFFu; }
BJ_HD bool scan_is_rst(unsigned prev, unsigned b) { return prev == 0xFFu && b >= 0xD0u && b <= 0xD7u; }

// ------------------------------------------------------------------------------------------------ bit reader
// The un-stuffed stream is stored as 32-bit words whose most significant byte is the earliest byte (the
// un-stuff kernel writes byte o to address o ^ 3), so a window is two aligned words and one funnel shift.
BJ_HD uint32_t funnel_l(uint32_t hi, uint32_t lo, uint32_t s) {      // top 32 bits of (hi:lo) << s, s in 0..31
#ifdef __CUDA_ARCH__
    return __funnelshift_l(lo, hi, s);
#else
    return s ? (hi << s) | (lo >> (32 - s)) : hi;
#endif
}

struct BitReader {
    const uint32_t *w;      // image's un-stuffed stream, word 0 = bits 0..31
    uint32_t wi, cur, nxt;
    BJ_HD uint32_t ld(uint32_t i) const {
#ifdef __CUDA_ARCH__
        return __ldg(w + i);
#else
        return w[i];
#endif
    }
    BJ_HD void seek(uint32_t p) { wi = p >> 5; cur = ld(wi); nxt = ld(wi + 1); }
    // p never moves by more than 31 bits between calls
    BJ_HD uint32_t window(uint32_t p) {
        const uint32_t i = p >> 5;
        if (i != wi) { cur = nxt; nxt = ld(i + 1); wi = i; }
        return funnel_l(cur, nxt, p & 31u);
    }
};

// ------------------------------------------------------------------------------------------------ decoder state
// A decode position: bit offset in the image's un-stuffed stream, index of the data unit inside the MCU (selects
// the component, hence the table pair), zig-zag index inside the unit (0 = a DC symbol is next).
struct HuffState {
    uint32_t p;
    uint32_t cz;            // (c << 8) | z
};
BJ_HD bool same_state(const HuffState &a, const HuffState &b) { return a.p == b.p && a.cz == b.cz; }

// Per image: geometry of an MCU and where each component's tables sit (entry offsets into the staged tables).
struct HuffGeom {
    uint32_t bpm;           // data units per MCU
    uint32_t ny;            // luma units per MCU (hs * vs); unit c belongs to component c < ny ? 0 : c - ny + 1
    uint32_t tab[3];        // per component: dc table offset | ac table offset << 16  (units of kLutCap entries... see below)
};
BJ_HD uint32_t comp_of(const HuffGeom &g, uint32_t c) { return c < g.ny ? 0u : c - g.ny + 1u; }
BJ_HD uint32_t tabs_of(const HuffGeom &g, uint32_t c) { return c < g.ny ? g.tab[0] : (c == g.ny ? g.tab[1] : g.tab[2]); }

// What one step consumed/produced.
struct Sym {
    uint32_t bits;          // stream bits consumed (>= 1)
    int32_t value;          // extended coefficient / DC difference (0 when size == 0)
    uint32_t run, size;
    bool eob;               // AC end-of-block
    bool bad;               // the reference would stop here: no code, DC category > 11, AC size > 10
};

// Decode one symbol (+ its magnitude bits) from the window.  dc: a DC symbol is expected.
BJ_HD Sym decode_symbol(const uint16_t *lut, uint32_t win, bool dc) {
    Sym s;
    const uint32_t e = lut_lookup(lut, win);
    const uint32_t len = e >> 8, sym = e & 0xFFu;
    s.run = dc ? 0u : sym >> 4;
    s.size = dc ? sym : (sym & 15u);
    s.eob = !dc && sym == 0u && e != 0u;
    s.bad = e == 0u || (dc ? sym > 11u : (s.size > 10u));
    if (s.bad) s.size = 0;
    const uint32_t t = win << len;                        // len <= 16
    const uint32_t raw = s.size ? (t >> (32u - s.size)) : 0u;
    // magnitude extension, src/jpeg_scanner.cpp:480-482 / :513-516: first bit 0 => negative
    s.value = (int32_t)raw - ((s.size && !(t >> 31)) ? (int32_t)((1u << s.size) - 1u) : 0);
    s.bits = (len ? len : 1u) + s.size;
    return s;
}

// Advance (c, z) over one symbol.  Returns true when the unit ended.  `overflow` = the reference's
// "i + run >= 64" error (src/jpeg_scanner.cpp:497-500); the speculative passes treat it as a unit end.
BJ_HD bool advance_state(const HuffGeom &g, uint32_t &c, uint32_t &z, const Sym &s, bool *overflow) {
    *overflow = false;
    uint32_t zn;
    if (z == 0) zn = 1;
    else if (s.eob) zn = 64;
    else {
        *overflow = z + s.run >= 64u;
        zn = z + s.run + 1u;
    }
    if (zn >= 64u) { z = 0; c = (c + 1u == g.bpm) ? 0u : c + 1u; return true; }
    z = zn;
    return false;
}

// ------------------------------------------------------------------------------------------------ pass 1: synchronise
// Totals of one sub-sequence for the prefix sums: data units whose DC symbol starts inside it and, per component,
// the sum of those DC differences (mod 2^16 arithmetic is enough: the reference truncates the predictor to a
// short after every unit, src/jpeg_scanner.cpp:485-486, which is a ring homomorphism).
struct SubTotals {
    uint32_t n;
    uint32_t dc[3];
};

// Decode from `st` until the position reaches end_bit (first symbol boundary at or after it).  Errors do not
// stop a speculative decode (a wrong guess must not poison its successors): a bad symbol advances at least one
// bit and the unit simply continues.  With the true entry state the result is the true exit state up to the
// first real error, which the write pass detects and reports.
BJ_HD HuffState decode_span(BitReader &rd, const uint16_t *luts, const HuffGeom &g, HuffState st, uint32_t end_bit,
                            SubTotals *tot) {
    uint32_t p = st.p, c = st.cz >> 8, z = st.cz & 0xFFu;
    uint32_t n = 0, d0 = 0, d1 = 0, d2 = 0;
    uint32_t tabs = tabs_of(g, c);
    if (p < end_bit) rd.seek(p);
    while (p < end_bit) {
        const bool dc = z == 0;
        const Sym s = decode_symbol(luts + (dc ? (tabs & 0xFFFFu) : (tabs >> 16)), rd.window(p), dc);
        if (dc) {
            n++;
            const uint32_t k = comp_of(g, c);
            const uint32_t v = s.bad ? 0u : (uint32_t)s.value;
            if (k == 0) d0 += v; else if (k == 1) d1 += v; else d2 += v;
        }
        p += s.bits;
        bool ovf;
        if (advance_state(g, c, z, s, &ovf)) tabs = tabs_of(g, c);
    }
    tot->n = n; tot->dc[0] = d0; tot->dc[1] = d1; tot->dc[2] = d2;
    HuffState o;
    o.p = p; o.cz = (c << 8) | z;
    return o;
}

// ------------------------------------------------------------------------------------------------ pass 2: write
// Ownership rule: a data unit belongs to the sub-sequence in which its DC symbol starts; the owner decodes the
// whole unit (running past its own end if necessary) and stores all 64 coefficients at once, so every unit is
// written exactly once, by one thread, with no zero-fill pass.  A sub-sequence entered mid-unit first skips to
// the end of that unit without storing anything.
//
// Sink concept:  void put(uint32_t zz, int16_t v);  void flush(uint32_t du);   (flush also clears the unit)
struct WriteResult {
    uint32_t first_zero;    // first unit index that must read as zero because the reference stopped; UINT32_MAX if none
};

template <class Sink>
BJ_HD WriteResult write_span(BitReader &rd, const uint16_t *luts, const HuffGeom &g, HuffState st, uint32_t end_bit,
                             uint32_t data_end_bit, uint32_t du, uint32_t du_end, bool last_of_segment,
                             const uint32_t pred_in[3], Sink &sink) {
    WriteResult res;
    res.first_zero = 0xFFFFFFFFu;
    uint32_t p = st.p, c = st.cz >> 8, z = st.cz & 0xFFu;
    uint32_t pred0 = pred_in[0], pred1 = pred_in[1], pred2 = pred_in[2];
    uint32_t tabs = tabs_of(g, c);
    bool owned = false;
    rd.seek(p);
    for (;;) {
        const bool dc = z == 0;
        if (dc) {
            if (p >= end_bit || du >= du_end) break;
            owned = true;
        } else if (!owned && p >= end_bit) break;
        const Sym s = decode_symbol(luts + (dc ? (tabs & 0xFFFFu) : (tabs >> 16)), rd.window(p), dc);
        p += s.bits;
        const uint32_t z0 = z, k = comp_of(g, c);             // a DC symbol never ends a unit, so c is this unit's
        bool ovf;
        const bool ended = advance_state(g, c, z, s, &ovf);
        if (owned) {
            // the reference's failure points, src/jpeg_scanner.cpp:470-478 (DC) and :490-511 (AC); running out of
            // bits inside a symbol fails too (BitReader::read_bit returns -1)
            if (s.bad || ovf || p > data_end_bit) {
                // a failed DC leaves the unit untouched (zero); a failed AC keeps what was stored before it
                if (dc) res.first_zero = du;
                else { sink.flush(du); res.first_zero = du + 1; }
                return res;
            }
            if (dc) {
                uint32_t pr;
                if (k == 0) { pred0 += (uint32_t)s.value; pr = pred0; }
                else if (k == 1) { pred1 += (uint32_t)s.value; pr = pred1; }
                else { pred2 += (uint32_t)s.value; pr = pred2; }
                sink.put(0, (int16_t)(uint16_t)pr);
            } else if (!s.eob) {
                sink.put(z0 + s.run, (int16_t)s.value);       // size 0 (ZRL and friends) stores a literal 0
            }
            if (ended) { sink.flush(du); du++; owned = false; }
        }
        if (ended) tabs = tabs_of(g, c);
    }
    // the last sub-sequence of a segment must have produced the segment's last unit
    if (last_of_segment && du < du_end) res.first_zero = du;
    return res;
}

}  // namespace bj
