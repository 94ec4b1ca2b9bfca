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
// One table per Huffman table: kRootBits-bit root + second-level tables for longer codes.  Entries are 32 bits
// and carry everything a decode step needs, precomputed per table class (DC / AC), so the inner loop is a load,
// two adds and a compare:
//   bits  5..0   stream bits the step consumes: code length + magnitude bits (>= 1, also for "no such code")
//   bits 12..6   zig-zag advance: 1 for a DC symbol, run + 1 for an AC symbol, 64 for end-of-block
//   bits 16..13  magnitude size (0 for a step the reference would refuse)
//   bits 21..17  code length
//   bit  22      bad: the reference stops here - no code, DC category > 11, AC size > 10
//                (src/jpeg_scanner.cpp:470-478, :490-511)
//   bit  23      AC end-of-block
//   bit  31      link (root only): bits 19..16 = k, bits 15..0 = offset of a 2^k-entry second-level table
constexpr int kRootBitsDC = 9, kRootBitsAC = 11;      // > 11-bit AC codes are ~0.2 % of symbols at q = 90
constexpr int kLutCapDC = 1024, kLutCapAC = 2560;     // entries per table (4 KB / 10 KB)
constexpr uint32_t kLutLink = 0x80000000u;
constexpr uint32_t kLutBad = 1u << 22;
constexpr uint32_t kLutEob = 1u << 23;
BJ_HD constexpr int lut_root_bits(bool ac) { return ac ? kRootBitsAC : kRootBitsDC; }
BJ_HD constexpr int lut_cap(bool ac) { return ac ? kLutCapAC : kLutCapDC; }

inline uint32_t lut_leaf(int len, unsigned sym, bool ac) {
    const unsigned run = ac ? sym >> 4 : 0, size = ac ? (sym & 15u) : sym;
    const bool eob = ac && sym == 0;
    const bool bad = ac ? size > 10 : sym > 11;
    const unsigned sz = bad ? 0u : size;
    return (uint32_t)(len + sz) | ((eob ? 64u : run + 1u) << 6) | (sz << 13) | ((uint32_t)len << 17) | (bad ? kLutBad : 0u) | (eob ? kLutEob : 0u);
}
constexpr uint32_t kLutNoCode = 1u | (1u << 6) | kLutBad;      // no code with this prefix: skip one bit

// Host: build one table.  Returns the number of entries used, or -1 if the second-level tables do not fit.
// Over-subscribed (invalid) tables keep the reference's behaviour: the shortest matching code wins and codes
// that do not fit their length never match (get_next_symbol compares the l-bit prefix with the stored code).
inline int build_lut(const uint8_t offsets[17], const uint8_t symbols[162], bool ac, uint32_t *lut) {
    const int R = lut_root_bits(ac), kLutCap = lut_cap(ac);
    for (int i = 0; i < kLutCap; i++) lut[i] = 0;
    uint8_t maxlen[1 << kRootBitsAC];
    memset(maxlen, 0, sizeof(maxlen));
    uint32_t code = 0;
    for (int l = 1; l <= 16; l++) {                       // codes that fit the root
        for (unsigned j = offsets[l - 1]; j < offsets[l]; j++) {
            const uint32_t cv = code++;
            if (l > R || (cv >> l)) continue;
            const uint32_t first = cv << (R - l), cnt = 1u << (R - l);
            for (uint32_t i = 0; i < cnt; i++)
                if (lut[first + i] == 0) lut[first + i] = lut_leaf(l, symbols[j], ac);
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
            lut[pre] = kLutLink;
        }
        code <<= 1;
    }
    int next = 1 << R;
    for (int pre = 0; pre < (1 << R); pre++) {
        if (!maxlen[pre]) continue;
        const int k = maxlen[pre] - R;
        if (next + (1 << k) > kLutCap) return -1;
        lut[pre] = kLutLink | ((uint32_t)k << 16) | (uint32_t)next;
        next += 1 << k;
    }
    code = 0;
    for (int l = 1; l <= 16; l++) {                       // fill them
        for (unsigned j = offsets[l - 1]; j < offsets[l]; j++) {
            const uint32_t cv = code++;
            if (l <= R || (cv >> l)) continue;
            const uint32_t pre = cv >> (l - R);
            if (!(lut[pre] & kLutLink)) continue;
            const int k = (lut[pre] >> 16) & 15, rem = l - R;
            const uint32_t base = lut[pre] & 0xFFFFu;
            const uint32_t first = (cv & ((1u << rem) - 1)) << (k - rem), cnt = 1u << (k - rem);
            for (uint32_t i = 0; i < cnt; i++)
                if (lut[base + first + i] == 0) lut[base + first + i] = lut_leaf(l, symbols[j], ac);
        }
        code <<= 1;
    }
    for (int i = 0; i < kLutCap; i++) if (lut[i] == 0) lut[i] = kLutNoCode;
    return next;
}

// win = the next 32 bits of the stream, MSB first.
BJ_HD uint32_t lut_lookup(const uint32_t *lut, uint32_t win, bool ac) {
    const uint32_t R = ac ? kRootBitsAC : kRootBitsDC;
    uint32_t e = lut[win >> (32u - R)];
    if (e & kLutLink) e = lut[(e & 0xFFFFu) + ((win << R) >> (32u - ((e >> 16) & 15u)))];
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
    const uint32_t *q;      // -> word wi of the image's un-stuffed stream (word 0 = bits 0..31)
    uint32_t wi, cur, nxt, nx2;   // words wi, wi+1, wi+2: the load for a word is issued one word ahead of its use
    BJ_HD static uint32_t ld(const uint32_t *a) {
#ifdef __CUDA_ARCH__
        return __ldg(a);
#else
        return *a;
#endif
    }
    BJ_HD void seek(const uint32_t *words, uint32_t p) { wi = p >> 5; q = words + wi; cur = ld(q); nxt = ld(q + 1); nx2 = ld(q + 2); }
    // p never moves by more than 31 bits between calls, so the word index grows by at most one
    BJ_HD uint32_t window(uint32_t p) {
        if ((p >> 5) != wi) { cur = nxt; nxt = nx2; q++; wi++; nx2 = ld(q + 2); }
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
    uint32_t tab[3];        // per component: DC table entry offset | AC table entry offset << 16 (from the staged base)
};
BJ_HD uint32_t comp_of(const HuffGeom &g, uint32_t c) { return c < g.ny ? 0u : c - g.ny + 1u; }
BJ_HD uint32_t tabs_of(const HuffGeom &g, uint32_t c) { return c < g.ny ? g.tab[0] : (c == g.ny ? g.tab[1] : g.tab[2]); }

// Magnitude extension of the `size` bits that follow a `len`-bit code in the window
// (src/jpeg_scanner.cpp:480-482 / :513-516: first bit 0 => negative).  size 0 gives 0.
BJ_HD int32_t extend_value(uint32_t win, uint32_t len, uint32_t size) {
    const uint32_t t = win << len;                                        // len <= 16
    const uint32_t raw = funnel_l(0u, t, size);                           // top `size` bits of t
    return (int32_t)raw + (((int32_t)t >= 0) ? 1 - (int32_t)(1u << size) : 0);
}

// ------------------------------------------------------------------------------------------------ pass 1: synchronise
// Decode from `st` until the position reaches end_bit (first symbol boundary at or after it).  Only the decoder
// state is tracked - no values - plus the number of data units whose DC symbol starts inside the span (for the
// prefix sum that tells every sub-sequence which unit it starts in).  Errors do not stop a speculative decode
// (a wrong guess must not poison its successors): a refused symbol consumes its code bits (at least one) and the
// unit simply continues; an over-long run ends the unit.  With the true entry state the result is the true exit
// state up to the first real error, which the write pass detects and reports.
BJ_HD HuffState decode_span(const uint32_t *words, const uint32_t *luts, const HuffGeom &g, HuffState st, uint32_t end_bit,
                            uint32_t *units_started) {
    uint32_t p = st.p, c = st.cz >> 8, z = st.cz & 0xFFu;
    uint32_t ends = 0;
    const uint32_t entered_mid = z != 0 ? 1u : 0u;
    uint32_t tabs = tabs_of(g, c);
    BitReader rd;
    if (p < end_bit) rd.seek(words, p);
    while (p < end_bit) {
        const uint32_t win = rd.window(p);
        const bool ac = z != 0;
        const uint32_t e = lut_lookup(luts + (ac ? (tabs >> 16) : (tabs & 0xFFFFu)), win, ac);
        p += e & 63u;
        z += (e >> 6) & 127u;
        if (z >= 64u) {
            z = 0;
            ends++;
            c = (c + 1u == g.bpm) ? 0u : c + 1u;
            tabs = tabs_of(g, c);
        }
    }
    // started = ended + (one still open at the exit) - (the one that was already open at the entry)
    *units_started = ends + (z != 0 ? 1u : 0u) - entered_mid;
    HuffState o;
    o.p = p; o.cz = (c << 8) | z;
    return o;
}

// ------------------------------------------------------------------------------------------------ pass 2: write
// Ownership rule: a data unit belongs to the sub-sequence in which its DC symbol starts; the owner decodes the
// whole unit (running past its own end if necessary) and stores all 64 coefficients at once, so every unit is
// written exactly once, by one thread, with no zero-fill pass.  A sub-sequence entered mid-unit first skips to
// the end of that unit without storing anything.  DC DIFFERENCES go to a separate compact plane (2 bytes per
// unit); the prediction sums over it are a separate, tiny scan (K1c) and slot 0 of every unit stays zero.
//
// Sink concept:  void put(uint32_t zz, int16_t v);  void dc(uint32_t du, int16_t diff);  void flush(uint32_t du);
// (flush also clears the staged unit)
struct WriteResult {
    uint32_t first_zero;    // first unit index that must read as zero because the reference stopped; UINT32_MAX if none
};

template <class Sink>
BJ_HD WriteResult write_span(const uint32_t *words, const uint32_t *luts, const HuffGeom &g, HuffState st, uint32_t end_bit,
                             uint32_t data_end_bit, uint32_t du, uint32_t du_end, bool last_of_segment, Sink &sink) {
    WriteResult res;
    res.first_zero = 0xFFFFFFFFu;
    uint32_t p = st.p, c = st.cz >> 8, z = st.cz & 0xFFu;
    uint32_t tabs = tabs_of(g, c);
    bool owned = false;
    BitReader rd;
    rd.seek(words, p);
    for (;;) {
        const bool ac = z != 0;
        if ((!ac || !owned) && p >= end_bit) break;         // a new unit would start (or a foreign one continue) past my end
        if (!ac && du >= du_end) break;                     // the segment's last unit is done: the rest is padding
        owned = owned || !ac;
        const uint32_t win = rd.window(p);
        const uint32_t e = lut_lookup(luts + (ac ? (tabs >> 16) : (tabs & 0xFFFFu)), win, ac);
        p += e & 63u;
        const uint32_t zn = z + ((e >> 6) & 127u);
        if (owned) {
            // the reference's failure points: refused symbol, run past the end of the unit ("i + run >= 64",
            // src/jpeg_scanner.cpp:497-500), or bits running out inside a symbol (BitReader::read_bit returns -1)
            if ((e & kLutBad) || (!(e & kLutEob) && zn > 64u) || p > data_end_bit) {
                // a failed DC leaves the unit untouched (zero); a failed AC keeps what was stored before it
                if (!ac) res.first_zero = du;
                else { sink.flush(du); res.first_zero = du + 1; }
                return res;
            }
            const int32_t v = extend_value(win, (e >> 17) & 31u, (e >> 13) & 15u);
            if (!ac) sink.dc(du, (int16_t)v);               // |diff| < 2^11
            else if (v != 0) sink.put(zn - 1u, (int16_t)v); // size 0 (ZRL and friends) would store a literal 0: already there
        }
        if (zn >= 64u) {
            if (owned) { sink.flush(du); du++; owned = false; }
            z = 0;
            c = (c + 1u == g.bpm) ? 0u : c + 1u;
            tabs = tabs_of(g, c);
        } else z = zn;
    }
    // the last sub-sequence of a segment must have produced the segment's last unit
    if (last_of_segment && du < du_end) res.first_zero = du;
    return res;
}

// ------------------------------------------------------------------------------------------------ K1c: DC prediction
// Per MCU: turn the DC differences of its units (decode order: luma units, then Cb, then Cr) into predicted
// values given the three running predictors, which are updated (src/jpeg_scanner.cpp:485-486; everything mod 2^16).
BJ_HD void dc_predict_mcu(const HuffGeom &g, int16_t *dcs, uint32_t pred[3]) {
    for (uint32_t c = 0; c < g.bpm; c++) {
        const uint32_t k = comp_of(g, c);
        pred[k] += (uint32_t)(uint16_t)dcs[c];
        dcs[c] = (int16_t)(uint16_t)pred[k];
    }
}

}  // namespace bj
