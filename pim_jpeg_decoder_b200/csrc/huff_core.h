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
// One table per Huffman table: a kRootBits-bit root + second-level tables for longer codes.  Entries are 32 bits
// and carry everything a decode step needs, precomputed per table class (DC / AC).  The two low bytes ARE the
// increment of the packed decoder state S = (bit position << 8) | zig-zag index, so a step is: window, load, add.
//   bits  7..0   zig-zag advance: 1 for a DC symbol, run + 1 for an AC symbol, 64 for end-of-block
//   bits 15..8   stream bits the step consumes: code length + magnitude bits (1..27, also for "no such code")
//   bits 20..16  code length            } each at the bottom of a byte with the bits up to the next multiple of 5
//   bits 27..24  magnitude size         } clear, so `e >> 16` / `e >> 24` feed wrap-mode (mod 32) shifts unmasked;
//                                         size is 0 for a step the reference would refuse
//   bit  29      bad: the reference stops here - no code, DC category > 11, AC size > 10
//                (src/jpeg_scanner.cpp:470-478, :490-511)
//   bit  30      AC end-of-block
//   bit  31      link (root only): bits 19..16 = k, bits 15..0 = BYTE offset (from the table's first entry) of a
//                2^k-entry second-level table indexed by the k bits that follow the root bits
constexpr int kRootBits = 10;                         // > 10-bit codes are ~0.4 % of symbols at q = 90
// entries per table in the pools (root + second level).  For a canonical code the second level needs about one
// entry per long code plus < 64 per code length (long codes are numerically contiguous), far below these caps.
constexpr int kLutCapDC = 1536, kLutCapAC = 1792;
constexpr uint32_t kLutLink = 0x80000000u;
constexpr uint32_t kLutBad = 1u << 29;
constexpr uint32_t kLutEob = 1u << 30;
BJ_HD constexpr int lut_cap(bool ac) { return ac ? kLutCapAC : kLutCapDC; }
BJ_HD uint32_t lut_len(uint32_t e) { return (e >> 16) & 31u; }
BJ_HD uint32_t lut_size(uint32_t e) { return (e >> 24) & 15u; }

inline uint32_t lut_leaf(int len, unsigned sym, bool ac) {
    const unsigned run = ac ? sym >> 4 : 0, size = ac ? (sym & 15u) : sym;
    const bool eob = ac && sym == 0;
    const bool bad = ac ? size > 10 : sym > 11;
    const unsigned sz = bad ? 0u : size;
    return (eob ? 64u : run + 1u) | ((uint32_t)(len + sz) << 8) | ((uint32_t)len << 16) | (sz << 24) | (bad ? kLutBad : 0u) | (eob ? kLutEob : 0u);
}
constexpr uint32_t kLutNoCode = 1u | (1u << 8) | kLutBad;      // no code with this prefix: skip one bit

// Host: build one table.  Returns the number of entries used, or -1 if the second-level tables do not fit.
// Over-subscribed (invalid) tables keep the reference's behaviour: the shortest matching code wins and codes
// that do not fit their length never match (get_next_symbol compares the l-bit prefix with the stored code).
inline int build_lut(const uint8_t offsets[17], const uint8_t symbols[162], bool ac, uint32_t *lut) {
    const int R = kRootBits, kLutCap = lut_cap(ac);
    for (int i = 0; i < kLutCap; i++) lut[i] = 0;
    uint8_t maxlen[1 << kRootBits];
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
        lut[pre] = kLutLink | ((uint32_t)k << 16) | ((uint32_t)next * 4u);
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
            const uint32_t base = (lut[pre] & 0xFFFFu) / 4u;
            const uint32_t first = (cv & ((1u << rem) - 1)) << (k - rem), cnt = 1u << (k - rem);
            for (uint32_t i = 0; i < cnt; i++)
                if (lut[base + first + i] == 0) lut[base + first + i] = lut_leaf(l, symbols[j], ac);
        }
        code <<= 1;
    }
    for (int i = 0; i < next; i++) if (lut[i] == 0) lut[i] = kLutNoCode;
    return next;
}

// Where the staged tables live.  On the device: a 32-bit shared-memory address (loads are ld.shared with no
// generic-address arithmetic in the decode loop); on the host: a pointer.  Table positions are BYTE offsets.
struct LutMem {
#ifdef __CUDA_ARCH__
    uint32_t base;
    __device__ __forceinline__ void attach(const uint32_t *smem) { base = (uint32_t)__cvta_generic_to_shared(smem); }
    __device__ __forceinline__ uint32_t ld(uint32_t byte_off) const {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + byte_off));
        return v;
    }
#else
    const uint32_t *base;
    void attach(const uint32_t *tables) { base = tables; }
    uint32_t ld(uint32_t byte_off) const { return base[byte_off >> 2]; }
#endif
};

// win = the next 32 bits of the stream, MSB first; tab = byte offset of the table.
BJ_HD uint32_t lut_lookup(const LutMem &m, uint32_t tab, uint32_t win) {
    uint32_t e = m.ld(tab + ((win >> (32 - kRootBits)) << 2));
    if (__builtin_expect((int32_t)e < 0, 0)) e = m.ld(tab + (e & 0xFFFFu) + (((win << kRootBits) >> (32u - ((e >> 16) & 15u))) << 2));
    return e;
}

// ------------------------------------------------------------------------------------------------ scan bytes
// Classification of one raw scan byte from its neighbours (see the scan-byte walk cited above):
//   FF xx : FF kept iff xx == 00 (stuffed data byte); FF before FF is a fill byte; FF before RSTn is dropped
//   xx after FF : dropped (the stuffing zero or the marker code).  RSTn marks a segment boundary.
BJ_HD bool scan_keep(unsigned prev, unsigned b, unsigned next) { return b == 0xFFu ? next == 0x00u : prev != 0xFFu; }
BJ_HD bool scan_is_rst(unsigned prev, unsigned b) { return prev == 0xFFu && b >= 0xD0u && b <= 0xD7u; }

// The same rules on 16 bytes at once, four per 32-bit word (little-endian: byte i of the chunk sits in bits
// 8*(i&3).. of w[1 + i/4]); w[0] is the word before the chunk (only its top byte matters), w[5] the word after
// (only its low byte).  keep / rst: bit i = byte i survives / is the code byte of an RSTn marker.
BJ_HD uint32_t bytes_eq(uint32_t a, uint32_t b) {                          // 0xFF in every byte where a == b
#ifdef __CUDA_ARCH__
    return __vcmpeq4(a, b);
#else
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) if (((a >> (8 * i)) & 0xFFu) == ((b >> (8 * i)) & 0xFFu)) r |= 0xFFu << (8 * i);
    return r;
#endif
}
BJ_HD uint32_t movemask4(uint32_t m) { return (((m >> 7) & 0x01010101u) * 0x01020408u) >> 24; }   // bit 7 of each byte -> 4 bits
BJ_HD void classify_words(const uint32_t w[6], uint32_t &keep, uint32_t &rst) {
    keep = 0; rst = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int k = 0; k < 4; k++) {
        const uint32_t x = w[k + 1];
        const uint32_t before = (x << 8) | (w[k] >> 24);                   // byte i-1 under byte i
        const uint32_t behind = (x >> 8) | (w[k + 2] << 24);               // byte i+1 under byte i
        const uint32_t F = bytes_eq(x, 0xFFFFFFFFu), PF = bytes_eq(before, 0xFFFFFFFFu), Z = bytes_eq(behind, 0u);
        const uint32_t km = (F & Z) | (~F & ~PF);                          // scan_keep
        const uint32_t rm = PF & bytes_eq(x & 0xF8F8F8F8u, 0xD0D0D0D0u);   // scan_is_rst
        keep |= movemask4(km) << (4 * k);
        rst |= movemask4(rm) << (4 * k);
    }
}

// ------------------------------------------------------------------------------------------------ bit reader
// The un-stuffed stream is stored as 32-bit words whose most significant byte is the earliest byte (the
// un-stuff kernel writes byte o to address o ^ 3), so a window is two aligned words and one funnel shift.
BJ_HD uint32_t funnel_l(uint32_t hi, uint32_t lo, uint32_t s) {      // top 32 bits of (hi:lo) << (s mod 32)
#ifdef __CUDA_ARCH__
    return __funnelshift_l(lo, hi, s);
#else
    s &= 31u;
    return s ? (hi << s) | (lo >> (32 - s)) : hi;
#endif
}

// Decoder positions inside one span are kept as  S = (bit offset from `origin` << 8) | zig-zag index, `origin`
// being the 32-bit-aligned stream position at or before the span's first bit.  A step adds the low 16 bits of a
// table entry.  The zig-zag index never exceeds 127 (63 + 64), so bit 6 set means "unit complete" and the byte
// never carries into the position.  A symbol consumes at most 27 bits, so the word index grows by at most one.
constexpr uint32_t kWordS = 32u << 8;
struct BitStream {
    const uint32_t *w;            // -> word holding the current position
    uint32_t cur, nxt, nx2;       // that word and the two after it (the load runs one word ahead of its use)
    uint32_t word_end;            // S value of the first bit after `cur`
    BJ_HD static uint32_t ld(const uint32_t *a) {
#ifdef __CUDA_ARCH__
        return __ldg(a);
#else
        return *a;
#endif
    }
    BJ_HD void open(const uint32_t *words, uint32_t p) {
        w = words + (p >> 5);
        cur = ld(w); nxt = ld(w + 1); nx2 = ld(w + 2);
        word_end = kWordS;
    }
    BJ_HD void advance() { cur = nxt; nxt = nx2; w++; nx2 = ld(w + 2); word_end += kWordS; }
    BJ_HD uint32_t window(uint32_t S) {
        if (S >= word_end) advance();
        return funnel_l(cur, nxt, S >> 8);
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

// Per image: geometry of an MCU and where each component's tables sit.
struct HuffGeom {
    uint32_t bpm;           // data units per MCU
    uint32_t ny;            // luma units per MCU (hs * vs); unit c belongs to component c < ny ? 0 : c - ny + 1
    uint32_t tab[3];        // per component: DC table byte offset | AC table byte offset << 16 (from the staged base)
};
BJ_HD uint32_t comp_of(const HuffGeom &g, uint32_t c) { return c < g.ny ? 0u : c - g.ny + 1u; }
BJ_HD uint32_t tabs_of(const HuffGeom &g, uint32_t c) { return c < g.ny ? g.tab[0] : (c == g.ny ? g.tab[1] : g.tab[2]); }

// Magnitude extension of the `size` bits that follow a `len`-bit code in the window
// (src/jpeg_scanner.cpp:480-482 / :513-516: first bit 0 => negative).  size 0 gives 0.
BJ_HD int32_t extend_value(uint32_t win, uint32_t len, uint32_t size) {
    const uint32_t t = win << len;                                        // len <= 16
    const uint32_t raw = funnel_l(0u, t, size);                           // top `size` bits of t (size <= 11)
    return (int32_t)raw + (((int32_t)t >= 0) ? 1 - (int32_t)(1u << size) : 0);
}
// The same from a table entry; all shifts are wrap-mode (mod 32), fed by plain shifts of the entry.
BJ_HD int32_t extend_entry(uint32_t win, uint32_t e) {
    const uint32_t t = funnel_l(win, 0u, e >> 16);                        // win << len
    const uint32_t raw = funnel_l(0u, t, e >> 24);                        // top `size` bits of t
    const uint32_t m = funnel_l(0xFFFFFFFFu, 0u, e >> 24) + 1u;           // 1 - 2^size
    return (int32_t)(raw + (m & ~(uint32_t)((int32_t)t >> 31)));
}

// ------------------------------------------------------------------------------------------------ pass 1: synchronise
// Decode from `st` until the position reaches end_bit (first symbol boundary at or after it).  Only the decoder
// state is tracked - no values - plus the number of data units whose DC symbol starts inside the span (for the
// prefix sum that tells every sub-sequence which unit it starts in).  Errors do not stop a speculative decode
// (a wrong guess must not poison its successors): a refused symbol consumes its code bits (at least one) and the
// unit simply continues; an over-long run ends the unit.  With the true entry state the result is the true exit
// state up to the first real error, which the write pass detects and reports.
// The inner loop runs to the end of the current stream word or of the span, whichever comes first, so a step is:
// funnel shift, table load, add, unit-complete test, limit test.
BJ_HD HuffState decode_span(const uint32_t *words, const LutMem &luts, const HuffGeom &g, HuffState st, uint32_t end_bit,
                            uint32_t *units_started) {
    *units_started = 0;
    if (st.p >= end_bit) return st;
    const uint32_t origin = st.p & ~31u;
    uint32_t c = st.cz >> 8;
    uint32_t S = ((st.p - origin) << 8) | (st.cz & 0xFFu);
    const uint32_t endS = (end_bit - origin) << 8;
    const uint32_t entered_mid = (S & 0xFFu) ? 1u : 0u;
    uint32_t ends = 0;
    uint32_t tabs = tabs_of(g, c);
    uint32_t tab = entered_mid ? (tabs >> 16) : (tabs & 0xFFFFu);
    BitStream bs;
    bs.open(words, st.p);
    for (;;) {
        const uint32_t limit = endS < bs.word_end ? endS : bs.word_end;
        while (S < limit) {
            const uint32_t e = lut_lookup(luts, tab, funnel_l(bs.cur, bs.nxt, S >> 8));
            S += e & 0xFFFFu;
            tab = tabs >> 16;
            if (S & 0x40u) {                                              // zig-zag index >= 64: unit complete
                S &= ~0xFFu;
                ends++;
                c = (c + 1u == g.bpm) ? 0u : c + 1u;
                tabs = tabs_of(g, c);
                tab = tabs & 0xFFFFu;
            }
        }
        if (S >= endS) break;
        bs.advance();
    }
    // started = ended + (one still open at the exit) - (the one that was already open at the entry)
    *units_started = ends + ((S & 0xFFu) ? 1u : 0u) - entered_mid;
    HuffState o;
    o.p = origin + (S >> 8); o.cz = (c << 8) | (S & 0xFFu);
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
BJ_HD WriteResult write_span(const uint32_t *words, const LutMem &luts, const HuffGeom &g, HuffState st, uint32_t end_bit,
                             uint32_t data_end_bit, uint32_t du, uint32_t du_end, bool last_of_segment, Sink &sink) {
    WriteResult res;
    res.first_zero = 0xFFFFFFFFu;
    const uint32_t origin = st.p & ~31u;
    uint32_t c = st.cz >> 8;
    uint32_t S = ((st.p - origin) << 8) | (st.cz & 0xFFu);
    const uint32_t endS = end_bit > origin ? (end_bit - origin) << 8 : 0u;
    // S > dataS  <=>  the position is past the end of the segment's data: bits ran out inside the symbol
    // (BitReader::read_bit returns -1 in the reference)
    // (saturated: a position inside a span stays far below 2^24 bits from its origin)
    const uint32_t data_rel = data_end_bit >= origin ? data_end_bit - origin : 0u;
    const uint32_t dataS = ((data_rel < 0xFFFFFFu ? data_rel : 0xFFFFFFu) << 8) | 0xFFu;
    uint32_t tabs = tabs_of(g, c);
    BitStream bs;
    bs.open(words, st.p);
    bool stop = false;
    if (S & 0xFFu) {                                      // entered inside a unit that belongs to a predecessor: skip it
        for (;;) {
            if (S >= endS) { stop = true; break; }
            const uint32_t e = lut_lookup(luts, tabs >> 16, bs.window(S));
            S += e & 0xFFFFu;
            if (S & 0x40u) break;
        }
        if (!stop) {
            S &= ~0xFFu;
            c = (c + 1u == g.bpm) ? 0u : c + 1u;
            tabs = tabs_of(g, c);
        }
    }
    while (!stop) {
        // a unit starts here: mine if it starts before my end and the segment still has units to give
        if (S >= endS || du >= du_end) break;
        {
            const uint32_t win = bs.window(S);
            const uint32_t e = lut_lookup(luts, tabs & 0xFFFFu, win);
            const uint32_t Sn = S + (e & 0xFFFFu);
            // a failed DC leaves the unit untouched (zero)
            if ((e & kLutBad) || Sn > dataS) { res.first_zero = du; return res; }
            sink.dc(du, (int16_t)extend_entry(win, e));                   // |diff| < 2^11
            S = Sn;
        }
        const uint32_t tab = tabs >> 16;
        // AC symbols.  The reference's failure points: refused symbol, bits running out inside a symbol, run past
        // the end of the unit ("i + run >= 64", src/jpeg_scanner.cpp:497-500).  A failed AC keeps what was stored
        // before it.  The zig-zag index is >= 1 here, so an index above 64 means end-of-block or an over-long run.
        bool open = true, failed = false;
        while (open) {
            while (S < bs.word_end) {
                const uint32_t win = funnel_l(bs.cur, bs.nxt, S >> 8);
                const uint32_t e = lut_lookup(luts, tab, win);
                const uint32_t Sn = S + (e & 0xFFFFu);
                if (__builtin_expect((e & kLutBad) || Sn > dataS, 0)) { failed = true; open = false; break; }
                if (Sn & 0x40u) {                                         // index >= 64: the unit ends one way or another
                    const uint32_t zn = Sn & 0xFFu;
                    if (zn == 64u) {
                        const int32_t v = extend_entry(win, e);
                        if (v != 0) sink.put(63u, (int16_t)v);
                    } else if (!(e & kLutEob)) failed = true;
                    S = Sn;
                    open = false;
                    break;
                }
                const int32_t v = extend_entry(win, e);
                if (v != 0) sink.put((Sn & 0xFFu) - 1u, (int16_t)v);    // size 0 (ZRL and friends) would store a literal 0: already there
                S = Sn;
            }
            if (open) bs.advance();
        }
        sink.flush(du);
        if (failed) { res.first_zero = du + 1; return res; }
        du++;
        S &= ~0xFFu;
        c = (c + 1u == g.bpm) ? 0u : c + 1u;
        tabs = tabs_of(g, c);
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
