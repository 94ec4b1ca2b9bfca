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

// ------------------------------------------------------------------------------------------------ AC tables of the
// synchronisation pass.  That pass tracks only the decoder STATE (no values), so one lookup may take several
// symbols at once: every AC symbol (code + magnitude bits) that lies completely inside the kRootBits-bit window.
//   bits 15..0   the step for ALL of them:  consumed bits << 8 | zig-zag advance   (added to the packed state S)
//   bits 30..16  the same for the FIRST symbol alone
//   bit  31      link (root only), as above; second-level entries describe one symbol (both halves equal)
// End-of-block advances by 128 here (bit 7), a refused symbol is a plain step (the speculative decode never
// stops), and a group never contains an end-of-block and never advances by more than 63.  After a step from
// index z <= 63:   z' <= 63 the unit goes on;  z' == 64 its last coefficient was just read;  z' >= 128 end-of-block;
// 65..127 means a symbol INSIDE the group ended the unit (or the single symbol is an over-long run): the decoder
// takes the first symbol alone instead.
constexpr uint32_t kSyncEob = 128u;
inline uint32_t sync_half(int len, unsigned sym) {
    const unsigned run = sym >> 4, size = sym & 15u;
    const bool eob = sym == 0, bad = size > 10;
    return (eob ? kSyncEob : run + 1u) | ((uint32_t)(len + (bad ? 0u : size)) << 8);
}
// `write_lut` = the table build_lut made for the same Huffman table (same root/second-level layout).
inline void build_lut_sync(const uint32_t *write_lut, int used, uint32_t *lut) {
    const int R = kRootBits;
    auto half_of = [](uint32_t e) -> uint32_t {          // write-format leaf -> sync half
        const uint32_t adv = (e & kLutEob) ? kSyncEob : (e & 0xFFu);
        return adv | (e & 0xFF00u);
    };
    for (int i = 0; i < kLutCapAC; i++) lut[i] = 0;
    for (int i = 1 << R; i < used; i++) { const uint32_t h = half_of(write_lut[i]); lut[i] = h | (h << 16); }
    for (uint32_t w = 0; w < (1u << R); w++) {
        const uint32_t e = write_lut[w];
        if (e & kLutLink) { lut[w] = e; continue; }
        const uint32_t first = half_of(e);
        uint32_t bits = (first >> 8) & 0xFFu, adv = first & 0xFFu;
        // more symbols from the rest of the window, while they fit completely
        if (!(e & (kLutBad | kLutEob)) && bits <= (uint32_t)R) {
            for (;;) {
                const uint32_t rem = (uint32_t)R - bits;
                if (rem == 0) break;
                const uint32_t idx = (w << bits) & ((1u << R) - 1u);       // remaining bits, zero-padded
                const uint32_t n = write_lut[idx];
                if (n & (kLutLink | kLutBad | kLutEob)) break;
                const uint32_t nb = (n >> 8) & 0xFFu, na = n & 0xFFu;
                if (nb > rem || adv + na > 63u) break;
                bits += nb; adv += na;
            }
        }
        lut[w] = (adv | (bits << 8)) | (first << 16);
    }
}

// Where the staged tables live.  Table positions are BYTE addresses: on the device absolute shared-memory
// addresses (loads are ld.shared with no address arithmetic beyond the index), on the host offsets from `base`.
struct LutMem {
#ifdef __CUDA_ARCH__
    __device__ __forceinline__ uint32_t ld(uint32_t addr) const {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
        return v;
    }
#else
    const uint32_t *base;
    void attach(const uint32_t *tables) { base = tables; }
    uint32_t ld(uint32_t byte_off) const { return base[byte_off >> 2]; }
#endif
};

// Codes longer than the root (rare: ~0.4 % of symbols at q = 90).  Kept out of line on the device so that the hot
// loops carry a branch to it instead of its predicated body.
#ifdef __CUDA_ARCH__
__device__ __noinline__
#else
inline
#endif
uint32_t lut_second(uint32_t tab, uint32_t e, uint32_t win
#ifndef __CUDA_ARCH__
                    , const LutMem &m
#endif
) {
#ifdef __CUDA_ARCH__
    const LutMem m{};
#endif
    return m.ld(tab + (e & 0xFFFFu) + (((win << kRootBits) >> (32u - ((e >> 16) & 15u))) << 2));
}

// win = the next 32 bits of the stream, MSB first; tab = position of the table.
BJ_HD uint32_t lut_lookup(const LutMem &m, uint32_t tab, uint32_t win) {
    uint32_t e = m.ld(tab + ((win >> (32 - kRootBits)) << 2));
    if (__builtin_expect((int32_t)e < 0, 0)) {
#ifdef __CUDA_ARCH__
        e = lut_second(tab, e, win);
#else
        e = lut_second(tab, e, win, m);
#endif
    }
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
// Per-byte tests are exact bit tricks that flag a byte in its bit 7 (no carries cross a byte):
//   byte == FF  <=>  its low 7 bits + 1 carry into bit 7, and bit 7 is set
//   byte == 00  <=>  its low 7 bits + 7F do not reach bit 7, and bit 7 is clear
BJ_HD uint32_t flag_ff(uint32_t x) { return ((x & 0x7F7F7F7Fu) + 0x01010101u) & x & 0x80808080u; }
BJ_HD uint32_t flag_00(uint32_t x) { return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u; }
BJ_HD uint32_t movemask4(uint32_t m) { return (((m >> 7) & 0x01010101u) * 0x01020408u) >> 24; }   // bit 7 of each byte -> 4 bits
BJ_HD void classify_words(const uint32_t w[6], uint32_t &keep, uint32_t &rst) {
    keep = 0; rst = 0;
    uint32_t Fprev = flag_ff(w[0]);
    uint32_t Z = flag_00(w[1]);
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int k = 0; k < 4; k++) {
        const uint32_t x = w[k + 1];
        const uint32_t F = flag_ff(x), Znext = flag_00(w[k + 2]);
        const uint32_t PF = (F << 8) | (Fprev >> 24);                      // flag of byte i-1 under byte i
        const uint32_t ZN = (Z >> 8) | (Znext << 24);                      // flag of byte i+1 under byte i
        const uint32_t R = flag_00((x ^ 0xD0D0D0D0u) & 0xF8F8F8F8u);       // D0..D7
        const uint32_t km = (F & ZN) | (~(F | PF) & 0x80808080u);          // scan_keep
        const uint32_t rm = PF & R;                                        // scan_is_rst
        keep |= movemask4(km) << (4 * k);
        rst |= movemask4(rm) << (4 * k);
        Fprev = F; Z = Znext;
    }
}

// Where the entropy-coded segment ends (src/jpeg_scanner.cpp:405-433): at the first FF that is followed by something
// other than 00 (stuffing), FF (a fill byte) or RSTn.  EOI (D9) there is the regular end; any other marker makes the
// file invalid ("Invalid marker during compressed data scan").
BJ_HD bool scan_is_end(unsigned b, unsigned next) { return b == 0xFFu && next != 0x00u && next != 0xFFu && !(next >= 0xD0u && next <= 0xD7u); }
// The counting pass of K0 (which finds that FF on the device, so that the host never walks the scan): the same on 16
// bytes at once, plus  end: bit i = byte i is such an FF.
BJ_HD void classify_words_end(const uint32_t w[6], uint32_t &keep, uint32_t &rst, uint32_t &end) {
    keep = 0; rst = 0; end = 0;
    uint32_t Fprev = flag_ff(w[0]);
    uint32_t F = flag_ff(w[1]), Z = flag_00(w[1]), R = flag_00((w[1] ^ 0xD0D0D0D0u) & 0xF8F8F8F8u);
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int k = 0; k < 4; k++) {
        const uint32_t xn = w[k + 2];
        const uint32_t Fn = flag_ff(xn), Zn = flag_00(xn), Rn = flag_00((xn ^ 0xD0D0D0D0u) & 0xF8F8F8F8u);
        const uint32_t PF = (F << 8) | (Fprev >> 24);                      // flag of byte i-1 under byte i
        const uint32_t ZN = (Z >> 8) | (Zn << 24);                         // flags of byte i+1 under byte i
        const uint32_t FN = (F >> 8) | (Fn << 24);
        const uint32_t RN = (R >> 8) | (Rn << 24);
        const uint32_t km = (F & ZN) | (~(F | PF) & 0x80808080u);          // scan_keep
        const uint32_t rm = PF & R;                                        // scan_is_rst
        const uint32_t em = F & ~(ZN | FN | RN);                           // scan_is_end
        keep |= movemask4(km) << (4 * k);
        rst |= movemask4(rm) << (4 * k);
        end |= movemask4(em) << (4 * k);
        Fprev = F; F = Fn; Z = Zn; R = Rn;
    }
}

// A 16-byte chunk whose first byte has scan-relative index r0 (negative in front of the scan), clipped to the scan
// [0, raw_len): bytes outside do not count, and the FF that ends the scan must have its marker code inside, too.
BJ_HD void clip_chunk(int64_t r0, uint32_t raw_len, uint32_t &keep, uint32_t &rst, uint32_t &end) {
    const int64_t lo64 = -r0, hi64 = (int64_t)raw_len - r0;
    const uint32_t lo = lo64 <= 0 ? 0u : (lo64 >= 16 ? 16u : (uint32_t)lo64);
    const uint32_t hi = hi64 <= 0 ? 0u : (hi64 >= 16 ? 16u : (uint32_t)hi64);
    const uint32_t valid = ((1u << hi) - 1u) & ~((1u << lo) - 1u);
    keep &= valid; rst &= valid;
    end &= valid & (hi64 > 16 ? 0xFFFFu : (valid >> 1));
}
// The bytes of that chunk in front of scan-relative position e (the scan's end; kNoScanEnd: all of them).
constexpr uint32_t kNoScanEnd = 0xFFFFFFFFu;
BJ_HD uint32_t chunk_mask_before(int64_t r0, uint32_t e) {
    const int64_t lim = e == kNoScanEnd ? 16 : (int64_t)e - r0;
    return lim >= 16 ? 0xFFFFu : (lim <= 0 ? 0u : ((1u << (uint32_t)lim) - 1u));
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
#ifndef BJ_STREAM_LD
#define BJ_STREAM_LD "ld.global.nc.u32"
#endif
struct BitStream {
    const uint32_t *base;         // -> the word holding the span's origin
    uint32_t wi;                  // index (from base) of the word holding the current position
    uint32_t cur, nxt, nx2;       // that word and the two after it (the load runs one word ahead of its use)
    uint32_t word_end;            // S value of the first bit after `cur`
    BJ_HD void open(const uint32_t *words, uint32_t p) {
        base = words + (p >> 5);
        wi = 0u;
#ifdef __CUDA_ARCH__
        cur = __ldg(base); nxt = __ldg(base + 1); nx2 = __ldg(base + 2);
        asm volatile("mov.u64 %0, %0;" : "+l"(base));                       // a register pair, not an address rebuilt from the kernel's parameters in the symbol loops
#else
        cur = base[0]; nxt = base[1]; nx2 = base[2];
#endif
        word_end = kWordS;
    }
    // back to an earlier position of the same span (the write pass re-reads a unit that did not end well)
    BJ_HD void seek(uint32_t S) {
        wi = S >> 13;
#ifdef __CUDA_ARCH__
        cur = __ldg(base + wi); nxt = __ldg(base + wi + 1); nx2 = __ldg(base + wi + 2);
#else
        cur = base[wi]; nxt = base[wi + 1]; nx2 = base[wi + 2];
#endif
        word_end = (wi + 1u) << 13;
    }
    // the next 32 bits at S, moving on to the next word first if S has left the current one
    BJ_HD uint32_t window(uint32_t S) {
#ifdef __CUDA_ARCH__
        // predicated, no branch; the load writes nx2 directly (a select on the loaded value would stall on it).
        // The position is a 32-bit word index: its predicated increment is one instruction and the address of the word
        // the load would fetch one multiply-add (a 64-bit pointer moved by a 0-or-1 select came out of ptxas as four
        // instructions, a mad.wide inside the asm block likewise).
        const uint32_t *ahead = base + wi + 3;                             // (wi + 1) + 2: what becomes nx2 when the position moves on
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ge.u32 p, %6, %4;\n\t"
            "@p mov.b32 %0, %1;\n\t"
            "@p mov.b32 %1, %2;\n\t"
            "@p add.u32 %3, %3, 1;\n\t"
            "@p add.u32 %4, %4, 8192;\n\t"
            "@p " BJ_STREAM_LD " %2, [%5];\n\t}"
            : "+r"(cur), "+r"(nxt), "+r"(nx2), "+r"(wi), "+r"(word_end)
            : "l"(ahead), "r"(S));
#else
        if (S >= word_end) { cur = nxt; nxt = nx2; wi++; nx2 = base[wi + 2]; word_end += kWordS; }
#endif
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

// Per image: geometry of an MCU and where each component's tables sit (LutMem positions).
struct HuffGeom {
    uint32_t bpm;           // data units per MCU
    uint32_t ny;            // luma units per MCU (hs * vs); unit c belongs to component c < ny ? 0 : c - ny + 1
    uint32_t dc[3], ac[3];  // per component: DC / AC table
    uint32_t unit_tab;      // device: shared-memory address (256-byte aligned) of one uint4 per unit c of the MCU, about
                            // the unit c1 that FOLLOWS it: {DC table, AC table, UnitWalk step from c to c1, c1}
};
BJ_HD uint32_t comp_of(const HuffGeom &g, uint32_t c) { return c < g.ny ? 0u : c - g.ny + 1u; }
BJ_HD uint32_t dc_of(const HuffGeom &g, uint32_t c) { return c < g.ny ? g.dc[0] : (c == g.ny ? g.dc[1] : g.dc[2]); }
BJ_HD uint32_t ac_of(const HuffGeom &g, uint32_t c) { return c < g.ny ? g.ac[0] : (c == g.ny ? g.ac[1] : g.ac[2]); }
// The unit after unit c of the MCU: its index and both its tables - one 128-bit shared load on the device (the
// Huffman kernels always stage the table: stage_luts), selects on the host.
BJ_HD void next_unit(const HuffGeom &g, uint32_t c, uint32_t &c1, uint32_t &dc, uint32_t &ac) {
#ifdef __CUDA_ARCH__
    [[maybe_unused]] uint32_t step;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(dc), "=r"(ac), "=r"(step), "=r"(c1) : "r"(g.unit_tab + c * 16u));
#else
    c1 = (c + 1u == g.bpm) ? 0u : c + 1u;
    dc = dc_of(g, c1); ac = ac_of(g, c1);
#endif
}

// Where the synchronisation pass is: unit c of the MCU, and how many units it has completed.  On the device both
// live in one register, completed << 8 | c << 4, so that the hand-over to the next unit is one predicated add of
// the table's step (0x100 + 16 * (c1 - c)) and the table entry's address is one logic operation.
struct UnitWalk {
#ifdef __CUDA_ARCH__
    uint32_t cu;
    BJ_HD void start(uint32_t c) { cu = c << 4; }
    BJ_HD uint32_t unit() const { return (cu >> 4) & 15u; }
    BJ_HD uint32_t ended() const { return cu >> 8; }
    // `fin`: the unit is complete - move on to the next one.  tab = table of the next symbol, ac = AC table of the unit.
    BJ_HD void advance(const HuffGeom &g, bool fin, uint32_t &tab, uint32_t &ac) {
        uint32_t dcn, acn, step;
        [[maybe_unused]] uint32_t c1;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(dcn), "=r"(acn), "=r"(step), "=r"(c1) : "r"(g.unit_tab | (cu & 0xF0u)));
        cu += fin ? step : 0u;
        ac = fin ? acn : ac;
        tab = fin ? dcn : ac;
    }
#else
    uint32_t c, n;
    void start(uint32_t c0) { c = c0; n = 0; }
    uint32_t unit() const { return c; }
    uint32_t ended() const { return n; }
    void advance(const HuffGeom &g, bool fin, uint32_t &tab, uint32_t &ac) {
        uint32_t c1, dcn, acn;
        next_unit(g, c, c1, dcn, acn);
        c = fin ? c1 : c;
        n += fin ? 1u : 0u;
        ac = fin ? acn : ac;
        tab = fin ? dcn : ac;
    }
#endif
};
BJ_HD constexpr uint32_t unit_walk_step(uint32_t c, uint32_t c1) { return 0x100u + 16u * (c1 - c); }

// Magnitude extension of the `size` bits that follow a `len`-bit code in the window
// (src/jpeg_scanner.cpp:480-482 / :513-516: first bit 0 => negative).  size 0 gives 0.
BJ_HD int32_t extend_value(uint32_t win, uint32_t len, uint32_t size) {
    const uint32_t t = win << len;                                        // len <= 16
    const uint32_t raw = funnel_l(0u, t, size);                           // top `size` bits of t (size <= 11)
    return (int32_t)raw + (((int32_t)t >= 0) ? 1 - (int32_t)(1u << size) : 0);
}
// The same from a table entry; all shifts are wrap-mode (mod 32), fed by plain shifts of the entry.
BJ_HD int32_t extend_entry(uint32_t win, uint32_t e) {
    // a = -1 if the first magnitude bit is 1 (the value is the `size` bits as they are), 0 if it is 0 (the value is minus
    // the complemented bits):  x = top `size` bits of (a ? t : ~t),  value = a ? x : -x = (x ^ ~a) + a + 1
    const uint32_t t = funnel_l(win, 0u, e >> 16);                        // win << len
    const uint32_t a = (uint32_t)((int32_t)t >> 31);
#ifdef __CUDA_ARCH__
    uint32_t ta;
    asm("lop3.b32 %0, %1, %2, 0, 0xC3;" : "=r"(ta) : "r"(t), "r"(a));     // t ^ ~a in one instruction
#else
    const uint32_t ta = t ^ ~a;
#endif
    const uint32_t x = funnel_l(0u, ta, e >> 24);
    return (int32_t)((x ^ ~a) + a + 1u);
}

// ------------------------------------------------------------------------------------------------ pass 1: synchronise
// Decode from `st` until the position reaches end_bit (first step boundary at or after it).  Only the decoder
// state is tracked - no values - plus the number of data units whose DC symbol starts inside the span (for the
// prefix sum that tells every sub-sequence which unit it starts in).  Errors do not stop a speculative decode
// (a wrong guess must not poison its successors): a refused symbol consumes its code bits (at least one) and the
// unit simply continues; an over-long run ends the unit.  With the true entry state the result is the true exit
// state up to the first real error, which the write pass detects and reports.
//
// Slices.  The write pass works on finer pieces than this pass: a sub-sequence [start_bit, end_bit) is cut into
// slices of slice_bits, and whenever the decode crosses the start of slice k >= 1 it reports the state there
// (first step boundary at or after the slice's first bit) and the units started so far:  rec(k, p, cz, units).
// Every slice of the sub-sequence is reported exactly once per call, in order.  The AC tables are the grouped ones
// (build_lut_sync), so a "step" may be several symbols; they never span a unit boundary, so the units counted
// between two reported states are exactly the units whose DC symbol starts between them.
template <class Rec>
BJ_HD HuffState decode_span(const uint32_t *words, const LutMem &luts, const HuffGeom &g, HuffState st, uint32_t start_bit,
                            uint32_t end_bit, uint32_t slice_bits, Rec &rec, uint32_t *units_started) {
    const uint32_t origin = st.p & ~31u;
    UnitWalk u;
    u.start(st.cz >> 8);
    uint32_t S = ((st.p - origin) << 8) | (st.cz & 0xFFu);
    const uint32_t endS = end_bit > origin ? (end_bit - origin) << 8 : 0u;
    const uint32_t entered_mid = (S & 0xFFu) ? 1u : 0u;
    const uint32_t nslices = end_bit > start_bit ? (end_bit - start_bit + slice_bits - 1u) / slice_bits : 1u;
    uint32_t k = 1;
    uint32_t ckS = start_bit + slice_bits > origin ? (start_bit + slice_bits - origin) << 8 : 0u;   // slice k starts here
    uint32_t limS = (k < nslices && ckS < endS) ? ckS : endS;
    uint32_t ac = ac_of(g, st.cz >> 8);
    uint32_t tab = entered_mid ? ac : dc_of(g, st.cz >> 8);
    BitStream bs;
    bs.open(words, st.p);
    for (;;) {
        while (S < limS) {
            const uint32_t e = lut_lookup(luts, tab, bs.window(S));
            uint32_t Sn = S + (e & 0xFFFFu);
            // 65..127: a symbol inside the group ended the unit (or an over-long run): the first symbol alone
            if (__builtin_expect(((Sn & 0xFFu) - 65u) < 63u, 0)) Sn = S + ((e >> 16) & 0x7FFFu);
            // zig-zag index >= 64: unit complete (also: over-long run).  Branch-free: most iterations of a warp see one.
            const bool fin = (Sn & 0xC0u) != 0u;
            u.advance(g, fin, tab, ac);
            S = fin ? (Sn & ~0xFFu) : Sn;
        }
        while (k < nslices && S >= ckS) {                                 // once per slice
            rec(k, origin + (S >> 8), (u.unit() << 8) | (S & 0xFFu), u.ended() + ((S & 0xFFu) ? 1u : 0u) - entered_mid);
            k++;
            ckS += slice_bits << 8;
        }
        if (S >= endS) break;
        limS = (k < nslices && ckS < endS) ? ckS : endS;
    }
    // started = ended + (one still open at the exit) - (the one that was already open at the entry)
    *units_started = u.ended() + ((S & 0xFFu) ? 1u : 0u) - entered_mid;
    HuffState o;
    o.p = origin + (S >> 8); o.cz = (u.unit() << 8) | (S & 0xFFu);
    return o;
}

// ------------------------------------------------------------------------------------------------ pass 2: write
// Ownership rule: a data unit belongs to the slice in which its DC symbol starts; the owner decodes the whole
// unit (running past its own end if necessary) and stores all 64 coefficients at once, so every unit is written
// exactly once, by one thread, with no zero-fill pass.  A slice entered mid-unit first skips to the end of that
// unit without storing anything.  DC DIFFERENCES go to a separate compact plane (2 bytes per unit); the
// prediction sums over it are a separate, tiny scan (K1c) and slot 0 of every unit stays zero.
//
// The decode is a cursor that takes ONE symbol per step(), DC and AC alike, so the lanes of a warp stay together;
// a step says when a unit is complete and the caller stores it (the kernel does that warp-cooperatively).
// Sink concept:  void put(uint32_t zz, int16_t v)  - stage one non-zero value of the current unit (zz = 0: the DC
// difference; the stored unit's slot 0 is zero, the difference goes to the DC plane);  void reset()  - forget what
// has been staged for the current unit (it is decoded again: redo_unit).
constexpr uint32_t kEvDone = 2u;     // open(): nothing to do in this slice

struct WriteCursor {
    BitStream bs;
    uint32_t S, c, ac, tab, du;     // ac: AC table of the current unit; tab: table of the next symbol
    uint32_t S0, bad;               // where the current unit starts; OR of its symbols' table entries (kLutBad is sticky)
    uint32_t endS, dataS, du_end;
    uint32_t first_zero;    // first unit index that must read as zero because the reference stopped; UINT32_MAX if none
    uint32_t fail;          // 0, or why the slice stopped: 1 = refused symbol / bits ran out, 2 = over-long run
    uint32_t done;          // the slice is finished (set by the step that completes its last unit, or by a failure)
    uint32_t st_du;         // after a step that returned true: the unit to store from the stage (UINT32_MAX: none)
    uint32_t idle;          // device: position of an all-zero word.  A finished lane keeps stepping with its warp on a
                            // zero window and this "table": the entry is 0, the position does not move, nothing is staged

    // Returns kEvDone if there is nothing to do.  Skips the unit a predecessor owns (no values).
    BJ_HD uint32_t open(const uint32_t *words, const LutMem &luts, const HuffGeom &g, HuffState st, uint32_t end_bit,
                        uint32_t data_end_bit, uint32_t du0, uint32_t du_end_) {
        first_zero = 0xFFFFFFFFu;
        const uint32_t origin = st.p & ~31u;
        c = st.cz >> 8;
        S = ((st.p - origin) << 8) | (st.cz & 0xFFu);
        endS = end_bit > origin ? (end_bit - origin) << 8 : 0u;
        // S > dataS  <=>  the position is past the end of the segment's data: bits ran out inside the symbol
        // (BitReader::read_bit returns -1 in the reference).  Saturated: positions stay far below 2^24 bits.
        const uint32_t data_rel = data_end_bit >= origin ? data_end_bit - origin : 0u;
        dataS = ((data_rel < 0xFFFFFFu ? data_rel : 0xFFFFFFu) << 8) | 0xFFu;
        du = du0; du_end = du_end_;
        fail = 0u; done = 0u; st_du = 0xFFFFFFFFu;
        ac = ac_of(g, c);
        bs.open(words, st.p);
        if (S & 0xFFu) {                                  // inside a unit that belongs to a predecessor: skip it
            for (;;) {
                if (S >= endS) return kEvDone;
                const uint32_t e = lut_lookup(luts, ac, bs.window(S));
                S += e & 0xFFFFu;
                if (S & 0x40u) break;
            }
            S &= ~0xFFu;
            c = (c + 1u == g.bpm) ? 0u : c + 1u;
            ac = ac_of(g, c);
        }
        tab = dc_of(g, c);
        S0 = S; bad = 0u;
        // a unit starts here: mine if it starts before my end and the segment still has units to give
        return (S >= endS || du >= du_end) ? kEvDone : 0u;
    }

    // One symbol.  Returns true when there is something for the caller to do: the unit st_du is complete - store the
    // staged unit (slot 0 holds its DC difference) and clear the stage; st_du = UINT32_MAX: nothing to store.  `done`
    // is only ever set by a step that returns true.
    //
    // The reference's failure points: refused symbol, bits running out inside a symbol, run past the end of the unit
    // ("i + run >= 64", src/jpeg_scanner.cpp:497-500).  A failed DC leaves the unit untouched (zero); a failed AC
    // keeps what was stored before it; first_zero = the first unit the reference never reached.
    //
    // The symbol loop does not look for failures: it decodes on (a refused code is a table entry like any other,
    // with no value; a coefficient index past 63 wraps inside the staged unit) and looks back when the unit ends -
    // did any symbol carry kLutBad, does the unit end past the segment's data, did it end on an over-long run.
    // A unit that did not end well is decoded again from its first symbol by redo_unit(), which stops exactly where
    // the reference stops.  (Damaged data only; at most 63 further symbols are read before the unit ends.)
    template <class Sink>
    BJ_HD bool step(const LutMem &luts, const HuffGeom &g, Sink &sink) {
        if (!step_plain(luts, sink)) return false;
        unit_end(luts, g, sink);
        return true;
    }
    // The same in two halves, for a caller that lets a lane wait between the symbol that completes a unit and the
    // hand-over to the next unit (k_huff_write takes several symbols per round and hands over once per round, all lanes
    // that completed a unit together).  step_plain: one symbol, no checks; true = the unit is complete: call unit_end()
    // before the next symbol.  unit_end: the look back over the unit (see above), then the hand-over - or the unit again.
    template <class Sink>
    BJ_HD bool step_plain(const LutMem &luts, Sink &sink) {
        const uint32_t win = bs.window(S);
        const uint32_t e = lut_lookup(luts, tab, win);
        S += e & 0xFFFFu;
        bad |= e;
        const int32_t v = extend_entry(win, e);
        // index of the coefficient this symbol carries (0: the DC difference).  size 0 (ZRL, EOB, refused) stores nothing.
        if (v != 0) sink.put(((S & 0xFFu) - 1u) & 63u, (int16_t)v);
        tab = ac;
        return (S & 0x40u) != 0u;                         // index >= 64: the unit ends one way or another
    }
    template <class Sink>
    BJ_HD void unit_end(const LutMem &luts, const HuffGeom &g, Sink &sink) {
        // (kLutEob in `bad` can only come from the unit's last symbol: an end-of-block always ends the unit)
        if (__builtin_expect((bad & kLutBad) || S > dataS || ((S & 0xFFu) != 64u && !(bad & kLutEob)), 0)) redo_unit(luts, g, sink);
        else unit_done(g);
    }
    BJ_HD void unit_done(const HuffGeom &g) {
        st_du = du++;
        S &= ~0xFFu;
        next_unit(g, c, c, tab, ac);
        S0 = S; bad = 0u;
        if (S >= endS || du >= du_end) finish();
    }
    BJ_HD void finish() {
        done = 1u;
#ifdef __CUDA_ARCH__
        S = 0u; bs.cur = 0u; bs.nxt = 0u; tab = idle; ac = idle;     // (S stays below bs.word_end: no further loads)
#endif
    }
    BJ_HD void failed(uint32_t why) {
        fail = why;
        const bool dc = why == 1u && (S & 0xFFu) == 0u;
        first_zero = dc ? du : du + 1u;                   // a failed DC: the unit stays zero; a failed AC keeps what was stored before it
        st_du = dc ? 0xFFFFFFFFu : du;
        finish();
    }
    // The current unit again, from its first symbol, symbol by symbol with the reference's checks.
    template <class Sink>
    BJ_HD void redo_unit(const LutMem &luts, const HuffGeom &g, Sink &sink) {
        sink.reset();
        S = S0;
        bs.seek(S0);
        tab = dc_of(g, c);
        for (;;) {
            const uint32_t win = bs.window(S);
            const uint32_t e = lut_lookup(luts, tab, win);
            const uint32_t Sn = S + (e & 0xFFFFu);
            if ((e & kLutBad) || Sn > dataS) { failed(1u); return; }
            const int32_t v = extend_entry(win, e);
            const uint32_t zz = (Sn & 0xFFu) - 1u;
            if (v != 0 && zz < 64u) sink.put(zz, (int16_t)v);
            tab = ac;
            S = Sn;
            if (!(Sn & 0x40u)) continue;
            if ((Sn & 0xFFu) != 64u && !(e & kLutEob)) { failed(2u); return; }   // over-long run
            break;                                        // (not reached: a unit is only decoded again if one of the checks fires)
        }
        unit_done(g);
    }
};

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
