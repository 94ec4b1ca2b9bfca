// TEST INFRASTRUCTURE - host-side emulation of the entropy-stage kernels' data flow.
//
// Compiles the product's device/host-shared core (pim_jpeg_decoder_b200/csrc/huff_core.h, parse.h) with g++ and
// runs the same algorithm the CUDA kernels run - un-stuff + segment table, speculative sub-sequence decode,
// fix-up rounds until a fixed point, per-segment prefix sums, owner-writes-whole-unit pass - sequentially on the
// CPU, so the algorithm can be checked against the oracle without a GPU.  It is NOT linked into libb200jpeg.so
// and is only loaded by tests/test_huff_emu.py.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../pim_jpeg_decoder_b200/csrc/huff_core.h"
#include "../../pim_jpeg_decoder_b200/csrc/parse.h"

using namespace bj;

namespace {
struct HostSink {
    int16_t unit[64];
    HostSink() { memset(unit, 0, sizeof(unit)); }
    void put(uint32_t zz, int16_t v) { unit[zz] = v; }
    void reset() { memset(unit, 0, sizeof(unit)); }
};
struct SliceRec { uint32_t p, cz, cnt; };
}  // namespace


// ------------------------------------------------------------------------------------------------ K0, as the kernels run it
// The device buffer holds the file at `file_pos` (any alignment; what lies around it is arbitrary).  Tiles of
// kTile raw bytes, 256 "threads" of four consecutive aligned 16-byte chunks each, as k_unstuff works: per tile, where the scan ends
// (if in this tile), the bytes that survive and the RSTn markers in front of that point; what the earlier tiles of the
// image contribute (the kernel's look-back; here a running prefix, tiles in order); compaction (byte o stored at
// o ^ 3), segment starts, the image's state.  Same shared code as the kernel: classify_words_end, clip_chunk,
// chunk_mask_before.
namespace {
constexpr uint32_t kTile = 16384;
struct K0Out {
    uint32_t status = 0;         // 0 ok, 2 invalid (no EOI where the scan ends)
    uint32_t raw_len = 0, end_code = 0x100, clean_len = 0, nrst = 0;
    std::vector<uint8_t> clean;  // swizzled (o ^ 3), zero padded
    std::vector<uint32_t> seg_off;   // starts of the segments found (without the closing entry)
};
struct Chunk { uint32_t keep, rst, end; uint32_t w[4]; int64_t r0; };

// chunk q of the tile: thread q / 4 (a thread holds four consecutive chunks), its chunk q % 4
Chunk classify_chunk(const std::vector<uint8_t> &buf, uint64_t raw_off, uint32_t tile, uint32_t q, uint32_t raw_len) {
    Chunk c;
    const uint64_t a0 = (raw_off & ~(uint64_t)15) + (uint64_t)tile * kTile + (uint64_t)q * 16;
    c.r0 = (int64_t)a0 - (int64_t)raw_off;
    auto live_at = [&](int64_t r) { return r < (int64_t)raw_len && r + 16 > 0; };
    const bool live = live_at(c.r0);
    auto at = [&](uint64_t a) -> uint32_t { return a < buf.size() ? buf[a] : 0xEEu; };
    uint32_t w[6] = {0, 0, 0, 0, 0, 0};
    if (live) for (int b = 0; b < 16; b++) w[1 + b / 4] |= at(a0 + b) << (8 * (b & 3));
    // the neighbours' edge bytes, exactly as k_unstuff gets them: inside a thread and between the lanes of a warp from
    // registers (a chunk that is not live holds zeros), at the warp's edges from memory
    const uint32_t thread = q / 4, sub = q % 4, lane = thread % 32;
    if (sub == 0 && lane == 0) w[0] = (live && a0 > 0) ? at(a0 - 1) << 24 : 0u;
    else w[0] = live_at(c.r0 - 16) ? at(a0 - 1) << 24 : 0u;
    if (sub == 3 && lane == 31) w[5] = live ? at(a0 + 16) : 0u;
    else w[5] = live_at(c.r0 + 16) ? at(a0 + 16) : 0u;
    for (int k = 0; k < 4; k++) c.w[k] = w[k + 1];
    classify_words_end(w, c.keep, c.rst, c.end);
    clip_chunk(c.r0, raw_len, c.keep, c.rst, c.end);
    return c;
}

void k0_emulate(const std::vector<uint8_t> &buf, uint64_t raw_off, uint32_t raw_len_max, uint32_t nseg_expected, K0Out *o) {
    const uint32_t mis = (uint32_t)(raw_off & 15u);
    const uint32_t ntile = std::max<uint32_t>(1u, (uint32_t)((mis + (uint64_t)raw_len_max + kTile - 1) / kTile));
    o->clean.assign(((size_t)raw_len_max / 4 + 4 + 96) * 4, 0);
    o->seg_off.assign(nseg_expected, 0xFFFFFFFFu);
    // what the look-back hands a tile: surviving bytes / markers of the image's earlier tiles, and whether the scan has ended
    uint32_t pk = 0, pr = 0;
    bool ended = false, state_written = false;
    for (uint32_t t = 0; t < ntile; t++) {
        uint32_t e = kNoScanEnd;
        std::vector<Chunk> cs;
        for (uint32_t th = 0; th < kTile / 16; th++) {
            cs.push_back(classify_chunk(buf, raw_off, t, th, raw_len_max));
            if (cs.back().end) e = std::min(e, (uint32_t)(cs.back().r0 + __builtin_ctz(cs.back().end)));
        }
        uint32_t ta = 0, tb = 0;
        for (auto &c : cs) { const uint32_t m = chunk_mask_before(c.r0, e); c.keep &= m; c.rst &= m; ta += __builtin_popcount(c.keep); tb += __builtin_popcount(c.rst); }
        const uint32_t base_k = pk, base_r = pr;
        const bool dead = ended;
        pk += ta; pr += tb; ended = ended || e != kNoScanEnd;              // (the tile's inclusive word)
        if (dead) continue;
        if (e != kNoScanEnd || t + 1 == ntile) {                           // the image's state
            if (state_written) { o->status = 99; return; }
            state_written = true;
            o->raw_len = e == kNoScanEnd ? raw_len_max : e;
            o->end_code = e == kNoScanEnd ? 0x100u : buf[raw_off + e + 1];
            o->status = o->end_code != 0xD9u ? 2u : 0u;
            o->clean_len = base_k + ta; o->nrst = base_r + tb;
            o->seg_off[0] = 0;
        }
        uint32_t pos = base_k, sidx = base_r + 1;
        for (auto &c : cs)
            for (int i = 0; i < 16; i++) {
                if (c.keep & (1u << i)) { o->clean[pos ^ 3u] = (uint8_t)(c.w[i >> 2] >> ((i & 3) * 8)); pos++; }
                else if (c.rst & (1u << i)) { if (sidx < nseg_expected) o->seg_off[sidx] = pos; sidx++; }
            }
    }
    if (!state_written) { o->status = 98; return; }
    o->seg_off.resize(o->status ? 1 : std::min(o->nrst + 1u, nseg_expected));
}
}  // namespace

// K0 alone against the per-byte rules and the host's scan walk (find_scan_end): the file sits `file_pos` bytes into a
// device buffer filled with `fill` (what a neighbouring file or stale memory may hold).  Returns 0 when status, true
// scan length, surviving bytes, marker count and segment starts all agree; a positive code says what differs.
extern "C" int emu_k0_check(const uint8_t *file, size_t len, int file_pos, int fill, int nseg_expected) {
    bj_image_desc d;
    const int rc_hdr = parse_header(file, len, &d, /*walk_scan=*/false);
    if (rc_hdr == BJ_ERR_INVALID_JPEG) { bj_image_desc d2; return parse_header(file, len, &d2, true) != BJ_OK ? 0 : 7; }   // rejected on the header alone
    if (rc_hdr != BJ_OK && rc_hdr != BJ_ERR_UNSUPPORTED) return rc_hdr;
    std::vector<uint8_t> buf((size_t)file_pos + len + 8192, (uint8_t)fill);
    memcpy(buf.data() + file_pos, file, len);
    K0Out o;
    k0_emulate(buf, (uint64_t)file_pos + d.scan_off, (uint32_t)d.scan_len, (uint32_t)nseg_expected, &o);
    size_t end = 0;
    const int rc_walk = find_scan_end(file, len, d.scan_off, &end);
    if ((rc_walk != BJ_OK) != (o.status != 0)) return 1;
    if (rc_walk != BJ_OK) return 0;                               // both reject the file
    if (o.raw_len != end - d.scan_off) return 2;
    const uint8_t *raw = file + d.scan_off;
    std::vector<uint8_t> want;
    std::vector<uint32_t> seg(1, 0);
    uint32_t nrst = 0;
    for (size_t i = 0; i < o.raw_len; i++) {
        const unsigned prev = i ? raw[i - 1] : 0u, bb = raw[i], next = raw[i + 1];
        if (scan_keep(prev, bb, next)) want.push_back((uint8_t)bb);
        else if (scan_is_rst(prev, bb)) { nrst++; if (seg.size() < (size_t)nseg_expected) seg.push_back((uint32_t)want.size()); }
    }
    if (o.clean_len != want.size() || o.nrst != nrst) return 3;
    for (size_t i = 0; i < want.size(); i++) if (o.clean[i ^ 3] != want[i]) return 4;
    if (o.seg_off.size() != std::min<size_t>(nrst + 1, (size_t)nseg_expected)) return 5;
    for (size_t i = 0; i < o.seg_off.size(); i++) if (o.seg_off[i] != seg[i]) return 6;
    return 0;
}

// optional: per Jacobi round, how many sub-sequences were decoded (set by emu_set_round_hist; 64 entries)
static uint32_t *g_round_hist = nullptr;
extern "C" void emu_set_round_hist(uint32_t *hist) { g_round_hist = hist; }
// 0: the write pass one symbol per step, slice after slice; n > 0: warps of 32 slices in rounds of n symbols, as k_huff_write
static int g_write_rounds = 0;
extern "C" void emu_set_write_rounds(int n) { g_write_rounds = n; }

// slice_bytes = granularity of the write pass; a sub-sequence of the synchronisation pass is `slices` of them.
// Returns 0 ok, <0 parse status.  info[0] = fix-up rounds, info[1] = first_zero (UINT32_MAX none), info[2] = nsub,
// info[3] = number of units written more or less than once (must be 0 for a clean stream)
extern "C" int emu_entropy(const uint8_t *file, size_t len, int slice_bytes, int slices, int16_t *coef_zz, uint32_t *info) {
    bj_image_desc d;
    int rc = parse_header(file, len, &d, /*walk_scan=*/false);    // the decode path's parse: headers only
    if (rc != BJ_OK) return rc;
    const uint32_t sub_bytes = (uint32_t)slice_bytes * (uint32_t)slices, slice_bits = (uint32_t)slice_bytes * 8u;
    const uint32_t nmx = (d.mcu_w + d.hs - 1) / d.hs, nmy = (d.mcu_h + d.vs - 1) / d.vs, nmcu = nmx * nmy;
    uint32_t bpm = 0;
    for (int j = 0; j < d.ncomp; j++) bpm += d.comp_h[j] * d.comp_v[j];
    const uint32_t ndu = nmcu * bpm;
    const uint32_t ri = d.restart_interval;
    const uint32_t nseg_expected = ri ? (nmcu + ri - 1) / ri : 1;

    // K0 (count, scan, compact) on a device-buffer image of the file at an odd alignment
    K0Out k0;
    {
        const size_t file_pos = 16 + (len % 13);
        std::vector<uint8_t> buf(file_pos + len + 8192, 0xFF);
        memcpy(buf.data() + file_pos, file, len);
        k0_emulate(buf, file_pos + d.scan_off, (uint32_t)d.scan_len, nseg_expected, &k0);
    }
    if (k0.status) return BJ_ERR_INVALID_JPEG;
    std::vector<uint8_t> &clean = k0.clean;
    std::vector<uint32_t> seg_off = k0.seg_off;
    for (uint32_t sg : seg_off) if (sg == 0xFFFFFFFFu) return -203;
    const size_t o = k0.clean_len;
    const uint32_t clean_len = (uint32_t)o;
    const uint32_t nseg = (uint32_t)seg_off.size();
    seg_off.push_back(clean_len);
    uint32_t first_zero = nseg < nseg_expected ? nseg * ri * bpm : 0xFFFFFFFFu;

    // tables: the write pass' single-symbol tables, and the synchronisation pass' grouped AC tables
    const size_t lut_words = 3 * (kLutCapDC + kLutCapAC);
    std::vector<uint32_t> luts(lut_words), luts_sync(lut_words);
    HuffGeom g;
    g.bpm = bpm; g.ny = (uint32_t)d.hs * d.vs; g.unit_tab = 0;
    for (int j = 0; j < 3; j++) {
        const int jj = j < d.ncomp ? j : 0;
        const uint32_t odc = j * kLutCapDC, oac = 3 * kLutCapDC + j * kLutCapAC;
        if (build_lut(d.dc_offsets[d.dc_id[jj]], d.dc_symbols[d.dc_id[jj]], false, &luts[odc]) < 0) return -100;
        const int used = build_lut(d.ac_offsets[d.ac_id[jj]], d.ac_symbols[d.ac_id[jj]], true, &luts[oac]);
        if (used < 0) return -100;
        memcpy(&luts_sync[odc], &luts[odc], kLutCapDC * 4);
        build_lut_sync(&luts[oac], used, &luts_sync[oac]);
        g.dc[j] = odc * 4; g.ac[j] = oac * 4;
    }
    LutMem lm, lm_sync;
    lm.attach(luts.data());
    lm_sync.attach(luts_sync.data());

    // sub-sequence table
    struct Sub { uint32_t seg, start_bit, end_bit; bool head, last; };
    std::vector<Sub> subs;
    for (uint32_t s = 0; s < nseg; s++) {
        const uint32_t b0 = seg_off[s], b1 = seg_off[s + 1];
        const uint32_t ns = std::max<uint32_t>(1, (b1 - b0 + sub_bytes - 1) / sub_bytes);
        for (uint32_t k = 0; k < ns; k++) {
            Sub u;
            u.seg = s; u.head = k == 0; u.last = k + 1 == ns;
            u.start_bit = (b0 + k * sub_bytes) * 8;
            u.end_bit = std::min<uint32_t>(b0 + (k + 1) * sub_bytes, b1) * 8;
            subs.push_back(u);
        }
    }
    const size_t ns = subs.size();
    const uint32_t *words = reinterpret_cast<const uint32_t *>(clean.data());

    // pass 1: Jacobi rounds to the fixed point  in[i+1] == out[i]; every decode (re)writes the sub-sequence's slices
    std::vector<HuffState> in(ns), out(ns);
    std::vector<uint32_t> tot(ns);
    std::vector<uint8_t> need(ns, 1);
    std::vector<SliceRec> slice(ns * (size_t)slices);
    std::vector<uint8_t> slice_seen(ns * (size_t)slices, 0);
    for (size_t i = 0; i < ns; i++) { in[i].p = subs[i].start_bit; in[i].cz = 0; }
    uint32_t rounds = 0;
    int order_bad = 0;
    for (;;) {
        bool any = false;
        if (g_round_hist && rounds < 64) { uint32_t cnt = 0; for (size_t i = 0; i < ns; i++) cnt += need[i]; g_round_hist[rounds] = cnt; }
        for (size_t i = 0; i < ns; i++) {
            if (!need[i]) continue;
            uint32_t expect = 1;
            auto rec = [&](uint32_t k, uint32_t p, uint32_t cz, uint32_t cnt) {
                if (k != expect || k >= (uint32_t)slices) order_bad++;
                expect = k + 1;
                slice[i * slices + k] = SliceRec{p, cz, cnt};
                slice_seen[i * slices + k] = 1;
            };
            out[i] = decode_span(words, lm_sync, g, in[i], subs[i].start_bit, subs[i].end_bit, slice_bits, rec, &tot[i]);
            const uint32_t nsl = subs[i].end_bit > subs[i].start_bit ? (subs[i].end_bit - subs[i].start_bit + slice_bits - 1) / slice_bits : 1u;
            if (expect != nsl) order_bad++;
            need[i] = 0; any = true;
        }
        if (!any) break;
        rounds++;
        for (size_t i = ns; i-- > 1;) {
            if (subs[i].head) continue;
            if (!same_state(in[i], out[i - 1])) { in[i] = out[i - 1]; need[i] = 1; }
        }
    }
    if (order_bad) return -300;

    // prefix sum of the units started per segment, then the write pass (DC differences into their own plane).
    // g_write_rounds == 0: one cursor per slice, one symbol per step, the finished unit stored when the step says so.
    // g_write_rounds  > 0: as k_huff_write runs it - "warps" of 32 consecutive slices work in rounds: every lane takes up
    // to g_write_rounds symbols without any check (step_plain) and stops at the one that completes its unit; then the
    // lanes that completed a unit look back over it and hand over (unit_end), announce (lane, unit) in the warp's slot
    // array in lane order, and the units are stored from the lanes' stages in slot order; a lane that has finished sits
    // the rounds out.  Every stage is cleared when its unit has been stored; a unit is stored at most once.
    std::vector<uint8_t> written(ndu, 0);
    std::vector<int16_t> dcp(ndu, 0x5A5A);
    struct Lane { WriteCursor cur; HostSink sink; bool done, active, last; };
    std::vector<Lane> lanes;
    uint32_t n_ex = 0;
    for (size_t i = 0; i < ns; i++) {
        const Sub &u = subs[i];
        if (u.head) n_ex = 0;
        const uint32_t du0 = u.seg * ri * bpm;
        const uint32_t du_end = (ri ? std::min(nmcu, (u.seg + 1) * ri) : nmcu) * bpm;
        const uint32_t nsl = u.end_bit > u.start_bit ? (u.end_bit - u.start_bit + slice_bits - 1) / slice_bits : 1u;
        for (uint32_t k = 0; k < nsl; k++) {
            HuffState st = in[i];
            uint32_t cnt = 0;
            if (k) { st.p = slice[i * slices + k].p; st.cz = slice[i * slices + k].cz; cnt = slice[i * slices + k].cnt; }
            const uint32_t end_bit = std::min(u.start_bit + (k + 1) * slice_bits, u.end_bit);
            lanes.emplace_back();
            Lane &L = lanes.back();
            L.cur.idle = 0;
            L.active = true;
            L.last = u.last && k + 1 == nsl;
            L.done = (L.cur.open(words, lm, g, st, end_bit, seg_off[u.seg + 1] * 8, du0 + n_ex + cnt, du_end) & kEvDone) != 0;
        }
        n_ex += tot[i];
    }
    auto store_unit = [&](Lane &L, uint32_t du) {
        if (du < ndu) {
            dcp[du] = L.sink.unit[0];
            L.sink.unit[0] = 0;
            memcpy(coef_zz + (size_t)du * 64, L.sink.unit, sizeof(L.sink.unit));
            written[du]++;
        }
        memset(L.sink.unit, 0, sizeof(L.sink.unit));
    };
    if (g_write_rounds <= 0) {
        for (Lane &L : lanes) {
            while (!L.done) {
                const bool unit = L.cur.step(lm, g, L.sink);
                L.done = L.cur.done != 0;
                if (unit) store_unit(L, L.cur.st_du);
            }
        }
    } else {
        for (size_t w0 = 0; w0 < lanes.size(); w0 += 32) {
            const size_t w1 = std::min(lanes.size(), w0 + 32);
            for (;;) {
                bool all_done = true;
                for (size_t l = w0; l < w1; l++) all_done = all_done && lanes[l].done;
                if (all_done) break;
                struct Slot { uint32_t lane, du; } slot[32];
                uint32_t nun = 0;
                bool fin[32] = {};
                for (size_t l = w0; l < w1; l++) {                           // the symbols of the round
                    Lane &L = lanes[l];
                    if (L.done) continue;
                    for (int k = 0; k < g_write_rounds && !(L.cur.S & 0x40u); k++) L.cur.step_plain(lm, L.sink);
                    fin[l - w0] = (L.cur.S & 0x40u) != 0u;
                }
                for (size_t l = w0; l < w1; l++) {                           // look-back + hand-over, slots in lane order
                    if (!fin[l - w0]) continue;
                    Lane &L = lanes[l];
                    L.cur.unit_end(lm, g, L.sink);
                    L.done = L.cur.done != 0;
                    slot[nun].lane = (uint32_t)(l - w0); slot[nun].du = L.cur.st_du; nun++;
                }
                for (uint32_t k = 0; k < nun; k++) store_unit(lanes[w0 + slot[k].lane], slot[k].du);   // the flush
            }
        }
    }
    for (Lane &L : lanes) {
        // the last slice of a segment must have produced the segment's last unit
        if (L.cur.fail == 0 && L.last && L.cur.du < L.cur.du_end) L.cur.first_zero = L.cur.du;
        first_zero = std::min(first_zero, L.cur.first_zero);
    }
    // K1c: DC prediction over the plane, restarting at every restart interval; then merged into slot 0 for the comparison
    {
        uint32_t pred[3] = {0, 0, 0};
        for (uint32_t m = 0; m < nmcu; m++) {
            if (ri && m % ri == 0) pred[0] = pred[1] = pred[2] = 0;
            for (uint32_t c = 0; c < bpm; c++) if (m * bpm + c >= first_zero) dcp[m * bpm + c] = 0;
            dc_predict_mcu(g, &dcp[(size_t)m * bpm], pred);
        }
        for (uint32_t u = 0; u < ndu; u++) coef_zz[(size_t)u * 64] = u < first_zero ? dcp[u] : (int16_t)0;
    }
    // zero tail (what the cleanup kernel does)
    uint32_t odd = 0;
    const uint32_t lim = std::min(first_zero, ndu);
    for (uint32_t u = 0; u < lim; u++) odd += written[u] != 1;
    if (first_zero < ndu) memset(coef_zz + (size_t)first_zero * 64, 0, (size_t)(ndu - first_zero) * 128);
    info[0] = rounds; info[1] = first_zero; info[2] = (uint32_t)ns; info[3] = odd;
    return 0;
}

extern "C" int emu_parse(const uint8_t *file, size_t len, bj_image_desc *d) { return parse_header(file, len, d); }

// Direct access to the table builder: every 16-bit window is decoded bit-serially the way the reference does
// (generate_codes + get_next_symbol) and via the table; the entry's fields must describe that symbol.
extern "C" int emu_lut_check(const uint8_t *offsets, const uint8_t *symbols, int ac) {
    std::vector<uint32_t> lut(lut_cap(ac != 0));
    if (build_lut(offsets, symbols, ac != 0, lut.data()) < 0) return -1;
    uint32_t codes[162];
    uint32_t code = 0;
    for (int l = 0; l < 16; l++) { for (unsigned j = offsets[l]; j < offsets[l + 1]; j++) codes[j] = code++; code <<= 1; }
    int bad = 0;
    for (uint32_t w = 0; w < 65536; w++) {               // every 16-bit window
        uint32_t want = kLutNoCode;
        uint32_t cw = 0;
        bool found = false;
        for (int l = 0; l < 16 && !found; l++) {
            cw = (cw << 1) | ((w >> (15 - l)) & 1);
            for (unsigned j = offsets[l]; j < offsets[l + 1]; j++)
                if (cw == codes[j]) { want = lut_leaf(l + 1, symbols[j], ac != 0); found = true; break; }
        }
        LutMem lm;
        lm.attach(lut.data());
        if (lut_lookup(lm, 0, w << 16) != want) bad++;
    }
    return bad;
}

// Byte classification: the word-at-a-time rules (classify_words) against the per-byte ones on an arbitrary byte
// string (16-byte chunks; bytes before / after the string read as 0).  Returns the number of mismatching bytes.
extern "C" int emu_classify_check(const uint8_t *bytes, size_t n) {
    int bad = 0;
    auto at = [&](long long i) -> uint32_t { return (i >= 0 && (size_t)i < n) ? bytes[i] : 0u; };
    for (size_t i0 = 0; i0 < n; i0 += 16) {
        uint32_t w[6] = {0, 0, 0, 0, 0, 0};
        w[0] = at((long long)i0 - 1) << 24;
        for (int b = 0; b < 16; b++) w[1 + b / 4] |= at((long long)i0 + b) << (8 * (b & 3));
        w[5] = at((long long)i0 + 16);
        uint32_t keep, rst, keep2, rst2, end2;
        classify_words(w, keep, rst);
        classify_words_end(w, keep2, rst2, end2);
        bad += keep != keep2 || rst != rst2;
        for (int b = 0; b < 16; b++) {
            const unsigned prev = at((long long)i0 + b - 1), bb = at((long long)i0 + b), next = at((long long)i0 + b + 1);
            bad += ((keep >> b) & 1u) != (scan_keep(prev, bb, next) ? 1u : 0u);
            bad += ((rst >> b) & 1u) != (scan_is_rst(prev, bb) ? 1u : 0u);
            bad += ((end2 >> b) & 1u) != (scan_is_end(bb, next) ? 1u : 0u);
        }
    }
    return bad;
}

// The magnitude extension from a table entry (extend_entry: wrap-mode shifts fed by plain shifts of the entry, sign handled by
// complementing the window) against the plain form extend_value (src/jpeg_scanner.cpp:480-482 / :513-516), for every code
// length, every size (0 = no magnitude bits) and windows that exercise both signs and the extreme magnitudes.
extern "C" int emu_extend_check(void) {
    uint32_t x = 0x2545F491u;
    for (uint32_t len = 1; len <= 16; len++) {
        for (uint32_t size = 0; size <= 11; size++) {
            const uint32_t e = (1u) | ((len + size) << 8) | (len << 16) | (size << 24) | kLutEob;   // (flag bits above the fields must not disturb the shifts)
            for (int i = 0; i < 400; i++) {
                x ^= x << 13; x ^= x >> 17; x ^= x << 5;
                uint32_t win = x;
                if (i == 0) win = 0u;
                if (i == 1) win = 0xFFFFFFFFu;
                if (i == 2) win = 0x80000000u >> len;                     // first magnitude bit set, the rest clear
                if (i == 3) win = ~(0x80000000u >> len);
                if (extend_entry(win, e) != extend_value(win, len, size)) return (int)(len * 100 + size);
            }
        }
    }
    return 0;
}

// The device form of UnitWalk keeps (completed units << 8 | unit index << 4) in one register and advances it by the
// table's step (unit_walk_step); the BitStream can be moved back inside its span (seek).  Both are device-side
// shortcuts of what the host code above does with separate variables: check the arithmetic here.
extern "C" int emu_unit_walk_check(void) {
    for (uint32_t bpm = 1; bpm <= 10; bpm++) {
        for (uint32_t c0 = 0; c0 < bpm; c0++) {
            uint32_t cu = c0 << 4, c = c0, n = 0;
            for (int i = 0; i < 5000; i++) {
                const uint32_t c1 = c + 1 == bpm ? 0 : c + 1;
                if (((cu & 0xF0u) >> 4) != c || (cu >> 4 & 15u) != c || (cu >> 8) != n) return 1;
                cu += unit_walk_step(c, c1);
                c = c1; n++;
            }
        }
    }
    // seek: read windows going forward, jump back, the windows repeat
    uint32_t words[64];
    for (int i = 0; i < 64; i++) words[i] = 0x9E3779B9u * (uint32_t)(i + 1);
    for (uint32_t p0 = 0; p0 < 96; p0 += 7) {
        BitStream bs;
        bs.open(words, p0);
        const uint32_t origin = p0 & ~31u;
        uint32_t S = (p0 - origin) << 8, seen[40], at[40];
        for (int i = 0; i < 40; i++) { at[i] = S; seen[i] = bs.window(S); S += (uint32_t)(1 + (i * 5) % 27) << 8; }
        for (int back = 39; back >= 0; back -= 3) {
            bs.seek(at[back]);
            uint32_t T = at[back];
            for (int i = back; i < 40; i++) { if (bs.window(T) != seen[i]) return 2; T += (uint32_t)(1 + (i * 5) % 27) << 8; }
        }
    }
    return 0;
}
