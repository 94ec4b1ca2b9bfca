"""CPU tests of the entropy stage's ALGORITHM: the product's shared device/host core (huff_core.h, parse.h),
driven by a sequential host emulation of the kernels' data flow (tests/emu/huff_emu.cpp), against the oracle.

The GPU kernels themselves are covered by the ``-m gpu`` tests; this file checks what can be checked without one:
the lookup-table builder, the byte classification, the speculative decode + fix-up fixed point, the prefix sums
and the owner-writes-whole-unit rule, on every golden fixture and on synthetic images, at several sub-sequence
sizes (small sizes force units and even single symbols to straddle sub-sequences).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import jpeg_synth as js
import oracle_lib as ol

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu", "libhuffemu.so")


def emu():
    src = os.path.join(HERE, "emu", "huff_emu.cpp")
    core = os.path.join(ol.ROOT, "pim_jpeg_decoder_b200", "csrc", "huff_core.h")
    if not os.path.exists(EMU) or os.path.getmtime(EMU) < max(os.path.getmtime(src), os.path.getmtime(core)):
        subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-o", EMU, src], check=True)
    lib = C.CDLL(EMU)
    lib.emu_entropy.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.emu_lut_check.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.emu_set_write_rounds.argtypes = [C.c_int]
    return lib


def run_emu(data, sub_bytes, slices=4):
    r = ol.Restated(data, restart_mode=0)
    assert r.valid
    buf = np.frombuffer(data, dtype=np.uint8)
    got = np.full_like(r.coef_zz, 0x5A5A)          # poison: every unit must be written (or zero-filled)
    info = np.zeros(4, dtype=np.uint32)
    rc = emu().emu_entropy(ol._ptr(buf), len(data), sub_bytes, slices, ol._ptr(got), ol._ptr(info))
    assert rc == 0
    return r, got, info


def _golden_names():
    import json
    with open(os.path.join(HERE, "golden", "golden.json")) as f:
        g = json.load(f)
    return sorted(k for k, v in g.items() if not v.get("invalid"))


@pytest.mark.parametrize("sub_bytes,slices", [(4, 1), (4, 4), (32, 2), (128, 4), (1024, 1), (16, 8)])
@pytest.mark.parametrize("name", _golden_names())
def test_emulated_entropy_stage_matches_oracle(name, sub_bytes, slices, golden, golden_dir):
    """sub_bytes = slice (write-pass) granularity; the synchronisation pass works on `slices` of them at a time."""
    data = open(os.path.join(golden_dir, golden[name]["file"]), "rb").read()
    r, got, info = run_emu(data, sub_bytes, slices)
    assert r.huff_rc == 0
    assert info[3] == 0, "a data unit was written twice or never"
    assert info[1] == 0xFFFFFFFF
    assert np.array_equal(got, r.coef_zz)


@pytest.mark.parametrize("w,h,sub,gray,ri", [(500, 375, 2, False, 0), (500, 375, 0, False, 0), (640, 480, 1, False, 0),
                                             (320, 200, 2, True, 0), (512, 512, 2, False, 8), (333, 111, 0, False, 5)])
def test_emulated_entropy_stage_synthetic(w, h, sub, gray, ri):
    data = js.synth_jpeg(w, h, seed=w + h, subsampling=sub, gray=gray, restart_blocks=ri)
    r, got, info = run_emu(data, 128)
    assert r.huff_rc == 0 and info[3] == 0
    assert np.array_equal(got, r.coef_zz)
    assert info[0] < 40            # fix-up rounds stay small: the stream self-synchronises


def test_emulated_entropy_stage_long_restart_free_stream():
    """A restart-free scan longer than 2^24 bits: positions inside a span are kept relative to the span's origin
    in 24 bits, and the end of the DATA (the whole scan here) must saturate instead of wrapping."""
    data = js.synth_jpeg(2048, 1600, seed=77, subsampling=0)
    r = ol.Restated(data, 0)
    assert (r.h.scan_len * 8) > (1 << 24)
    _, got, info = run_emu(data, 512)
    assert info[1] == 0xFFFFFFFF and info[3] == 0
    assert np.array_equal(got, r.coef_zz)


def test_truncated_scan_zero_fills_like_the_reference():
    """Cut the entropy-coded data short: the reference stops at the failing unit and leaves the rest zero."""
    data = bytearray(js.synth_jpeg(160, 120, seed=3, subsampling=2))
    cut = len(data) - 2 - 900
    bad = bytes(data[:cut]) + b"\xFF\xD9"
    r, got, info = run_emu(bad, 128)
    assert r.huff_rc != 0
    assert info[1] != 0xFFFFFFFF
    assert np.array_equal(got, r.coef_zz)


@pytest.fixture
def write_rounds():
    """Sets how the emulator runs the write pass (0: one symbol per step, slice after slice; n: warps of 32 slices in
    rounds of n symbols, look-back + hand-over once per round, units stored in slot order - the control flow of
    k_huff_write) and puts the default back afterwards."""
    lib = emu()
    yield lib.emu_set_write_rounds
    lib.emu_set_write_rounds(0)


@pytest.mark.parametrize("rounds", [1, 3, 8])
@pytest.mark.parametrize("sub_bytes,slices", [(4, 1), (16, 8), (128, 4)])
def test_write_pass_in_rounds_matches_oracle(rounds, sub_bytes, slices, write_rounds, golden, golden_dir):
    """The write pass as the kernel runs it: several symbols per lane and round with no check in between, the look-back
    over a completed unit, the hand-over to the next unit and the stores once per round."""
    write_rounds(rounds)
    for name in _golden_names():
        data = open(os.path.join(golden_dir, golden[name]["file"]), "rb").read()
        r, got, info = run_emu(data, sub_bytes, slices)
        assert r.huff_rc == 0, name
        assert info[3] == 0, "a data unit was written twice or never"
        assert np.array_equal(got, r.coef_zz), name


@pytest.mark.parametrize("rounds", [1, 4, 8])
@pytest.mark.parametrize("sub,sub_bytes,slices", [(2, 128, 4), (0, 16, 8), (1, 1024, 1)])
def test_write_pass_in_rounds_stops_where_the_reference_stops(rounds, sub, sub_bytes, slices, write_rounds):
    """Damaged Huffman data through the rounds: a unit that did not end well is found by the look-back at the end of the
    round (no symbol is checked on the way) and decoded again, exactly; truncated data likewise."""
    write_rounds(rounds)
    rng = np.random.default_rng(77 * rounds + 1000 * sub + slices)
    base = js.synth_jpeg(200, 152, seed=40 + sub, subsampling=sub)
    failed = 0
    for trial in range(16):
        bad, _ = _corrupt_scan(base, rng, 1 + trial % 8)
        r, got, info = run_emu(bad, sub_bytes, slices)
        failed += r.huff_rc != 0
        assert (info[1] != 0xFFFFFFFF) == (r.huff_rc != 0), trial
        assert np.array_equal(got, r.coef_zz), trial
    assert failed >= 2
    cut = bytes(base[:len(base) - 2 - 700]) + b"\xFF\xD9"
    r, got, info = run_emu(cut, sub_bytes, slices)
    assert r.huff_rc != 0 and info[1] != 0xFFFFFFFF
    assert np.array_equal(got, r.coef_zz)


def _corrupt_scan(data, rng, flips):
    """Flip `flips` random bits inside the entropy-coded data; bytes that become FF are set to FE so that no marker
    or stuffing pair appears (the damage stays inside the Huffman data: refused codes, over-long runs, bits running
    out, wrong unit counts).  Returns the damaged file and the restart segment of the first damaged byte."""
    h = ol.Restated(data, 0).h
    a = bytearray(data)
    lo, hi = h.scan_off + 4, h.scan_off + h.scan_len - 4
    first = None
    for pos in rng.integers(lo, hi, flips):
        if a[pos] == 0xFF or a[pos - 1] == 0xFF:
            continue                                    # leave stuffed pairs and markers alone
        a[pos] ^= 1 << int(rng.integers(0, 8))
        if a[pos] == 0xFF:
            a[pos] = 0xFE
        first = int(pos) if first is None else min(first, int(pos))
    seg = 0
    if first is not None:
        scan = bytes(a[h.scan_off:first])
        seg = sum(1 for i in range(len(scan) - 1) if scan[i] == 0xFF and 0xD0 <= scan[i + 1] <= 0xD7)
    return bytes(a), seg


@pytest.mark.parametrize("sub,sub_bytes,slices", [(2, 128, 4), (0, 32, 2), (2, 64, 1), (0, 16, 8), (1, 1024, 1)])
def test_corrupted_scans_stop_where_the_reference_stops(sub, sub_bytes, slices):
    """Damaged Huffman data (no restart markers): the reference stops at the first refused code / over-long run /
    missing bit and leaves every later coefficient zero (src/jpeg_scanner.cpp:467-520, result ignored at
    src/decoder_host.cpp:181); units before the failure keep their values, the failing unit keeps what was stored
    before the failing symbol.  Bit-exact, including which unit is the first zero one."""
    rng = np.random.default_rng(1000 * sub + slices)
    base = js.synth_jpeg(200, 152, seed=40 + sub, subsampling=sub)
    failed = 0
    for trial in range(24):
        bad, _ = _corrupt_scan(base, rng, 1 + trial % 8)
        r, got, info = run_emu(bad, sub_bytes, slices)
        failed += r.huff_rc != 0
        assert (info[1] != 0xFFFFFFFF) == (r.huff_rc != 0), trial
        assert np.array_equal(got, r.coef_zz), trial
    assert failed >= 2                                   # the damage does make the reference fail


@pytest.mark.parametrize("sub,ri,sub_bytes,slices", [(2, 4, 64, 1), (0, 3, 16, 8), (0, 7, 128, 2)])
def test_corrupted_scans_with_restart_markers_agree_up_to_the_damaged_segment(sub, ri, sub_bytes, slices):
    """With restart markers a damaged file is where this back end and the reference part ways ON PURPOSE: the reference's
    scanner has thrown the marker positions away, so after damage inside a segment its bit
    position drifts and everything that follows is garbage or zero; here every segment starts at its marker.  What
    must hold: every unit of the segments before the damaged one is the reference's, bit for bit."""
    rng = np.random.default_rng(1000 * sub + 10 * ri + slices)
    base = js.synth_jpeg(200, 152, seed=40 + sub + ri, subsampling=sub, restart_blocks=ri)
    good = ol.Restated(base, 0)
    bpm = {0: 3, 1: 4, 2: 6}[sub]                        # data units per MCU
    for trial in range(12):
        bad, seg = _corrupt_scan(base, rng, 1 + trial % 4)
        r, got, info = run_emu(bad, sub_bytes, slices)
        n = seg * ri * bpm
        assert np.array_equal(got[:n], r.coef_zz[:n]), trial
        assert np.array_equal(got[:n], good.coef_zz[:n]), trial


def test_word_at_a_time_byte_classification():
    """FF00 un-stuffing / RSTn / fill-byte rules four bytes per word (bit tricks) == the per-byte rules, on byte
    strings dense in the special values and on every pair of adjacent byte values."""
    lib = emu()
    lib.emu_classify_check.argtypes = [C.c_void_p, C.c_size_t]
    rng = np.random.default_rng(7)
    special = np.array([0xFF, 0x00, 0xD0, 0xD3, 0xD7, 0xD8, 0xCF, 0x7F, 0x80, 0xFE, 0x01], dtype=np.uint8)
    for dens in (0.1, 0.5, 0.9):
        a = rng.integers(0, 256, 1 << 16, dtype=np.uint8)
        m = rng.random(a.size) < dens
        a[m] = special[rng.integers(0, special.size, int(m.sum()))]
        assert lib.emu_classify_check(ol._ptr(a), a.size) == 0
    pairs = np.stack(np.meshgrid(np.arange(256), np.arange(256), indexing="ij"), -1).reshape(-1).astype(np.uint8)
    for shift in range(4):                      # every pair at every alignment inside a word
        b = np.concatenate([np.full(shift, 0x55, np.uint8), pairs])
        assert lib.emu_classify_check(ol._ptr(b), b.size) == 0


def _scan_soup(rng, nbytes):
    """Scan bytes dense in everything the scan-byte rules care about (never an end-of-scan marker)."""
    parts, n = [], 0
    while n < nbytes:
        k = int(rng.integers(0, 6))
        if k == 0:
            piece = bytes([0xFF, 0x00])
        elif k == 1:
            piece = bytes([0xFF, 0xD0 + int(rng.integers(0, 8))])
        elif k == 2:
            piece = bytes([0xFF] * int(rng.integers(1, 4)) + [0xFF, 0xD0 + int(rng.integers(0, 8))])   # fill bytes before a marker
        else:
            piece = bytes(int(x) for x in rng.integers(0, 255, int(rng.integers(1, 40))))            # no FF
        parts.append(piece)
        n += len(piece)
    return b"".join(parts)


def test_device_side_scan_end_counts_and_compaction():
    """K0 never gets a scan length from the host: k_unstuff finds the FF that ends the scan, counts what
    survives in front of it per 16 KB tile (look-back between the tiles), and compacts.  Their shared code, run tile by tile and chunk by
    chunk like the kernels, against the per-byte rules and the host's scan walk: scans dense in stuffed FFs, restart
    markers and fill bytes, every alignment of the file in the device buffer, arbitrary bytes around the file, ends on
    tile and chunk boundaries, trailing garbage behind EOI, and files the reference rejects (no EOI, another marker
    inside the scan, FF as the last byte)."""
    lib = emu()
    lib.emu_k0_check.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int]
    base = js.synth_jpeg(64, 48, seed=1, subsampling=2)
    h = ol.Restated(base, 0).h
    head = base[:h.scan_off]
    rng = np.random.default_rng(11)

    def check(data, expect_valid, nseg=1 << 20):
        buf = np.frombuffer(data, dtype=np.uint8)
        for pos, fill in ((0, 0x00), (1, 0xFF), (7, 0xD9), (15, 0xFF), (16, 0x00), (16387, 0xFF)):
            assert lib.emu_k0_check(ol._ptr(buf), len(data), pos, fill, nseg) == 0, (len(data), pos, fill)
        assert (ol.Restated(data, 0).rc == 0) == expect_valid

    for trial in range(6):
        soup = _scan_soup(rng, 2 * 16384 + 3000 * trial)
        check(head + soup + b"\xFF\xD9", True)
        check(head + soup + b"\xFF\xD9" + _scan_soup(rng, 700) + b"\xFF\xD9\xFF\xC0junk", True)      # garbage behind EOI
        check(head + soup, False)                                                                    # file ends inside the scan
        check(head + soup + b"\xFF", False)                                                          # ... on an FF
        check(head + soup + b"\xFF\xC4" + soup[:100] + b"\xFF\xD9", False)                           # another marker first
        check(head + soup + b"\xFF\xFF\xFF\xD9", True)                                               # fill bytes before EOI
        check(head + soup + b"\xFF\xD9", True, nseg=3)                                               # more markers than segments expected
    # the end at every position around a tile / chunk boundary
    soup = _scan_soup(rng, 2 * 16384)
    for cut in list(range(16384 - len(head) - 20, 16384 - len(head) + 20)) + [16, 17, 31, 32, 33, 63, 64, 65, 2047 - len(head), 2048 - len(head), 2049 - len(head)]:
        check(head + soup[:cut].rstrip(b"\xFF") + b"\x01\xFF\xD9", True)
    check(head + b"\xFF\xD9", True)                                                                  # empty scan
    check(head, False)


def test_lut_matches_bit_serial_search():
    lib = emu()
    dht, _ = js.std_tables()
    for key, (counts, syms) in dht.items():
        off = np.zeros(17, dtype=np.uint8)
        off[1:] = np.cumsum(counts)
        sy = np.zeros(162, dtype=np.uint8)
        sy[:len(syms)] = syms
        assert lib.emu_lut_check(ol._ptr(off), ol._ptr(sy), key[0]) == 0, key
    # an optimised table set and a deliberately odd one (sparse lengths incl. 16-bit DC codes)
    off = np.zeros(17, dtype=np.uint8)
    counts = [0, 1, 0, 0, 2, 0, 0, 0, 3, 0, 0, 5, 0, 0, 7, 20]
    off[1:] = np.cumsum(counts)
    sy = np.arange(162, dtype=np.uint8)
    assert lib.emu_lut_check(ol._ptr(off), ol._ptr(sy), 0) == 0
    assert lib.emu_lut_check(ol._ptr(off), ol._ptr(sy), 1) == 0


def test_magnitude_extension_from_table_entry():
    """extend_entry (what the write pass computes per symbol: eight instructions on the device) against the plain
    form of the reference's extension, for every code length x size and both signs."""
    assert emu().emu_extend_check() == 0


def test_unit_walk_and_stream_seek_arithmetic():
    """The packed unit counter of the synchronisation pass (completed << 8 | unit << 4, advanced by the table's step)
    and the bit reader's seek (the write pass re-reads a damaged unit) against the plain forms."""
    assert emu().emu_unit_walk_check() == 0
