"""Synthetic baseline-JPEG generators for the tests and the benchmark (no network, no datasets).

* ``synth_rgb`` / ``pil_jpeg``: the generator SURVEY.md section 8d prescribes (smooth field + noise, Pillow /
  libjpeg-turbo encoder, quality 90, Annex-K tables unless ``optimize``).
* ``encode_from_coefs``: a tiny coefficient-level baseline encoder, so tests can place exact quantised
  coefficients (zig-zag 48 / 52 aliasing, ZRL landings, extreme values that exercise the 16-bit wraps),
  choose any sampling the reference accepts (incl. h1v2, which Pillow cannot write) and any restart interval.
"""
import io

import numpy as np


_FIELDS = {}


def _field(w, h):
    """The seed-independent smooth part of synth_rgb (cached per size: it dominates generation time)."""
    f = _FIELDS.get((w, h))
    if f is None:
        y, x = np.mgrid[0:h, 0:w].astype(np.float64)
        r = 128 + 100 * np.sin(x / 37.0) * np.cos(y / 23.0)
        g = 128 + 90 * np.sin((x + y) / 51.0)
        b = 128 + 80 * np.cos(x / 17.0 - y / 29.0)
        f = np.stack([r, g, b], axis=-1)
        if len(_FIELDS) > 16:
            _FIELDS.clear()
        _FIELDS[(w, h)] = f
    return f


def synth_rgb(w, h, seed):
    rng = np.random.default_rng(seed)
    img = _field(w, h) + rng.normal(0.0, 12.0, size=(h, w, 3))
    return np.clip(img, 0, 255).astype(np.uint8)


def pil_jpeg(arr, quality=90, subsampling=2, optimize=False, restart_blocks=0, gray=False):
    from PIL import Image

    im = Image.fromarray(arr)
    if gray:
        im = im.convert("L")
    buf = io.BytesIO()
    kw = dict(quality=quality, optimize=optimize, progressive=False)
    if not gray:
        kw["subsampling"] = subsampling
    if restart_blocks:
        kw["restart_marker_blocks"] = restart_blocks
    im.save(buf, "JPEG", **kw)
    return buf.getvalue()


def synth_jpeg(w, h, seed, subsampling=2, quality=90, restart_blocks=0, gray=False, optimize=False):
    return pil_jpeg(synth_rgb(w, h, seed), quality, subsampling, optimize, restart_blocks, gray)


# --------------------------------------------------------------------------------------------------
# coefficient-level encoder

def _std_tables():
    """Annex-K Huffman tables + quality-90 quantisation tables, lifted from a Pillow file (optimize=False)."""
    data = pil_jpeg(synth_rgb(16, 16, 0), 90, 2)
    i, dht, dqt = 2, {}, {}
    while i < len(data):
        assert data[i] == 0xFF
        m = data[i + 1]
        if m == 0xDA:
            break
        ln = (data[i + 2] << 8) | data[i + 3]
        seg = data[i + 4:i + 2 + ln]
        if m == 0xC4:
            j = 0
            while j < len(seg):
                info = seg[j]
                counts = list(seg[j + 1:j + 17])
                n = sum(counts)
                dht[(info >> 4, info & 15)] = (counts, list(seg[j + 17:j + 17 + n]))
                j += 17 + n
        elif m == 0xDB:
            j = 0
            while j < len(seg):
                info = seg[j]
                assert info >> 4 == 0
                dqt[info & 15] = list(seg[j + 1:j + 65])
                j += 65
        i += 2 + ln
    return dht, dqt


_STD = None


def std_tables():
    global _STD
    if _STD is None:
        _STD = _std_tables()
    return _STD


def _code_map(counts, symbols):
    code, k, out = 0, 0, {}
    for ln in range(1, 17):
        for _ in range(counts[ln - 1]):
            out[symbols[k]] = (code, ln)
            code += 1
            k += 1
        code <<= 1
    return out


class _Bits:
    def __init__(self):
        self.out = bytearray()
        self.acc = 0
        self.n = 0

    def put(self, v, ln):
        if ln == 0:
            return
        self.acc = (self.acc << ln) | (v & ((1 << ln) - 1))
        self.n += ln
        while self.n >= 8:
            b = (self.acc >> (self.n - 8)) & 0xFF
            self.out.append(b)
            if b == 0xFF:
                self.out.append(0)
            self.n -= 8
        self.acc &= (1 << self.n) - 1

    def flush(self):
        if self.n:
            self.put((1 << (8 - self.n)) - 1, 8 - self.n)


def _mag(v):
    a = abs(int(v))
    s = a.bit_length()
    return s, (int(v) if v >= 0 else int(v) + (1 << s) - 1)


def encode_from_coefs(width, height, comps, coefs, qts=None, restart_interval=0, tables=None,
                      qt16=False, trailing=b""):
    """comps: [(h, v, qt_id, dc_id, ac_id)] (first = luma); coefs: per component an int array
    [rows_of_units][cols_of_units][64] of QUANTISED coefficients in ZIG-ZAG order (DC absolute), covering the
    MCU-padded grid.  Returns JPEG bytes (baseline, interleaved single scan)."""
    dht, dqt = tables if tables is not None else std_tables()
    qts = qts if qts is not None else dqt
    hs, vs = comps[0][0], comps[0][1]
    out = bytearray(b"\xFF\xD8")
    for tid in sorted(set(c[2] for c in comps)):
        if qt16:
            out += b"\xFF\xDB" + (2 + 129).to_bytes(2, "big") + bytes([0x10 | tid])
            for q in qts[tid]:
                out += int(q).to_bytes(2, "big")
        else:
            out += b"\xFF\xDB" + (2 + 65).to_bytes(2, "big") + bytes([tid]) + bytes(int(q) for q in qts[tid])
    out += b"\xFF\xC0" + (8 + 3 * len(comps)).to_bytes(2, "big") + b"\x08" + height.to_bytes(2, "big") + width.to_bytes(2, "big")
    out += bytes([len(comps)])
    for i, c in enumerate(comps):
        out += bytes([i + 1, (c[0] << 4) | c[1], c[2]])
    used = sorted(set((0, c[3]) for c in comps) | set((1, c[4]) for c in comps))
    for cls, tid in used:
        counts, syms = dht[(cls, tid)]
        out += b"\xFF\xC4" + (2 + 17 + len(syms)).to_bytes(2, "big") + bytes([(cls << 4) | tid]) + bytes(counts) + bytes(syms)
    if restart_interval:
        out += b"\xFF\xDD\x00\x04" + restart_interval.to_bytes(2, "big")
    out += b"\xFF\xDA" + (6 + 2 * len(comps)).to_bytes(2, "big") + bytes([len(comps)])
    for i, c in enumerate(comps):
        out += bytes([i + 1, (c[3] << 4) | c[4]])
    out += b"\x00\x3F\x00"

    maps = {k: _code_map(*v) for k, v in dht.items()}
    mw, mh = (width + 7) // 8, (height + 7) // 8
    bits = _Bits()
    pred = [0] * len(comps)
    mcu, rst = 0, 0
    for y in range(0, mh, vs):
        for x in range(0, mw, hs):
            if restart_interval and mcu and mcu % restart_interval == 0:
                bits.flush()
                bits.out += bytes([0xFF, 0xD0 + (rst & 7)])
                rst += 1
                pred = [0] * len(comps)
            for j, c in enumerate(comps):
                dcm, acm = maps[(0, c[3])], maps[(1, c[4])]
                for v in range(c[1]):
                    for h in range(c[0]):
                        by = (y // vs) * c[1] + v
                        bx = (x // hs) * c[0] + h
                        zz = coefs[j][by][bx]
                        d = int(zz[0]) - pred[j]
                        pred[j] = int(zz[0])
                        s, bitsv = _mag(d)
                        bits.put(*dcm[s])
                        bits.put(bitsv, s)
                        run = 0
                        last = max([k for k in range(1, 64) if zz[k] != 0], default=0)
                        for k in range(1, last + 1):
                            if zz[k] == 0:
                                run += 1
                                continue
                            while run > 15:
                                bits.put(*acm[0xF0])
                                run -= 16
                            s, bitsv = _mag(zz[k])
                            bits.put(*acm[(run << 4) | s])
                            bits.put(bitsv, s)
                            run = 0
                        if last < 63:
                            bits.put(*acm[0x00])
            mcu += 1
    bits.flush()
    out += bits.out + b"\xFF\xD9" + trailing
    return bytes(out)


def random_coefs(comps, width, height, seed, density=0.15, amp=40, dc_amp=200):
    """Random sparse quantised coefficients on the MCU-padded grid of every component."""
    rng = np.random.default_rng(seed)
    hs, vs = comps[0][0], comps[0][1]
    mw, mh = (width + 7) // 8, (height + 7) // 8
    nmx, nmy = (mw + hs - 1) // hs, (mh + vs - 1) // vs
    out = []
    for c in comps:
        rows, cols = nmy * c[1], nmx * c[0]
        a = rng.integers(-amp, amp + 1, size=(rows, cols, 64))
        a[rng.random(size=a.shape) > density] = 0
        a[..., 0] = rng.integers(-dc_amp, dc_amp + 1, size=(rows, cols))
        out.append(a)
    return out
