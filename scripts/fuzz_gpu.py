#!/usr/bin/env python
"""One-off parity fuzz on the GPU (not part of the test suite): random geometry / sampling / quality / Huffman tables
(Annex K or optimised) / restart interval, clean and damaged, decoded in mixed batches through the one-call path and
compared bit for bit with the oracle.  Usage: python scripts/fuzz_gpu.py [n_images] [seed]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import jpeg_synth as js  # noqa: E402
import oracle_lib as ol  # noqa: E402
import pim_jpeg_decoder_b200 as bj  # noqa: E402
from test_huff_emu import _corrupt_scan  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1200
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2026)
    dec = bj.Decoder(0)
    t0 = time.time()
    done = bad = damaged = invalid = 0
    while done < n:
        files = []
        for _ in range(64):
            w, h = int(rng.integers(1, 700)), int(rng.integers(1, 500))
            if rng.integers(0, 12) == 0:
                w, h = int(rng.integers(1, 3000)), int(rng.integers(1, 40))
            sub, gray = int(rng.integers(0, 3)), bool(rng.integers(0, 6) == 0)
            ri = int(rng.choice([0, 0, 0, 1, 2, 5, 8, 33]))
            q, opt = int(rng.choice([35, 60, 75, 90, 95, 100])), bool(rng.integers(0, 2))
            if q == 100 and opt:
                opt = False                 # (Pillow's in-memory encoder gives up on quality 100 + optimised tables: "Suspension not allowed here")
            data = js.synth_jpeg(w, h, seed=int(rng.integers(0, 1 << 30)), subsampling=sub, gray=gray, restart_blocks=ri, quality=q, optimize=opt)
            if ri == 0 and rng.integers(0, 4) == 0:
                try:
                    data, _ = _corrupt_scan(data, rng, int(rng.integers(1, 6)))
                    damaged += 1
                except ValueError:          # (a scan of a few bytes: nothing to flip)
                    pass
            elif rng.integers(0, 25) == 0:
                data = data[: int(rng.integers(len(data) // 2, len(data)))]        # truncated: no EOI
            files.append(data)
        fmt = bj.BJ_OUT_BMP if rng.integers(0, 2) else bj.BJ_OUT_RGB8
        dec.set_option("debug_poison", int(rng.integers(0, 2)))
        outs = []
        for f in files:
            st, d = bj.parse_header(f)
            outs.append(np.full(bj.lib().bj_output_size(d, fmt), 0x77, dtype=np.uint8) if st in (0,) or bj.lib().bj_peek_header(np.frombuffer(f, np.uint8).ctypes.data, len(f), d) == 0 else None)
        res, status = dec.decode(files, fmt, outs=outs)
        for f, o, st in zip(files, outs, status):
            r = ol.Restated(f, 0)
            if not r.valid:
                invalid += 1
                if st not in (bj.BJ_ERR_INVALID_JPEG, bj.BJ_ERR_UNSUPPORTED):
                    bad += 1
                    print("status mismatch on a file the reference rejects:", st, len(f))
                continue
            want = r.bmp if fmt == bj.BJ_OUT_BMP else r.rgb.reshape(-1)
            ok = st == (0 if r.huff_rc == 0 else bj.BJ_ERR_CORRUPT_SCAN) and o is not None and np.array_equal(o, want)
            if not ok:
                bad += 1
                print("MISMATCH", st, r.huff_rc, r.h.width, r.h.height, r.h.hs, r.h.vs, r.h.ncomp, r.h.restart_interval, len(f))
        done += len(files)
    print(f"{done} images ({damaged} damaged, {invalid} rejected) in {time.time() - t0:.1f} s: {bad} mismatches")
    dec.close()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
