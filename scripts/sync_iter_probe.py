#!/usr/bin/env python
"""How the synchronisation pass' time splits over its in-CTA iterations (config 2): option "debug_sync_iters" = k stops
the fix-up after k iterations (the result is then not the fixed point - timing only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import pim_jpeg_decoder_b200 as bj

specs, _ = bench.workload_specs("config2", 4096, 1024, 0)
blobs = bench.generate(specs, 16)
dec = bj.Decoder(0)
for iters in (1, 2, 3, 4, 0):
    dec.set_option("debug_sync_iters", iters)
    b = bj.Batch(dec, blobs, bj.BJ_OUT_BMP)
    b.upload()
    t = []
    for _ in range(6):
        b.decode(); b.sync()
        t.append(b.info().ms_sync)
    print(f"iterations {iters or 'all'}: sync {sorted(t)[len(t)//2]:.3f} ms  (write {b.info().ms_write:.3f})")
    b.destroy()
dec.close()
