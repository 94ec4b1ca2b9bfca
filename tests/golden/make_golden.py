"""Generates tests/golden/: small fixture JPEGs + golden.json (hashes produced by the REAL reference).

Run in the dev container (needs /root/reference and `make -C oracle ref`):  python tests/golden/make_golden.py

For every fixture the reference's own CLI (oracle/_ref/decoder = /root/reference/src compiled verbatim against
the fake UPMEM runtime) writes the BMP, and oracle/_ref/libref.so dumps the post-Huffman and post-exec buffers.
golden.json pins SHA-256 of all three.  `expect` names the fixture whose BMP the GPU path must reproduce:
itself, except for subsampled files WITH restart markers, where the reference mis-decodes (SURVEY.md section 0,
fact 7) and the restart-parity rule (section 8c) points at the restart-free twin made from the same coefficients.
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import jpeg_synth as js  # noqa: E402
import oracle_lib as ol  # noqa: E402

REF_IMAGE = "/root/reference/ILSVRC2012_val_00000001.JPEG"
Y420 = [(2, 2, 0, 0, 0), (1, 1, 1, 1, 1), (1, 1, 1, 1, 1)]
Y422 = [(2, 1, 0, 0, 0), (1, 1, 1, 1, 1), (1, 1, 1, 1, 1)]
Y440 = [(1, 2, 0, 0, 0), (1, 1, 1, 1, 1), (1, 1, 1, 1, 1)]
Y444 = [(1, 1, 0, 0, 0), (1, 1, 1, 1, 1), (1, 1, 1, 1, 1)]
GRAY = [(1, 1, 0, 0, 0)]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def quirk_coefs(comps, w, h):
    """Blocks that exercise the zig-zag 48/52 aliasing and ZRL landings."""
    co = js.random_coefs(comps, w, h, 11, density=0.0)
    pats = [{48: 5}, {52: -7}, {48: 5, 52: -7}, {36: 3, 52: 9}, {20: 2, 48: -4, 63: 1}, {32: 1, 48: 6}, {1: 300, 2: -300, 52: 1023, 48: -1023}]
    k = 0
    for a in co:
        for by in range(a.shape[0]):
            for bx in range(a.shape[1]):
                for z, v in pats[k % len(pats)].items():
                    a[by, bx, z] = v
                k += 1
    return co


def extreme_coefs(comps, w, h):
    """Large coefficients: drives the (short) wraps after dequantisation and between the IDCT passes."""
    co = js.random_coefs(comps, w, h, 13, density=0.5, amp=1023, dc_amp=1000)
    return co


def fixtures():
    f = {}
    f["ilsvrc_444"] = dict(data=open(REF_IMAGE, "rb").read(), note="the reference's bundled image (config 1)")
    f["p420_48x40"] = dict(data=js.synth_jpeg(48, 40, 1, 2))
    f["p420_50x37"] = dict(data=js.synth_jpeg(50, 37, 2, 2), note="odd position counts -> padded MCU grid")
    f["p422_70x33"] = dict(data=js.synth_jpeg(70, 33, 3, 1))
    f["p444_33x17"] = dict(data=js.synth_jpeg(33, 17, 4, 0))
    f["gray_40x24"] = dict(data=js.synth_jpeg(40, 24, 5, gray=True))
    f["p420_1x1"] = dict(data=js.synth_jpeg(1, 1, 6, 2))
    f["p420_17x17"] = dict(data=js.synth_jpeg(17, 17, 7, 2))
    f["p420_opt_96x64"] = dict(data=js.synth_jpeg(96, 64, 8, 2, optimize=True), note="optimised Huffman tables")
    f["p420_q30_200x120"] = dict(data=js.synth_jpeg(200, 120, 9, 2, quality=30), note="long zero runs / short blocks")
    f["p420_q100_64x64"] = dict(data=js.synth_jpeg(64, 64, 10, 2, quality=100), note="dense blocks, long codes")
    f["p444_ri2_64x48"] = dict(data=js.synth_jpeg(64, 48, 11, 0, restart_blocks=2))
    f["gray_ri3_64x48"] = dict(data=js.synth_jpeg(64, 48, 12, gray=True, restart_blocks=3))
    f["p420_320x240"] = dict(data=js.synth_jpeg(320, 240, 13, 2))
    # coefficient-level files
    for name, comps in (("420", Y420), ("422", Y422), ("440", Y440), ("444", Y444), ("gray", GRAY)):
        co = js.random_coefs(comps, 45, 30, 21)
        f["enc_%s_45x30" % name] = dict(data=js.encode_from_coefs(45, 30, comps, co))
        co = quirk_coefs(comps, 40, 24)
        f["enc_%s_zz48_52" % name] = dict(data=js.encode_from_coefs(40, 24, comps, co), note="zig-zag 48/52 aliasing, ZRL landings")
    co = extreme_coefs(Y420, 32, 32)
    f["enc_420_extreme"] = dict(data=js.encode_from_coefs(32, 32, Y420, co), note="16-bit wraps")
    f["enc_444_qt16"] = dict(data=js.encode_from_coefs(24, 16, Y444, js.random_coefs(Y444, 24, 16, 5), qt16=True,
                                                       qts={0: [257 + 3 * i for i in range(64)], 1: [300 + i for i in range(64)]}),
                             note="16-bit quantisation tables")
    # restart files: subsampled ones point at their restart-free twin (restart-parity rule)
    for name, comps in (("420", Y420), ("422", Y422), ("440", Y440)):
        co = js.random_coefs(comps, 100, 60, 31)
        f["enc_%s_100x60" % name] = dict(data=js.encode_from_coefs(100, 60, comps, co))
        f["enc_%s_100x60_ri4" % name] = dict(data=js.encode_from_coefs(100, 60, comps, co, restart_interval=4), expect="enc_%s_100x60" % name,
                                              note="reference mis-decodes subsampled+DRI; twin rule")
    co = js.random_coefs(Y444, 100, 60, 32)
    f["enc_444_100x60_ri1"] = dict(data=js.encode_from_coefs(100, 60, Y444, co, restart_interval=1))
    f["enc_444_100x60_ri13"] = dict(data=js.encode_from_coefs(100, 60, Y444, co, restart_interval=13))
    f["enc_gray_trailing"] = dict(data=js.encode_from_coefs(30, 30, GRAY, js.random_coefs(GRAY, 30, 30, 33), trailing=b"\x00garbage after EOI\xff\xd9"))
    # invalid files the reference rejects
    good = js.synth_jpeg(32, 32, 14, 2)
    f["bad_truncated"] = dict(data=good[: len(good) - 40], invalid=True)
    f["bad_not_jpeg"] = dict(data=b"\x89PNG\r\n\x1a\n" + bytes(64), invalid=True)
    return f


def main():
    ol.build_oracle()
    tmp = tempfile.mkdtemp()
    out = {}
    fx = fixtures()
    for name, d in fx.items():
        path = os.path.join(HERE, name + ".jpg")
        with open(path, "wb") as fh:
            fh.write(d["data"])
        e = dict(file=name + ".jpg", bytes=len(d["data"]), note=d.get("note", ""), expect=d.get("expect", name))
        work = os.path.join(tmp, name + ".jpg")
        shutil.copy(path, work)
        subprocess.run([os.path.join(ol.ORACLE, "_ref", "decoder"), work], check=True, stdout=subprocess.DEVNULL,
                       env=dict(os.environ, ORACLE_NR_DPUS="64"))
        bmp = os.path.join(tmp, name + ".bmp")
        if d.get("invalid"):
            assert not os.path.exists(bmp), name
            e["invalid"] = True
        else:
            e["bmp_sha256"] = sha(np.fromfile(bmp, dtype=np.uint8))
            r = ol.RefDecoded(work)
            e.update(width=r.info.width, height=r.info.height, ncomp=r.info.ncomp, hs=r.info.h_samp, vs=r.info.v_samp,
                     restart_interval=r.info.restart_interval, nchunks=r.nchunk, huffman_ok=int(r.info.huffman_ok),
                     mcus_pre_sha256=sha(r.mcus_pre), mcus_post_sha256=sha(r.mcus_post), metadata_sha256=sha(r.metadata))
        out[name] = e
    with open(os.path.join(HERE, "golden.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print("wrote", len(out), "fixtures,", sum(e["bytes"] for e in out.values()), "bytes")


if __name__ == "__main__":
    main()
