"""The <dpu> facade (pim_jpeg_decoder_b200/host/compat/dpu) under the reference's UNMODIFIED host.

CPU test: the reference's decoder_host.cpp / jpeg_scanner.cpp / bmp_writer.cpp are compiled from /root/reference
against the facade and a TEST STUB of the C ABI whose bj_exec_mcus is the oracle (tests/emu/stub_b200jpeg.c), so
what is checked here is the facade's gather/scatter and the host's batching through it - not the kernels.
GPU test (test_gpu_host.py) runs the same host against the real libb200jpeg.so.
"""
import hashlib
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
STUB = os.path.join(HERE, "emu", "_stub")

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "src", "decoder_host.cpp")),
                                reason="reference sources not present (GPU box)")


@pytest.fixture(scope="module")
def compat_emu():
    os.makedirs(STUB, exist_ok=True)
    inc = os.path.join(ROOT, "include")
    lib = os.path.join(STUB, "libb200jpeg.so")
    subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-I", inc, "-I", os.path.join(ROOT, "oracle"), "-o", lib,
                    os.path.join(HERE, "emu", "stub_b200jpeg.c"), os.path.join(ROOT, "oracle", "restate.c")], check=True)
    exe = os.path.join(STUB, "decoder_compat_emu")
    subprocess.run(["g++", "--std=c++11", "-O2", "-DMAX_MCU_PER_DPU=100", "-include", "cstdint", "-include", "cstdlib", "-w",
                    "-I", os.path.join(ROOT, "pim_jpeg_decoder_b200", "host", "compat"), "-I", inc, "-I", os.path.join(REF, "src"),
                    "-o", exe] + [os.path.join(REF, "src", f) for f in ("decoder_host.cpp", "jpeg_scanner.cpp", "bmp_writer.cpp")] +
                   ["-L", STUB, "-lb200jpeg", "-Wl,-rpath," + STUB, "-lpthread"], check=True)
    return exe


@pytest.mark.parametrize("nr_dpus", [64, 7])
def test_reference_host_over_facade(compat_emu, nr_dpus, golden, golden_dir, tmp_path):
    """Several images per exec() batch, and (7 DPUs) a batch flush in the middle of the list."""
    names = ["p420_50x37", "p444_33x17", "gray_40x24", "p422_70x33", "p420_48x40", "enc_444_100x60_ri13"]
    paths = []
    for n in names:
        p = str(tmp_path / golden[n]["file"])
        shutil.copy(os.path.join(golden_dir, golden[n]["file"]), p)
        paths.append(p)
    env = dict(os.environ, B200JPEG_NR_DPUS=str(nr_dpus))
    out = subprocess.run([compat_emu] + paths, env=env, capture_output=True, text=True, check=True).stdout
    assert f"{nr_dpus} dpus are allocated" in out
    for n, p in zip(names, paths):
        got = hashlib.sha256(open(p[:-4] + ".bmp", "rb").read()).hexdigest()
        assert got == golden[n]["bmp_sha256"], n          # the hash the reference itself produced
