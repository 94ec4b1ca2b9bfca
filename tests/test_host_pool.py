"""CPU test of the library's host worker pool (bj::HostPool, csrc/bj_host.h): the per-image host work of a batch
(parse + pack) is split over it, so every index must be visited exactly once, for any thread count / chunking."""
import ctypes as C
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(HERE, "emu", "libhostpoolemu.so")
    subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-I", "/usr/local/cuda/include", "-o", so,
                    os.path.join(HERE, "emu", "hostpool_emu.cpp"), "-lpthread"], check=True)
    return C.CDLL(so)


@pytest.mark.parametrize("threads,n,chunk", [(1, 100, 7), (2, 1, 1), (4, 4096, 32), (8, 585, 16), (8, 0, 4), (16, 31, 64), (3, 1000, 1)])
def test_parallel_for_visits_each_index_once(emu, threads, n, chunk):
    assert emu.emu_hostpool(threads, n, chunk, 50) == 0
