#!/usr/bin/env python
"""What the host can take back from all its GPUs at once, by kind of host memory.

One process, one stream per GPU, 1 GiB device -> host per GPU, all copies in flight together; wall clock around them.
Kinds of destination memory: cudaHostAlloc default, cudaHostAlloc write-combined, anonymous mmap + transparent huge pages
+ cudaHostRegister, MAP_HUGETLB + cudaHostRegister (if the box has huge pages configured).
Usage: python scripts/d2h_probe_multi.py > profiles/<name>.txt
"""
import ctypes as C
import mmap
import time

import torch  # loads libcudart

rt = C.CDLL("libcudart.so.12")
rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
rt.cudaHostRegister.argtypes = [C.c_void_p, C.c_size_t, C.c_uint]
rt.cudaHostUnregister.argtypes = [C.c_void_p]
rt.cudaFreeHost.argtypes = [C.c_void_p]
rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
rt.cudaFree.argtypes = [C.c_void_p]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaStreamCreate.argtypes = [C.POINTER(C.c_void_p)]
rt.cudaStreamSynchronize.argtypes = [C.c_void_p]
libc = C.CDLL("libc.so.6", use_errno=True)
libc.mmap.restype = C.c_void_p
libc.mmap.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_long]
libc.madvise.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
libc.munmap.argtypes = [C.c_void_p, C.c_size_t]
NB = 1 << 30
D2H, H2D = 2, 1
MAP_HUGETLB, MADV_HUGEPAGE = 0x40000, 14


def host_buffer(kind):
    p = C.c_void_p()
    if kind == "pinned":
        assert rt.cudaHostAlloc(C.byref(p), NB, 0) == 0
        return p.value, lambda: rt.cudaFreeHost(p)
    if kind == "write-combined":
        assert rt.cudaHostAlloc(C.byref(p), NB, 4) == 0
        return p.value, lambda: rt.cudaFreeHost(p)
    flags = mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS | (MAP_HUGETLB if kind == "hugetlb+register" else 0)
    a = libc.mmap(None, NB, mmap.PROT_READ | mmap.PROT_WRITE, flags, -1, 0)
    if a is None or a == C.c_void_p(-1).value:
        return None, None
    if kind == "thp+register":
        libc.madvise(a, NB, MADV_HUGEPAGE)
    C.memset(a, 1, NB)                                  # touch
    if rt.cudaHostRegister(a, NB, 0) != 0:
        libc.munmap(a, NB)
        return None, None
    return a, lambda: (rt.cudaHostUnregister(a), libc.munmap(a, NB))


def main():
    ngpu = torch.cuda.device_count()
    print(f"{ngpu} GPUs, 1 GiB per copy, best of 3")
    for kind in ("pinned", "write-combined", "thp+register", "hugetlb+register"):
        for devs in sorted({1, min(2, ngpu), min(4, ngpu), ngpu}):
            bufs = []
            ok = True
            for d in range(devs):
                rt.cudaSetDevice(d)
                h, free = host_buffer(kind)
                if h is None:
                    ok = False
                    break
                dp, st = C.c_void_p(), C.c_void_p()
                assert rt.cudaMalloc(C.byref(dp), NB) == 0 and rt.cudaStreamCreate(C.byref(st)) == 0
                bufs.append((d, h, free, dp, st))
            if not ok:
                print(f"{kind:18s} {devs} GPU(s): not available on this box")
                for d, h, free, dp, st in bufs:
                    free(); rt.cudaFree(dp)
                break
            res = {}
            for name, direction in (("d2h", D2H), ("h2d", H2D)):
                best = 0.0
                for rep in range(3):
                    for d, h, free, dp, st in bufs:
                        rt.cudaSetDevice(d); rt.cudaStreamSynchronize(st)
                    t0 = time.perf_counter()
                    for d, h, free, dp, st in bufs:
                        rt.cudaSetDevice(d)
                        if direction == D2H:
                            rt.cudaMemcpyAsync(h, dp, NB, D2H, st)
                        else:
                            rt.cudaMemcpyAsync(dp, h, NB, H2D, st)
                    for d, h, free, dp, st in bufs:
                        rt.cudaSetDevice(d); rt.cudaStreamSynchronize(st)
                    best = max(best, devs * NB / (time.perf_counter() - t0) / 1e9)
                res[name] = best
            print(f"{kind:18s} {devs} GPU(s): aggregate D2H {res['d2h']:7.1f} GB/s   H2D {res['h2d']:7.1f} GB/s")
            for d, h, free, dp, st in bufs:
                rt.cudaSetDevice(d); free(); rt.cudaFree(dp)


if __name__ == "__main__":
    main()
