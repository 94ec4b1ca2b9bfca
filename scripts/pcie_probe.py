"""What slows a pinned device->host copy down on this box?  13 copies of 190 MB back to back (the copy-out of one config-2
step), alone and with the things the one-call path runs beside it."""
import threading, time
import numpy as np, torch

N, SZ = 13, 190 << 20
dev = torch.empty(SZ, dtype=torch.uint8, device="cuda")
host = torch.empty(N * SZ, dtype=torch.uint8).pin_memory()
h_in = torch.empty(24 << 20, dtype=torch.uint8).pin_memory()
d_in = torch.empty(24 << 20, dtype=torch.uint8, device="cuda")
big_a = torch.empty(1 << 28, dtype=torch.float32, device="cuda"); big_b = torch.empty_like(big_a)
s_out, s_in, s_k = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
stop = False

def host_traffic():
    a = np.empty(24 << 20, dtype=np.uint8); b = np.empty_like(a)
    while not stop:
        np.copyto(b, a)

def run(h2d=False, kernels=False, threads=0):
    global stop
    stop = False
    ts = [threading.Thread(target=host_traffic) for _ in range(threads)]
    for t in ts: t.start()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(N):
        if h2d:
            with torch.cuda.stream(s_in): d_in.copy_(h_in, non_blocking=True)
        if kernels:
            with torch.cuda.stream(s_k):
                for _ in range(2): big_b.copy_(big_a)          # ~1.3 ms of HBM-bound work
        with torch.cuda.stream(s_out): host[i * SZ:(i + 1) * SZ].copy_(dev, non_blocking=True)
    s_out.synchronize(); dt = time.perf_counter() - t0
    torch.cuda.synchronize(); stop = True
    for t in ts: t.join()
    return N * SZ / dt / 1e9

for _ in range(2): run()
print("D2H alone                      %.1f GB/s" % max(run() for _ in range(3)))
print("+ H2D 24 MB per copy           %.1f GB/s" % max(run(h2d=True) for _ in range(3)))
print("+ kernels (HBM-bound)          %.1f GB/s" % max(run(kernels=True) for _ in range(3)))
print("+ H2D + kernels                %.1f GB/s" % max(run(h2d=True, kernels=True) for _ in range(3)))
for th in (1, 2, 4, 8):
    print("+ %d host memcpy threads        %.1f GB/s" % (th, max(run(threads=th) for _ in range(2))))
print("+ H2D + kernels + 4 threads    %.1f GB/s" % max(run(h2d=True, kernels=True, threads=4) for _ in range(2)))
