"""Copy-engine probe: do a pinned H2D and a pinned D2H on two streams overlap on this box?"""
import torch, time
p = torch.cuda.get_device_properties(0)
print("device", p.name, "multi_processor_count", p.multi_processor_count)
try:
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    v = ctypes.c_int(0)
    rt.cudaDeviceGetAttribute(ctypes.byref(v), 40, 0)   # cudaDevAttrAsyncEngineCount
    print("asyncEngineCount", v.value)
except Exception as e:
    print("attr query failed", e)
n = 1 << 29
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(both):
    torch.cuda.synchronize(); t = time.perf_counter()
    with torch.cuda.stream(s1):
        h1.copy_(d1, non_blocking=True)
    if both:
        with torch.cuda.stream(s2):
            d2.copy_(h2, non_blocking=True)
    torch.cuda.synchronize(); return time.perf_counter() - t
for _ in range(2): run(True)
a = min(run(False) for _ in range(3)); b = min(run(True) for _ in range(3))
print("D2H alone %.1f GB/s; D2H + H2D together: %.2f ms vs %.2f ms alone -> %s" % (n / a / 1e9, b * 1e3, a * 1e3, "overlap" if b < 1.5 * a else "SERIALISED"))
