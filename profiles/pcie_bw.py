"""Pinned D2H / H2D bandwidth of the box (what bounds bench.py's e2e): python profiles/pcie_bw.py"""
import torch, time
n = 482344960
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
for name, fn in (("D2H one copy", lambda: h.copy_(d, non_blocking=True)),
                 ("H2D one copy", lambda: d.copy_(h, non_blocking=True))):
    for _ in range(2):
        fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("%s: %.2f ms  %.1f GB/s" % (name, ms, n / ms / 1e6))
# two streams, halves in parallel
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
t = time.perf_counter()
with torch.cuda.stream(s1):
    h[: n // 2].copy_(d[: n // 2], non_blocking=True)
with torch.cuda.stream(s2):
    h[n // 2:].copy_(d[n // 2:], non_blocking=True)
torch.cuda.synchronize()
ms = (time.perf_counter() - t) * 1e3
print("D2H two streams: %.2f ms  %.1f GB/s" % (ms, n / ms / 1e6))
