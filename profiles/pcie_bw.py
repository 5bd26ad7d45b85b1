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
# chunked like flyby_render's copy-out: 8 chunks x (features, ids) on one stream
feat_d = torch.empty(92 * 512 * 512 * 4, dtype=torch.float32, device="cuda")
ids_d = torch.empty(92 * 512 * 512, dtype=torch.int32, device="cuda")
feat_h = torch.empty(92 * 512 * 512 * 4, dtype=torch.float32).pin_memory()
ids_h = torch.empty(92 * 512 * 512, dtype=torch.int32).pin_memory()
s = torch.cuda.Stream()
for trial in range(2):
    torch.cuda.synchronize()
    t = time.perf_counter()
    with torch.cuda.stream(s):
        for k in range(8):
            a, b = feat_d.numel() * k // 8, feat_d.numel() * (k + 1) // 8
            feat_h[a:b].copy_(feat_d[a:b], non_blocking=True)
            a, b = ids_d.numel() * k // 8, ids_d.numel() * (k + 1) // 8
            ids_h[a:b].copy_(ids_d[a:b], non_blocking=True)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t) * 1e3
print("D2H 16 chunked copies: %.2f ms  %.1f GB/s" % (ms, n / ms / 1e6))
# the same while another stream streams through DRAM (stands in for the render kernels)
big = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
s2 = torch.cuda.Stream()
for load in (0, 1):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if load:
        with torch.cuda.stream(s2):
            for _ in range(60):   # ~20 ms of DRAM-bound kernels
                big.add_(1)
    with torch.cuda.stream(s):
        e0.record(s)
        for k in range(8):
            a, b = feat_d.numel() * k // 8, feat_d.numel() * (k + 1) // 8
            feat_h[a:b].copy_(feat_d[a:b], non_blocking=True)
            a, b = ids_d.numel() * k // 8, ids_d.numel() * (k + 1) // 8
            ids_h[a:b].copy_(ids_d[a:b], non_blocking=True)
        e1.record(s)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("D2H chunked, DRAM-bound kernels on another stream: %s  %.2f ms  %.1f GB/s" % ("yes" if load else "no", ms, n / ms / 1e6))
