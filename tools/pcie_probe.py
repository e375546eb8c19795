"""Measure pinned H2D / D2H bandwidth and their overlap on this box (context for the e2e number)."""
import time
import torch
dev = torch.device("cuda", 0)
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device=dev)
d2 = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
a = t(lambda: d.copy_(h, non_blocking=True))
b = t(lambda: h2.copy_(d2, non_blocking=True))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = t(both)
print({"h2d_GBps": n / a / 1e9, "d2h_GBps": n / b / 1e9, "both_concurrent_ms": c * 1e3, "h2d_ms": a * 1e3, "d2h_ms": b * 1e3})
