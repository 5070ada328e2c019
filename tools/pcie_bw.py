#!/usr/bin/env python3
"""Development helper: pinned-memory PCIe bandwidth on this box (the floor of the e2e figure)."""
import time
import torch
n = 67_240_000 // 8
h = torch.empty(n, dtype=torch.float64).pin_memory()
d = torch.empty(n, dtype=torch.float64, device="cuda")
h2 = torch.empty(32_000_000 // 8, dtype=torch.float64).pin_memory()
d2 = torch.empty(32_000_000 // 8, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for name, fn in (("D2H 67 MB", lambda: h.copy_(d, non_blocking=True)),
                 ("H2D 32 MB", lambda: d2.copy_(h2, non_blocking=True))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    nb = 67.24e6 if "D2H" in name else 32e6
    print("%s: %.3f ms  %.1f GB/s" % (name, dt * 1e3, nb / dt / 1e9))
# both directions at once
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize()
print("both: %.3f ms per pair" % ((time.perf_counter() - t0) / 10 * 1e3))
# small copies (2.4 MB, like one output field of one chunk)
hs = torch.empty(2_400_000 // 8, dtype=torch.float64).pin_memory()
dsm = torch.empty(2_400_000 // 8, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(28): hs.copy_(dsm, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("28 x 2.4 MB D2H: %.3f ms  %.1f GB/s" % (dt * 1e3, 28 * 2.4e6 / dt / 1e9))
