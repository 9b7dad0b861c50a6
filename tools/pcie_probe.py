#!/usr/bin/env python
"""What the PCIe link of this box gives a pinned device->host / host->device copy (GB/s), alone and both directions at once:
the ceiling of the host-buffer C-ABI call (bench.py `e2e`), which moves 46 B out and 16 B in per env-step."""
import json
import time

import torch

dev = torch.device("cuda", 0)
out = {}
for mb in (5, 40, 256):
    n = mb << 20
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        out[f"{name}_{mb}MB_GBs"] = round(10 * n / (time.perf_counter() - t0) / 1e9, 2)
# both directions at once on two streams (40 MB out, 16 MB in: the shape of one e2e step)
d1 = torch.empty(40 << 20, dtype=torch.uint8, device=dev); h1 = torch.empty(40 << 20, dtype=torch.uint8).pin_memory()
d2 = torch.empty(16 << 20, dtype=torch.uint8, device=dev); h2 = torch.empty(16 << 20, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1):
        h1.copy_(d1, non_blocking=True)
    with torch.cuda.stream(s2):
        d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
out["duplex_40MB_out_16MB_in_ms_per_pair"] = round(dt / 10 * 1e3, 3)
out["duplex_d2h_GBs"] = round(10 * (40 << 20) / dt / 1e9, 2)
print(json.dumps(out))
