#!/usr/bin/env python
"""Pinned host <-> device copy bandwidth of this box (the ceiling of the e2e number)."""
import torch

n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = timed(lambda: d.copy_(h, non_blocking=True)); print(f"H2D 1 GiB: {n/ms/1e6:.1f} GB/s")
ms = timed(lambda: h2.copy_(d2, non_blocking=True)); print(f"D2H 1 GiB: {n/ms/1e6:.1f} GB/s")
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
ms = timed(both); print(f"H2D + D2H concurrently: {n/ms/1e6:.1f} GB/s each direction")
