#!/usr/bin/env python
"""Timing of the fused pass on a float32 pool (wave_pool_filtered): the warp-per-record kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from waveformanalysis_b200 import engine

n, L = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000, 800
run = engine.DeviceRun.synth(n, L, 16, seed=1235)
pool_f32 = run.pool.view(torch.int16).to(torch.float32)  # same values as float32 samples
frun = engine.DeviceRun(run.meta, pool_f32, n, 1, L)
torch.cuda.synchronize()
for mode in ("both", "features", "hits"):
    kw = dict(features=mode in ("both", "features"), hits=mode in ("both", "hits"), threshold=15.0)
    res = frun.features_hits(hit_cap=1024, **kw)
    torch.cuda.synchronize()
    nh = int(res["total"].item()) if kw["hits"] else 0
    out = {"features": torch.empty(n * 36, dtype=torch.uint8, device="cuda"), "hits": torch.empty((nh + 1024) * 60, dtype=torch.uint8, device="cuda"),
           "total": torch.zeros(1, dtype=torch.int64, device="cuda")}
    for _ in range(2):
        frun.features_hits(out=out, hit_cap=nh + 1024, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        frun.features_hits(out=out, hit_cap=nh + 1024, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    bpr = 4 * L + 72 + 60 * nh / n
    print(f"f32 pool mode={mode:9s} records={n} hits/rec={nh/n:.2f} ms={ms:.3f} Mrec/s={n/ms/1e3:.1f} GB/s(alg)={n*bpr/ms/1e6:.1f}")
