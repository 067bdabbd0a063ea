#!/usr/bin/env python
"""Quick kernel timing of the fused pass variants (diagnostics, not the bench contract)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from waveformanalysis_b200 import engine

ap = argparse.ArgumentParser()
ap.add_argument("--records", type=int, default=2_000_000)
ap.add_argument("--samples", type=int, default=800)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--modes", default="both,features,hits")
ap.add_argument("--threshold", type=float, default=15.0)
args = ap.parse_args()

run = engine.DeviceRun.synth(args.records, args.samples, 16, seed=1235)
torch.cuda.synchronize()
for mode in args.modes.split(","):
    kw = dict(features=mode in ("both", "features"), hits=mode in ("both", "hits"), threshold=args.threshold)
    res = run.features_hits(hit_cap=1024, **kw)
    torch.cuda.synchronize()
    nh = int(res["total"].item()) if kw["hits"] else 0
    out = {"features": torch.empty(args.records * 36, dtype=torch.uint8, device="cuda"),
           "hits": torch.empty((nh + 1024) * 60, dtype=torch.uint8, device="cuda"),
           "total": torch.zeros(1, dtype=torch.int64, device="cuda")}
    for _ in range(2):
        run.features_hits(out=out, hit_cap=nh + 1024, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        run.features_hits(out=out, hit_cap=nh + 1024, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    bpr = 2 * args.samples + 72 + 60 * nh / args.records
    print(f"mode={mode:9s} variant={os.environ.get('WFB_FUSED_VARIANT','auto'):7s} records={args.records} L={args.samples} hits/rec={nh/args.records:.2f} "
          f"ms={ms:.3f} Mrec/s={args.records/ms/1e3:.1f} GB/s(alg)={args.records*bpr/ms/1e6:.1f}")
