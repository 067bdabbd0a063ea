#!/usr/bin/env python
"""Device-resident timing of the filter / width kernels at BASELINE configs[2] scale (diagnostics and ncu target)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from waveformanalysis_b200 import _lib, engine, ops

ap = argparse.ArgumentParser()
ap.add_argument("--records", type=int, default=2_000_000)
ap.add_argument("--samples", type=int, default=800)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
n, L = args.records, args.samples
run = engine.DeviceRun.synth(n, L, 64, dt_ns=4, seed=303)
torch.cuda.synchronize()


def timed(label, fn, bytes_):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    print(f"{label:28s} {ms:9.3f} ms  {bytes_ / ms / 1e6:8.1f} GB/s (algorithmic)  {n / ms / 1e3:8.1f} M records/s")


sos = ops.butter_bandpass_sos(4, 0.01, 0.1, 0.25)
timed("sg_filter 11/2", lambda: ops.filter_run_device(run, {"filter_type": "SG", "sg_window_size": 11, "sg_poly_order": 2}), n * L * 6)
timed("bw_filter order 4", lambda: ops.filter_run_device(run, {"filter_type": "BW", "sos": sos}), n * L * 6)
lib = _lib.load()
out = torch.empty(n * 52, dtype=torch.uint8, device="cuda")
timed("width_integral", lambda: _lib.check(lib.wfb_width_integral(engine._ptr(run.pool), 0, run.pool_len, engine._ptr(run.meta), n, 0.1, 0.9, 4.0, 0, 0,
                                                                  engine._ptr(out), engine._stream()), "wi"), n * (2 * L + 100))
