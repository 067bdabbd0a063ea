#!/usr/bin/env python
"""Device-resident timing of wfb_find_peaks at BASELINE configs[2] scale on an SG-filtered float32 pool (ncu target)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from waveformanalysis_b200 import _lib, engine, ops
from waveformanalysis_b200.dtypes import HIT_DTYPE

n, L = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, 800
run = engine.DeviceRun.synth(n, L, 64, dt_ns=4, seed=303)
run.pool.copy_(16383 - run.pool)
meta = run.meta.view(torch.uint8).view(-1, 48)
base = meta[:, 8:16].contiguous().view(torch.float64).view(-1)
meta[:, 8:16] = (16383.0 - base).view(torch.uint8).view(-1, 8)
meta[:, 36] = 1
d_sg = ops.filter_run_device(run, {"filter_type": "SG", "sg_window_size": 11, "sg_poly_order": 2})
frun = engine.DeviceRun(run.meta, d_sg, n, 1, L)
lib = _lib.load()
p = _lib.PeakParams(wave_kind=_lib.WAVE_REC_F32, use_derivative=1, height=8.0, prominence=0.7, width=2.0, threshold=0.0, has_threshold=0,
                    distance=2, height_method=0, height_window_extension=4, lmax=L, level_f32=0)
ws = torch.empty(lib.wfb_find_peaks_workspace_bytes(n), dtype=torch.uint8, device="cuda")
total = torch.zeros(1, dtype=torch.int64, device="cuda")
cap = 8 * n
rows = torch.empty(cap * HIT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")


def peaks():
    _lib.check(lib.wfb_find_peaks(engine._ptr(frun.pool), frun.pool_len, engine._ptr(frun.meta), n, C.byref(p), engine._ptr(rows), cap,
                                  C.c_void_p(0), engine._ptr(total), engine._ptr(ws), ws.numel(), engine._stream()), "wfb_find_peaks")


peaks()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    peaks()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"find_peaks {n} records x {L}: {ms:.3f} ms  {n / ms / 1e3:.1f} M records/s  peaks/record {int(total.item()) / n:.2f}")
