#!/usr/bin/env python
"""Where the time of the plugin-level host path goes (diagnostics): 4 M records x 800 samples, pinned host inputs."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from waveformanalysis_b200 import engine, residency
from waveformanalysis_b200.dtypes import BASIC_FEATURES_DTYPE, RECORDS_DTYPE, THRESHOLD_HIT_DTYPE

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
dev = engine.DeviceRun.synth(n, 800, 16, seed=77, with_rows=True)
pool_pin = torch.empty(n * 800, dtype=torch.int16).pin_memory()
rows_pin = torch.empty(n * 102, dtype=torch.uint8).pin_memory()
pool_pin.copy_(dev.pool)
rows_pin.copy_(dev.records_rows)
torch.cuda.synchronize()
del dev
records = rows_pin.numpy().view(RECORDS_DTYPE)
pool = pool_pin.numpy().view(np.uint16)


def T(label, fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    print(f"{label:40s} {1e3 * (time.perf_counter() - t0):9.2f} ms")
    return out


for rep in range(2):
    print("--- pass", rep)
    T("fingerprint", lambda: residency.fingerprint(records, pool))
    T("dt min/max", lambda: (np.ascontiguousarray(records["dt"]).min(), records["event_length"].max()))
    run = T("DeviceRun.from_host (upload)", lambda: engine.DeviceRun.from_host(records, pool))
    res = T("features_hits kernel (cap 8n)", lambda: run.features_hits(threshold=15.0))
    total = int(res["total"].item())
    feats = T("features D2H (.cpu().numpy())", lambda: res["features"][: n * 36].cpu().numpy().view(BASIC_FEATURES_DTYPE))
    hits = T("hits D2H (.cpu().numpy())", lambda: res["hits"][: total * 60].cpu().numpy().view(THRESHOLD_HIT_DTYPE))
    T("hits D2H into np.empty via cudaMemcpy", lambda: torch.from_numpy(np.empty(total * 60, np.uint8)).copy_(res["hits"][: total * 60]))
    pin = T("pinned alloc of the hit rows", lambda: torch.empty(total * 60, dtype=torch.uint8).pin_memory())
    T("hits D2H into pinned", lambda: pin.copy_(res["hits"][: total * 60]))
    T("run_to_host (kernel + both D2H)", lambda: run.run_to_host(threshold=15.0))
    del run, res
    T("empty_cache", lambda: torch.cuda.empty_cache())
