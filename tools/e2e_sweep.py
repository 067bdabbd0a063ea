#!/usr/bin/env python
"""Chunk-size sweep of the host-buffer path (wfb_process_host): records/s and H2D GB/s per chunk size."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from waveformanalysis_b200 import engine  # noqa: E402
from waveformanalysis_b200.dtypes import BASIC_FEATURES_DTYPE, RECORDS_DTYPE, THRESHOLD_HIT_DTYPE  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--records", type=int, default=4_000_000)
ap.add_argument("--chunks", default="131072,262144,524288,1048576")
args = ap.parse_args()
n, L = args.records, 800
dev = engine.DeviceRun.synth(n, L, 16, seed=77, with_rows=True)
pool_pin = torch.empty(n * L, dtype=torch.int16).pin_memory()
rows_pin = torch.empty(n * 102, dtype=torch.uint8).pin_memory()
pool_pin.copy_(dev.pool)
rows_pin.copy_(dev.records_rows)
torch.cuda.synchronize()
del dev
torch.cuda.empty_cache()
records = rows_pin.numpy().view(RECORDS_DTYPE)
pool = pool_pin.numpy().view(np.uint16)
feat = torch.empty(n * 36, dtype=torch.uint8).pin_memory().numpy().view(BASIC_FEATURES_DTYPE)
first = engine.process_host(records, pool, threshold=15.0, out_features=feat)
hits = torch.empty((first["n_hits"] + 1024) * 60, dtype=torch.uint8).pin_memory().numpy().view(THRESHOLD_HIT_DTYPE)
for c in [int(x) for x in args.chunks.split(",")]:
    engine.process_host(records, pool, threshold=15.0, out_features=feat, out_hits=hits, chunk_records=c)
    t0 = time.perf_counter()
    for _ in range(3):
        engine.process_host(records, pool, threshold=15.0, out_features=feat, out_hits=hits, chunk_records=c)
    dt = (time.perf_counter() - t0) / 3
    print(f"chunk={c:8d} records/s={n / dt / 1e6:7.2f} M  H2D={n * (2 * L + 102) / dt / 1e9:6.2f} GB/s")
