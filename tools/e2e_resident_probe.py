#!/usr/bin/env python
"""Diagnostics: the chunked host pipeline (plain / pinned results / resident) and the plugin-level step, 4 M records."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from waveformanalysis_b200 import engine, residency
from waveformanalysis_b200.dtypes import RECORDS_DTYPE

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
    dt = time.perf_counter() - t0
    print(f"{label:50s} {1e3 * dt:9.2f} ms  {n / dt / 1e6:7.2f} M rec/s", flush=True)
    return out


for rep in range(int(os.environ.get("PROBE_HOST_PASSES", "3"))):
    print("--- pass", rep)
    for ck in (1 << 17, 1 << 18, 1 << 19):
        r = T(f"process_host pageable results chunk={ck}", lambda: engine.process_host(records, pool, threshold=15.0, chunk_records=ck))
        del r
    r = T("process_host pinned results", lambda: engine.process_host(records, pool, threshold=15.0, pinned_results=True))
    del r
    r = T("process_host resident + pinned", lambda: engine.process_host(records, pool, threshold=15.0, pinned_results=True, keep_resident=True))
    del r
    r = T("process_host resident + pinned, cap 6.5n", lambda: engine.process_host(records, pool, threshold=15.0, pinned_results=True, keep_resident=True, hit_cap=int(6.5 * n)))
    del r

from waveformanalysis_b200.plugins import B200BasicFeaturesPlugin, B200RecordsPlugin, B200ThresholdHitPlugin, B200WavePoolPlugin

plugins = {"records": B200RecordsPlugin(), "wave_pool": B200WavePoolPlugin(), "basic_features": B200BasicFeaturesPlugin(),
           "hit_threshold": B200ThresholdHitPlugin()}
ctx = bench.PluginContext({"wave_source": "records", "hit_threshold": {"threshold": 15.0}}, plugins)
K = [0]


def step():
    K[0] += 1
    run_id = f"e2e_{K[0]}"
    ctx._results[(run_id, "records")] = records
    ctx._results[(run_id, "wave_pool")] = pool
    feats = plugins["basic_features"].compute(ctx, run_id)
    hits = plugins["hit_threshold"].compute(ctx, run_id)
    residency.release(run_id)
    ctx._results.clear()
    return feats, hits


for rep in range(4):
    out = T("plugin step", step)
    del out
pr = cProfile.Profile()
pr.enable()
out = step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)

# ---- the streaming backend over the same host arrays
from waveformanalysis_b200.plugins import B200HitThresholdStreamPlugin

ctx.config["hit_threshold_stream"] = {"threshold": 15.0}
ctx._results[("stream", "records")] = records
ctx._results[("stream", "wave_pool")] = pool
sp = B200HitThresholdStreamPlugin()


def stream_pass():
    rows = 0
    for chunk in sp.compute(ctx, "stream"):
        rows += len(chunk.data)
    return rows


def host_allocs():
    try:
        st = torch.cuda.host_memory_stats()
        return {k: st[k] for k in st if ("num_host_alloc" in k or "host_alloc_time" in k or k.endswith("allocated_bytes.allocated") or k.endswith("segment.allocated")) }
    except Exception as exc:  # noqa: BLE001
        return str(exc)


for rep in range(6):
    T("stream pass", stream_pass)
    print("   pinned allocator:", host_allocs(), sp.stream_stats)
pr = cProfile.Profile()
pr.enable()
stream_pass()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(30)
