#!/usr/bin/env python
"""Run every non-fused hot-path op once on a synthetic run (kernel timings come from an ncu launch list).

  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python tools/opbench.py
  python tools/opbench.py --summarise launches.csv      # kernel -> ms, algorithmic GB/s

Sizes are printed as JSON on the last line so the summary can turn durations into GB/s."""
import argparse
import collections
import csv
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--records", type=int, default=262144)
ap.add_argument("--samples", type=int, default=800)
ap.add_argument("--summarise", default=None)
ap.add_argument("--sizes", default=None, help="sizes JSON printed by the run (for --summarise)")
args = ap.parse_args()

if args.summarise:
    sizes = json.loads(args.sizes) if args.sizes else {}
    n, L, H = sizes.get("records", args.records), sizes.get("samples", args.samples), sizes.get("hits", 0)
    by = collections.OrderedDict()
    for r in csv.reader(open(args.summarise)):
        if len(r) > 14 and r[0].isdigit():
            by.setdefault(r[4].split("(")[0].replace("void ", "").replace("wfb::", ""), []).append(float(r[14]) / 1e6)
    # algorithmic bytes per launch; None: bookkeeping kernels
    alg = {
        "sg_filter_kernel": n * 6 * L, "bw_filter_kernel": n * 6 * L, "k1_gather_kernel": n * (4 * L + 126),
        "width_integral_kernel": n * (2 * L + 100), "waveform_width_kernel": None,
        "find_peaks_kernel<0>": n * 4 * L, "find_peaks_kernel<1>": None, "peaks_emit_cached_kernel": None,
    }
    peak = 6453.7
    print(f"{'kernel':44s} {'launches':>8s} {'total ms':>10s} {'alg GB/s':>10s} {'% of HBM peak':>14s}")
    for k, v in by.items():
        b = next((alg[a] for a in alg if k.startswith(a)), None)
        gbs = (b * len(v) / (sum(v) * 1e-3) / 1e9) if b else None
        print(f"{k[:44]:44s} {len(v):8d} {sum(v):10.3f} {gbs if gbs is None else round(gbs, 1)!s:>10s} {'' if gbs is None else round(100 * gbs / peak, 1)!s:>14s}")
    sys.exit(0)

import numpy as np
import torch

from waveformanalysis_b200 import engine, ops
from waveformanalysis_b200.dtypes import HIT_DTYPE

n, L = args.records, args.samples
run = engine.DeviceRun.synth(n, L, 16, seed=4321, with_rows=True)
rec = run.records_to_host()
pool = run.pool_to_host()
torch.cuda.synchronize()

# K1: rebuild the records from "raw" rows in shuffled order (time sort + gather + baseline)
rng = np.random.default_rng(7)
perm = rng.permutation(n)
raw = pool.view(np.int16).reshape(n, L)[perm]
rec2, pool2 = ops.build_records(rec["timestamp"][perm], rec["board"][perm], rec["channel"][perm], raw, dt_ns=2)
assert np.array_equal(pool2, pool), "K1 round trip"

# wave_pool_filtered: SG 11/2 and Butterworth order 4
sg = ops.filter_pool(rec, pool, configs={}, default={"filter_type": "SG", "sg_window_size": 11, "sg_poly_order": 2})
bw = ops.filter_pool(rec, pool, configs={}, default={"filter_type": "BW", "sos": ops.butter_bandpass_sos(4, 0.01, 0.1, 0.5)})

# hits -> merge -> grouping, widths
out = engine.process_host(rec, pool, threshold=15.0)
hits = out["hits"]
cl, mg, cp = ops.hit_merge_default(hits)
ev = ops.group_hit_windows(mg, 100.0)
ev2 = ops.group_time_window(out["features"]["timestamp"], out["features"]["channel"], 100.0)
wi = ops.width_integral(rec, pool)
h2 = np.zeros(len(hits), dtype=HIT_DTYPE)
for f in ("position", "height", "integral", "edge_start", "edge_end", "dt", "timestamp", "board", "channel", "record_id"):
    if f in h2.dtype.names and f in hits.dtype.names:
        h2[f] = hits[f]
# positive-going copy of the waves so waveform_width keeps its rows (it drops peaks below the baseline)
waves = (16383 - pool.view(np.int16).reshape(n, L)).astype(np.int16)
ww = ops.waveform_width(h2, rec["record_id"], waves)
# hit = find_peaks on the SG-filtered pool (records source) and on the raw rows
pk = ops.find_peaks_records(rec, sg, height=8.0, width=2)
pk2 = ops.find_peaks_records(rec, pool, height=12.0, width=2)
torch.cuda.synchronize()
print(json.dumps({"records": n, "samples": L, "hits": int(len(hits)), "width_rows": int(len(ww)), "peaks": int(len(pk)), "events": int(len(ev["event_id"])) if isinstance(ev, dict) and "event_id" in ev else None}))
