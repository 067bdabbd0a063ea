#!/usr/bin/env python
"""Diagnostics: run the golden hit cases through both C-ABI paths and print the first differing rows."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from waveformanalysis_b200 import engine as eng

g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "hotpath_golden.npz"), allow_pickle=False)
rec, pool = g["records"], g["wave_pool"]


def show(name, got, want):
    n = min(len(got), len(want))
    print(f"{name}: got {len(got)} rows, want {len(want)}")
    bad = [i for i in range(n) if got[i].tobytes() != want[i].tobytes()]
    print(f"  {len(bad)} differing rows; first: {bad[:5]}")
    for i in bad[:6]:
        print("   got ", got[i])
        print("   want", want[i])
        rid = int(want[i]["record_id"])
        r = rec[rid]
        print("   record", rid, "len", int(r["event_length"]), "off", int(r["wave_offset"]), "mis", int(r["wave_offset"]) % 8, "baseline", float(r["baseline"]))


for kw, key in ((dict(threshold=15.0), "hits_thr15"),
                (dict(threshold=12.0, thresholds={(0, 2): 40.0}, left_extension=5, right_extension=0), "hits_chan")):
    a = eng.process_host(rec, pool, features=False, **kw)
    kw2 = {k: v for k, v in kw.items() if k != "thresholds"}
    b = eng.DeviceRun.from_host(rec, pool).run_to_host(rules=eng.make_rules(kw.get("thresholds"), None), features=False, **kw2)
    show(key + " host-path", a["hits"], g[key])
    show(key + " device-path", b["hits"], g[key])
