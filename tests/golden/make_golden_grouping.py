#!/usr/bin/env python
"""Golden vectors for hit_grouped on chain-merged hits whose clusters span several records (their sample
window is -1 and the absolute window comes from the component hits, event_grouping.py:369-414): the LIVE
reference HitGroupedPlugin on the m50 rows of hotpath_golden.npz.

    python tests/golden/make_golden_grouping.py   # rewrites tests/golden/grouping_golden.npz"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

from make_golden import Ctx, import_reference  # noqa: E402


def main():
    import_reference()
    from waveform_analysis.core.plugins.builtin.cpu.event_analysis import HitGroupedPlugin

    g = np.load(os.path.join(HERE, "hotpath_golden.npz"))
    h, mg, cp = g["hits_thr15"], g["m50_merged"], g["m50_components"]
    assert (mg["sample_start"] < 0).any()
    G = {}
    for wname, w in (("w100", 100.0), ("w0", 0.0), ("w5000", 5000.0)):
        df = HitGroupedPlugin().compute(Ctx({"time_window_ns": w}, {"hit_merged": mg, "hit_merged_components": cp, "hit_threshold": h}), "run")
        G[f"hg50_{wname}_t_min"] = df["t_min"].to_numpy(np.int64)
        G[f"hg50_{wname}_t_max"] = df["t_max"].to_numpy(np.int64)
        G[f"hg50_{wname}_n_hits"] = df["n_hits"].to_numpy(np.int64)
        G[f"hg50_{wname}_record_ids"] = np.concatenate([np.asarray(v, np.int64) for v in df["record_ids"]])
        G[f"hg50_{wname}_timestamps"] = np.concatenate([np.asarray(v, np.int64) for v in df["timestamps"]])
        G[f"hg50_{wname}_channels"] = np.concatenate([np.asarray(v, np.int64) for v in df["channels"]])
        G[f"hg50_{wname}_sample_starts"] = np.concatenate([np.asarray(v, np.int64) for v in df["sample_starts"]])
    out = os.path.join(HERE, "grouping_golden.npz")
    np.savez_compressed(out, **G)
    print("wrote", out, {k: v.shape for k, v in G.items() if k.endswith("t_min")})


if __name__ == "__main__":
    main()
