#!/usr/bin/env python
"""Golden vectors for the step after the hot path - df, df_events -> df_paired, s1_s2 - from the LIVE reference
plugins (DataFramePlugin, GroupedEventsPlugin + EventAnalyzer.pair_events, S1S2ClassifierPlugin) on rows of
hotpath_golden.npz.

    python tests/golden/make_golden_after.py   # rewrites tests/golden/after_golden.npz"""

from __future__ import annotations

import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

from make_golden import Ctx, import_reference  # noqa: E402

S1S2_CASES = {
    "none": {},
    "width": {"s1_width_range": (None, 420.0), "s2_width_range": (420.0, None)},
    "samples_conflict": {"width_unit": "samples", "s1_width_range": (75.0, 230.0), "s2_width_range": (200.0, 600.0),
                         "s1_height_range": (6.0, None), "conflict_policy": "prefer_s2"},
    "area_prefer_s1": {"s1_area_range": (0.0, 25000.0), "s2_area_range": (15000.0, None), "s2_height_range": (None, 7.5),
                       "conflict_policy": "prefer_s1"},
    "conflict_unknown": {"s1_width_range": (0.0, 500.0), "s2_width_range": (400.0, 1200.0)},
    "only_s2_height": {"s2_height_range": (5.0, 8.0)},
}


def df_to_npz(G, prefix, df):
    G[f"{prefix}_index"] = df.index.to_numpy(np.int64)
    for c in df.columns:
        G[f"{prefix}_{c}"] = df[c].to_numpy()


def main():
    import_reference()
    from waveform_analysis.core.plugins.builtin.cpu.dataframe import DataFramePlugin
    from waveform_analysis.core.plugins.builtin.cpu.event_analysis import GroupedEventsPlugin
    from waveform_analysis.core.plugins.builtin.cpu.s1_s2_classifier import S1S2ClassifierPlugin
    from waveform_analysis.core.processing.analyzer import EventAnalyzer

    g = np.load(os.path.join(HERE, "hotpath_golden.npz"))
    rng = np.random.default_rng(20261018)
    records, bf = g["records"], g["bf_default"]
    assert len(np.unique(records["timestamp"])) == len(records)  # pandas' quicksort: only unique keys have a defined order
    perm = rng.permutation(len(records))
    records, bf = records[perm].copy(), bf[perm].copy()
    bf["timestamp"], bf["board"], bf["channel"] = records["timestamp"], records["board"], records["channel"]
    G = {"df_in_records": records, "df_in_features": bf}

    # df, records mode, no calibration / calibration with a missing and an invalid channel
    cfg = {"wave_source": "records"}
    df = DataFramePlugin().compute(Ctx(cfg, {"records": records, "basic_features": bf}), "run")
    df_to_npz(G, "df_plain", df)
    gains = {"0:0": 12.5, "0:1": 13.2, "0:2": {"gain_adc_per_pe": 7.0}, "0:3": -1.0}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        df_pe = DataFramePlugin().compute(Ctx({**cfg, "gain_adc_per_pe": gains}, {"records": records, "basic_features": bf}), "run")
    df_to_npz(G, "df_pe", df_pe)
    # df, st_waveforms mode with a source that has neither board nor record_id columns
    st = np.zeros(len(records), dtype=[("timestamp", "i8"), ("channel", "i2")])
    st["timestamp"], st["channel"] = records["timestamp"], records["channel"]
    df_st = DataFramePlugin().compute(Ctx({}, {"st_waveforms": st, "basic_features": bf}), "run")
    df_to_npz(G, "df_st", df_st)
    G["df_in_st"] = st

    # df_events -> df_paired: group with a wide window (multi-member events), pair with narrower ones
    spans = np.diff(np.sort(records["timestamp"])) / 1e3
    w_group = float(np.quantile(spans, 0.6))
    ev = GroupedEventsPlugin().compute(Ctx({"time_window_ns": w_group, "use_numba": False}, {"df": df}), "run")
    G["pair_group_window_ns"] = np.float64(w_group)
    G["pair_ev_offsets"] = np.concatenate([[0], np.cumsum([len(x) for x in ev["timestamps"]])]).astype(np.int64)
    G["pair_ev_timestamps"] = np.concatenate([np.asarray(x, np.int64) for x in ev["timestamps"]])
    G["pair_ev_areas"] = np.concatenate([np.asarray(x, np.float32) for x in ev["areas"]])
    G["pair_ev_heights"] = np.concatenate([np.asarray(x, np.float32) for x in ev["heights"]])
    G["pair_ev_channels"] = np.concatenate([np.asarray(x, np.int64) for x in ev["channels"]])
    G["pair_ev_dt_ns"] = ev["dt/ns"].to_numpy(np.float64)
    G["pair_ev_n_hits"] = ev["n_hits"].to_numpy(np.int64)
    for name, tw, nch, start in (("a", w_group, 2, 6), ("b", w_group * 0.35, 3, 0), ("c", 0.0, 1, 2)):
        paired = EventAnalyzer(n_channels=nch, start_channel_slice=start).pair_events(ev, time_window_ns=tw)
        G[f"pair_{name}_tw"] = np.float64(tw)
        G[f"pair_{name}_nch_start"] = np.asarray([nch, start], dtype=np.int64)
        G[f"pair_{name}_index"] = paired.index.to_numpy(np.int64)
        for c in paired.columns:
            if paired[c].dtype != object:
                G[f"pair_{name}_{c.replace('/', '_')}"] = paired[c].to_numpy()
                G[f"pair_{name}_{c.replace('/', '_')}_dtype"] = np.asarray(str(paired[c].dtype))
        print("pair", name, "kept", len(paired), "of", len(ev), {c: str(paired[c].dtype) for c in paired.columns if "ch" in c and c != "channels"})

    # s1_s2 on the waveform_width golden rows; the feature table is shorter than the largest record_id
    ww = g["ww_default"]
    feats = g["bf_default"][:150]
    G["s1s2_in_features"] = feats
    assert (ww["record_id"] >= 150).any()
    for name, conf in S1S2_CASES.items():
        out = S1S2ClassifierPlugin().compute(Ctx(dict(conf), {"waveform_width": ww, "basic_features": feats}), "run")
        G[f"s1s2_{name}"] = out
        print("s1_s2", name, np.bincount(out["label"], minlength=3))
    path = os.path.join(HERE, "after_golden.npz")
    np.savez_compressed(path, **G)
    print("wrote", path, len(G), "arrays")


if __name__ == "__main__":
    main()
