#!/usr/bin/env python
"""Generate golden input/output vectors by running the LIVE reference (read-only at
/root/reference) on seeded synthetic inputs.  Run in the build container only:

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

The fixtures travel with the repo; the GPU box never reads /root/reference.
Every output array below is produced by the reference's own plugin ``compute`` (or the
reference function named in the key), not by anything in this repository.
"""

from __future__ import annotations

import os
import sys
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("WFB_REFERENCE_ROOT", "/root/reference")


def import_reference():
    for mod in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.colors",
                "matplotlib.figure", "matplotlib.axes", "matplotlib.gridspec", "matplotlib.lines",
                "matplotlib.collections", "matplotlib.cm", "matplotlib.ticker", "matplotlib.dates"):
        sys.modules.setdefault(mod, MagicMock())
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import waveform_analysis  # noqa: F401

    return waveform_analysis


class Ctx:
    """Minimal dict-backed context with the reference's config resolution order
    (same contract as the reference's tests/utils.py DummyContext)."""

    def __init__(self, config=None, data=None):
        self.config = config or {}
        self._results = {}
        for k, v in (data or {}).items():
            self._results[("run", k)] = v

    def get_config(self, plugin, name):
        p = plugin.provides
        if p in self.config and isinstance(self.config[p], dict) and name in self.config[p]:
            return self.config[p][name]
        if f"{p}.{name}" in self.config:
            return self.config[f"{p}.{name}"]
        if name in self.config:
            return self.config[name]
        if name in getattr(plugin, "options", {}):
            return plugin.options[name].default
        return None

    def get_data(self, run_id, name):
        return self._results.get((run_id, name))

    def _set_data(self, run_id, name, data):
        self._results[(run_id, name)] = data

    def key_for(self, run_id, name):
        return f"{run_id}-{name}"


def st_from_records(records, pool, create_record_dtype):
    L = int(records["event_length"][0])
    st = np.zeros(len(records), dtype=create_record_dtype(L))
    for f in ("baseline", "baseline_upstream", "polarity", "timestamp", "record_id", "dt", "event_length", "board", "channel"):
        st[f] = records[f]
    st["wave"] = pool.reshape(len(records), L).view(np.int16)
    return st


def main(out_dir=HERE):
    sys.path.insert(0, ROOT)
    import_reference()
    from waveform_analysis.core.plugins.builtin.cpu.basic_features import BasicFeaturesPlugin
    from waveform_analysis.core.plugins.builtin.cpu.hit_finder import ThresholdHitPlugin
    from waveform_analysis.core.plugins.builtin.cpu.hit_merge import (
        HitMergeClustersPlugin, HitMergedComponentsPlugin, HitMergePlugin)
    from waveform_analysis.core.plugins.builtin.cpu.peak_finding import HitFinderPlugin
    from waveform_analysis.core.plugins.builtin.cpu.records import WavePoolFilteredPlugin
    from waveform_analysis.core.plugins.builtin.cpu.waveform_width import WaveformWidthPlugin
    from waveform_analysis.core.plugins.builtin.cpu.waveform_width_integral import (
        WaveformWidthIntegralPlugin)
    from waveform_analysis.core.plugins.builtin.cpu.event_analysis import HitGroupedPlugin
    from waveform_analysis.core.processing import records_builder as rb
    from waveform_analysis.core.processing.dtypes import create_record_dtype
    from waveform_analysis.core.processing.event_grouping import group_multi_channel_hits
    from waveform_analysis.utils.formats import get_adapter
    import pandas as pd

    from waveformanalysis_b200.synth import make_raw_run, make_ragged_records, records_from_raw

    G = {}
    # ---------------------------------------------------------------- records builder (K1)
    raw = make_raw_run(4, 150, 800, seed=1234)
    adapter = get_adapter("vx2730")
    cols = adapter.format_spec.columns
    parts = []
    for c in range(4):
        sel = raw["channels"] == c
        n_c = int(sel.sum())
        arr = np.zeros((n_c, 7 + 800), dtype=np.int64)
        arr[:, cols.board] = raw["boards"][sel]
        arr[:, cols.channel] = raw["channels"][sel]
        arr[:, cols.timestamp] = raw["timestamps_ps"][sel]
        arr[:, 7:] = raw["samples"][sel]
        parts.append(rb._build_records_part_from_raw_array(
            arr, channel_idx=c, default_dt_ns=2, cols=cols,
            normalize_timestamp_to_ps=adapter.format_spec.normalize_timestamp_to_ps,
            baseline_samples=None))
    bundle = rb.merge_records_parts(parts)
    # the streaming builder renumbers record_id after the merge (records_builder.py:425)
    bundle.records["record_id"] = np.arange(len(bundle.records))
    G["raw_timestamps_ps"] = raw["timestamps_ps"]
    G["raw_boards"] = raw["boards"]
    G["raw_channels"] = raw["channels"]
    G["raw_samples"] = raw["samples"]
    G["records"] = bundle.records
    G["wave_pool"] = bundle.wave_pool
    records, pool = bundle.records.copy(), bundle.wave_pool.copy()

    def run(plugin, data, config=None):
        return plugin.compute(Ctx(config, data), "run")

    # ---------------------------------------------------------------- basic_features
    base = {"records": records, "wave_pool": pool}
    G["bf_default"] = run(BasicFeaturesPlugin(), base, {"wave_source": "records"})
    G["bf_fullrange"] = run(BasicFeaturesPlugin(), base, {"wave_source": "records", "height_range": (0, None), "area_range": (100, -50)})
    for pol in ("negative", "positive"):
        r2 = records.copy()
        r2["polarity"] = pol
        G[f"bf_{pol}"] = run(BasicFeaturesPlugin(), {"records": r2, "wave_pool": pool}, {"wave_source": "records", "height_range": (0, None)})
    G["bf_fixed"] = run(BasicFeaturesPlugin(), base, {"wave_source": "records", "channel_config": {"channels": {"0:1": {"fixed_baseline": 8000.5}, "0:3": {"fixed_baseline": 7990.0}}}})
    # ---------------------------------------------------------------- hit_threshold
    G["hits_thr15"] = run(ThresholdHitPlugin(), base, {"wave_source": "records", "threshold": 15.0})
    G["hits_chan"] = run(ThresholdHitPlugin(), base, {"wave_source": "records", "threshold": 12.0, "left_extension": 5, "right_extension": 0,
                                                     "channel_config": {"channels": {"0:2": {"threshold": 40.0}}}})
    rpos = records.copy()
    rpos["polarity"] = "positive"
    G["hits_positive"] = run(ThresholdHitPlugin(), {"records": rpos, "wave_pool": pool}, {"wave_source": "records", "threshold": -20.0})
    # ---------------------------------------------------------------- ragged records
    rr, rp = make_ragged_records(400, seed=7)
    G["rag_records"] = rr
    G["rag_pool"] = rp
    ragged = {"records": rr, "wave_pool": rp}
    G["rag_bf"] = run(BasicFeaturesPlugin(), ragged, {"wave_source": "records", "height_range": (5, -5), "area_range": (0, None)})
    G["rag_hits"] = run(ThresholdHitPlugin(), ragged, {"wave_source": "records", "threshold": 15.0, "left_extension": 3, "right_extension": 4})
    G["rag_wint"] = run(WaveformWidthIntegralPlugin(), ragged, {"wave_source": "records"})
    # ---------------------------------------------------------------- wave_pool_filtered
    sub = slice(0, 120)
    rs, ps = records[sub].copy(), pool[: 120 * 800].copy()
    fbase = {"records": rs, "wave_pool": ps}
    G["filt_records"] = rs
    G["filt_pool"] = ps
    G["filt_sg"] = run(WavePoolFilteredPlugin(), fbase, {"max_workers": 1})
    G["filt_sg_21_3"] = run(WavePoolFilteredPlugin(), fbase, {"max_workers": 1, "sg_window_size": 21, "sg_poly_order": 3})
    bwcfg = {"max_workers": 1, "filter_type": "BW", "lowcut": 0.01, "highcut": 0.1, "fs": 0.5, "filter_order": 4}
    G["filt_bw"] = run(WavePoolFilteredPlugin(), fbase, bwcfg)
    G["filt_mixed"] = run(WavePoolFilteredPlugin(), fbase, {"max_workers": 1, "channel_config": {"channels": {"0:1": {"filter_type": "BW", "lowcut": 0.02, "highcut": 0.2, "fs": 1.0, "filter_order": 2}}}})
    fsg = {"records": rs, "wave_pool": ps, "wave_pool_filtered": G["filt_sg"]}
    G["filt_bf"] = run(BasicFeaturesPlugin(), fsg, {"wave_source": "records", "use_filtered": True})
    G["filt_hits"] = run(ThresholdHitPlugin(), fsg, {"wave_source": "records", "use_filtered": True, "threshold": 15.0})
    rs_neg = rs.copy()
    rs_neg["polarity"] = "negative"
    G["filt_bf_negative"] = run(BasicFeaturesPlugin(), {"records": rs_neg, "wave_pool": ps, "wave_pool_filtered": G["filt_sg"]},
                                {"wave_source": "records", "use_filtered": True, "height_range": (0, None)})
    # ---------------------------------------------------------------- width integral
    G["wint_default"] = run(WaveformWidthIntegralPlugin(), {"records": records[:200], "wave_pool": pool}, {"wave_source": "records"})
    rneg = records[:200].copy()
    rneg["polarity"] = "negative"
    G["wint_negative"] = run(WaveformWidthIntegralPlugin(), {"records": rneg, "wave_pool": pool}, {"wave_source": "records", "q_low": 0.2, "q_high": 0.8, "dt": 2.0})
    # ---------------------------------------------------------------- waveform_width (positive pulses)
    rawp = make_raw_run(3, 60, 800, seed=99, positive_pulses=True)
    rp_rec, rp_pool = records_from_raw(rawp, polarity="positive")
    st = st_from_records(rp_rec, rp_pool, create_record_dtype)
    hits = run(HitFinderPlugin(), {"records": rp_rec, "wave_pool": rp_pool}, {"use_filtered": False, "wave_source": "records", "use_derivative": False,
                                                          "height": 30.0, "prominence": 5.0, "width": 2, "distance": 5})
    G["ww_records"] = rp_rec
    G["ww_pool"] = rp_pool
    G["ww_hit"] = hits
    G["ww_default"] = run(WaveformWidthPlugin(), {"hit": hits, "st_waveforms": st}, {})
    G["ww_50"] = run(WaveformWidthPlugin(), {"hit": hits, "st_waveforms": st}, {"rise_low": 0.1, "rise_high": 0.5, "fall_high": 0.5, "fall_low": 0.1, "sampling_rate": 0.25})
    G["ww_nointerp"] = run(WaveformWidthPlugin(), {"hit": hits, "st_waveforms": st}, {"interpolation": False})
    # filtered (float32) waves
    fw = run(WavePoolFilteredPlugin(), {"records": rp_rec, "wave_pool": rp_pool}, {"max_workers": 1})
    from waveform_analysis.core.plugins.builtin.cpu.filtering import create_filtered_waveform_dtype
    stf = np.zeros(len(st), dtype=create_filtered_waveform_dtype(st.dtype))
    for f in st.dtype.names:
        if f != "wave":
            stf[f] = st[f]
    stf["wave"] = fw.reshape(len(st), 800)
    G["ww_filtered_pool"] = fw
    G["ww_filtered"] = run(WaveformWidthPlugin(), {"hit": hits, "filtered_waveforms": stf}, {"use_filtered": True})
    # ---------------------------------------------------------------- merge + grouping
    h = G["hits_thr15"]
    for tag, cfg in (("m0", {}), ("m50", {"merge_gap_ns": 50.0, "max_total_width_ns": 400.0})):
        cl = run(HitMergeClustersPlugin(), {"hit_threshold": h}, cfg)
        mg = run(HitMergePlugin(), {"hit_threshold": h, "hit_merge_clusters": cl}, cfg)
        cp = run(HitMergedComponentsPlugin(), {"hit_threshold": h, "hit_merge_clusters": cl, "hit_merged": mg}, cfg)
        G[f"{tag}_clusters"], G[f"{tag}_merged"], G[f"{tag}_components"] = cl, mg, cp
        if tag == "m0":
            for wname, w in (("w100", 100.0), ("w0", 0.0), ("w2000", 2000.0)):
                df = run(HitGroupedPlugin(), {"hit_merged": mg, "hit_merged_components": cp, "hit_threshold": h}, {"time_window_ns": w})
                G[f"hg_{wname}_t_min"] = df["t_min"].to_numpy(np.int64)
                G[f"hg_{wname}_t_max"] = df["t_max"].to_numpy(np.int64)
                G[f"hg_{wname}_dt_ns"] = df["dt/ns"].to_numpy(np.float64)
                G[f"hg_{wname}_n_hits"] = df["n_hits"].to_numpy(np.int64)
                G[f"hg_{wname}_record_ids"] = np.concatenate([np.asarray(v, np.int64) for v in df["record_ids"]])
                G[f"hg_{wname}_timestamps"] = np.concatenate([np.asarray(v, np.int64) for v in df["timestamps"]])
                G[f"hg_{wname}_channels"] = np.concatenate([np.asarray(v, np.int64) for v in df["channels"]])
    bf = G["bf_default"]
    dfin = pd.DataFrame({"timestamp": bf["timestamp"], "channel": bf["channel"], "area": bf["area"], "height": bf["height"]})
    for wname, w in (("w100", 100.0), ("w30000", 30000.0)):
        for numba in (True, False):
            ev = group_multi_channel_hits(dfin, w, use_numba=numba)
            tag = f"ge_{wname}_{'nb' if numba else 'np'}"
            G[f"{tag}_t_min"] = ev["t_min"].to_numpy(np.int64)
            G[f"{tag}_t_max"] = ev["t_max"].to_numpy(np.int64)
            G[f"{tag}_n_hits"] = ev["n_hits"].to_numpy(np.int64)
            G[f"{tag}_timestamps"] = np.concatenate([np.asarray(v, np.int64) for v in ev["timestamps"]])
            G[f"{tag}_channels"] = np.concatenate([np.asarray(v, np.int64) for v in ev["channels"]])

    # ---------------------------------------------------------------- st_waveforms / filtered_waveforms sources
    from unittest.mock import patch

    st_small = st_from_records(records[:300], pool[: 300 * 800], create_record_dtype)
    G["st_bf_default"] = run(BasicFeaturesPlugin(), {"st_waveforms": st_small}, {})
    st_pos = st_small.copy()
    st_pos["polarity"][::2] = "positive"
    st_pos["polarity"][1::4] = "negative"
    G["st_polarity"] = st_pos["polarity"]
    G["st_bf_polarity"] = run(BasicFeaturesPlugin(), {"st_waveforms": st_pos}, {"height_range": (0, None), "area_range": (10, 700),
                                                                             "channel_config": {"channels": {"0:2": {"fixed_baseline": 8011.25}}}})
    st_neg = st_pos.copy()
    st_neg["wave"] = st_neg["wave"] - 9000  # genuinely negative int16 samples
    st_neg["baseline"] = st_neg["baseline"] - 9000
    G["st_neg_bf"] = run(BasicFeaturesPlugin(), {"st_waveforms": st_neg}, {"height_range": (0, None)})

    def st_hits(stw, cfg):
        bundle = rb.build_records_from_st_waveforms(stw, default_dt_ns=2)
        with patch("waveform_analysis.core.plugins.builtin.cpu.records.get_records_bundle", return_value=bundle):
            return run(ThresholdHitPlugin(), {"st_waveforms": stw}, cfg)

    G["st_hits"] = st_hits(st_pos, {"threshold": 15.0})
    G["st_neg_hits"] = st_hits(st_neg, {"threshold": 15.0, "left_extension": 4, "right_extension": 1})
    G["st_wint"] = run(WaveformWidthIntegralPlugin(), {"st_waveforms": st_pos}, {})
    stf_small = np.zeros(len(st_small), dtype=create_filtered_waveform_dtype(st_small.dtype))
    for f in st_small.dtype.names:
        if f != "wave":
            stf_small[f] = st_pos[f]
    stf_small["wave"] = run(WavePoolFilteredPlugin(), {"records": records[:300], "wave_pool": pool[: 300 * 800]}, {"max_workers": 1}).reshape(300, 800)
    G["stf_wave"] = stf_small["wave"]
    G["stf_bf"] = run(BasicFeaturesPlugin(), {"filtered_waveforms": stf_small}, {"use_filtered": True})

    out = os.path.join(out_dir, "hotpath_golden.npz")
    np.savez_compressed(out, **G)
    print("wrote", out, "keys:", len(G), "size MB:", os.path.getsize(out) / 1e6)
    return G


if __name__ == "__main__":
    main()
