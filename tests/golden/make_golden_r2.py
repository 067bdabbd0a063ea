#!/usr/bin/env python
"""Golden vectors for the inputs and shapes added in round 2, from the LIVE reference:

  * `hit` on uint16 structured rows (numpy's uint16 diff / negation wrap), `hit` and `signal_peaks_stream` on rows
    without a baseline field (level = np.mean of the row), `waveform_width_integral` on genuinely negative int16
    rows, `hit_threshold` on structured rows whose event_length is shorter than the row;
  * BASELINE config 5 shapes - records of 256, 2048 and 8192 samples: basic_features, hit_threshold, `hit`,
    wave_pool_filtered (SG);
  * BASELINE config 3 - a 64-channel V1725-like run (dt = 4 ns) with positive pulses, chained
    wave_pool_filtered (SG) -> `hit` on the filtered rows -> waveform_width.

    python tests/golden/make_golden_r2.py        # rewrites tests/golden/r2_golden.npz

Build container only; the fixtures travel with the repo."""

from __future__ import annotations

import os
import sys
from unittest.mock import patch

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

from make_golden import Ctx, import_reference, st_from_records  # noqa: E402
from waveformanalysis_b200.synth import make_raw_run, records_from_raw  # noqa: E402


def strip_fields(arr, drop):
    keep = [n for n in arr.dtype.names if n not in drop]
    dt = np.dtype([(n, arr.dtype.fields[n][0]) for n in keep])
    out = np.zeros(len(arr), dtype=dt)
    for n in keep:
        out[n] = arr[n]
    return out


def main():
    import_reference()
    from waveform_analysis.core.plugins.builtin.cpu.basic_features import BasicFeaturesPlugin
    from waveform_analysis.core.plugins.builtin.cpu.filtering import create_filtered_waveform_dtype
    from waveform_analysis.core.plugins.builtin.cpu.hit_finder import ThresholdHitPlugin
    from waveform_analysis.core.plugins.builtin.cpu.peak_finding import HitFinderPlugin
    from waveform_analysis.core.plugins.builtin.cpu.records import WavePoolFilteredPlugin
    from waveform_analysis.core.plugins.builtin.cpu.waveform_width import WaveformWidthPlugin
    from waveform_analysis.core.plugins.builtin.cpu.waveform_width_integral import WaveformWidthIntegralPlugin
    from waveform_analysis.core.plugins.builtin.streaming.cpu.signal_peaks import SignalPeaksStreamPlugin
    from waveform_analysis.core.processing import records_builder as rb
    from waveform_analysis.core.processing.dtypes import create_record_dtype

    def run(plugin, data, cfg):
        return plugin.compute(Ctx(cfg, data), "run")

    def filtered_rows(st, fw):
        stf = np.zeros(len(st), dtype=create_filtered_waveform_dtype(st.dtype))
        for f in st.dtype.names:
            if f != "wave":
                stf[f] = st[f]
        stf["wave"] = fw.reshape(len(st), -1)
        return stf

    G = {}
    # ------------------------------------------------------------------ uint16 rows / rows without a baseline field
    raw = make_raw_run(3, 40, 400, seed=2024)
    rec, pool = records_from_raw(raw)
    st = st_from_records(rec, pool, create_record_dtype)
    G["a_records"], G["a_pool"] = rec, pool
    dt_u16 = np.dtype([(n, (np.uint16, (400,)) if n == "wave" else st.dtype.fields[n][0]) for n in st.dtype.names])
    st_u16 = np.zeros(len(st), dtype=dt_u16)
    for n in st.dtype.names:
        st_u16[n] = st[n] if n != "wave" else st["wave"].view(np.uint16)
    G["u16_deriv"] = run(HitFinderPlugin(), {"st_waveforms": st_u16}, {"use_filtered": False, "height": 12.0, "width": 2})
    G["u16_level"] = run(HitFinderPlugin(), {"st_waveforms": st_u16}, {"use_filtered": False, "use_derivative": False, "height": 20.0,
                                                                      "prominence": 4.0, "width": 3})
    G["u16_diffheight"] = run(HitFinderPlugin(), {"st_waveforms": st_u16}, {"use_filtered": False, "height": 60000.0, "width": 1,
                                                                           "prominence": 1.0, "height_method": "diff"})
    st_nb = strip_fields(st, {"baseline", "baseline_upstream"})
    G["nobase_i16_level"] = run(HitFinderPlugin(), {"st_waveforms": st_nb}, {"use_filtered": False, "use_derivative": False, "height": 20.0,
                                                                            "prominence": 4.0, "width": 3})
    G["nobase_i16_deriv"] = run(HitFinderPlugin(), {"st_waveforms": st_nb}, {"use_filtered": False, "height": 12.0, "width": 2})
    fw = run(WavePoolFilteredPlugin(), {"records": rec, "wave_pool": pool}, {"max_workers": 1})
    stf = filtered_rows(st, fw)
    G["a_filtered_pool"] = fw
    stf_nb = strip_fields(stf, {"baseline", "baseline_upstream"})
    G["nobase_f32_level"] = run(HitFinderPlugin(), {"filtered_waveforms": stf_nb, "st_waveforms": st_nb},
                                {"use_derivative": False, "height": 20.0, "prominence": 4.0, "width": 3})

    def stream(cfg, stw, stfw):
        plugin = SignalPeaksStreamPlugin()
        plugin.parallel = False
        chunks = list(plugin.compute(Ctx(cfg, {"filtered_waveforms": stfw, "st_waveforms": stw}), "run"))
        return np.concatenate([c.data for c in chunks])

    G["stream_nobase_level"] = stream({"use_derivative": False, "height": 25.0, "prominence": 4.0, "width": 3}, st_nb, stf_nb)
    # negative int16 samples
    st_neg = st.copy()
    st_neg["wave"] = st_neg["wave"] - 9000
    st_neg["baseline"] = st_neg["baseline"] - 9000
    st_neg["polarity"][::3] = "positive"
    G["neg_polarity"] = st_neg["polarity"]
    G["wint_neg"] = run(WaveformWidthIntegralPlugin(), {"st_waveforms": st_neg}, {})
    # event_length shorter than the row: the whole row is scanned, edges are clamped (hit_finder.py:183-230, 388-391)
    st_short = st.copy()
    rng = np.random.default_rng(5)
    st_short["event_length"] = rng.integers(150, 401, size=len(st))
    G["short_event_length"] = st_short["event_length"]
    bundle = rb.build_records_from_st_waveforms(st_short, default_dt_ns=2)
    with patch("waveform_analysis.core.plugins.builtin.cpu.records.get_records_bundle", return_value=bundle):
        G["short_hits"] = run(ThresholdHitPlugin(), {"st_waveforms": st_short}, {"threshold": 15.0})
        G["short_hits_ext"] = run(ThresholdHitPlugin(), {"st_waveforms": st_short}, {"threshold": 12.0, "left_extension": 5, "right_extension": 0})
    # ------------------------------------------------------------------ config 5 shapes
    for L, n_ch, n_per, seed in ((256, 8, 40, 61), (2048, 4, 12, 62), (8192, 2, 6, 63)):
        raw = make_raw_run(n_ch, n_per, L, seed=seed)
        r, p = records_from_raw(raw)
        G[f"L{L}_records"], G[f"L{L}_pool"] = r, p
        G[f"L{L}_bf"] = run(BasicFeaturesPlugin(), {"records": r, "wave_pool": p}, {"wave_source": "records"})
        G[f"L{L}_hits"] = run(ThresholdHitPlugin(), {"records": r, "wave_pool": p}, {"wave_source": "records", "threshold": 15.0})
        G[f"L{L}_sg"] = run(WavePoolFilteredPlugin(), {"records": r, "wave_pool": p}, {"max_workers": 1})
        G[f"L{L}_hit"] = run(HitFinderPlugin(), {"records": r, "wave_pool": p}, {"use_filtered": False, "wave_source": "records", "height": 12.0,
                                                                                  "width": 2})
        G[f"L{L}_hit_filt"] = run(HitFinderPlugin(), {"records": r, "wave_pool_filtered": G[f"L{L}_sg"]},
                                  {"use_filtered": True, "wave_source": "records", "height": 8.0, "width": 2})
    # ------------------------------------------------------------------ config 3: 64 channels, V1725-like, positive pulses
    raw = make_raw_run(64, 6, 800, seed=303, dt_ns=4, positive_pulses=True, n_boards=4)
    r, p = records_from_raw(raw, polarity="positive")
    stc = st_from_records(r, p, create_record_dtype)
    G["c3_records"], G["c3_pool"] = r, p
    G["c3_sg"] = run(WavePoolFilteredPlugin(), {"records": r, "wave_pool": p}, {"max_workers": 1})
    stfc = filtered_rows(stc, G["c3_sg"])
    # `hit` looks for negative pulses on structured rows; positive pulses are found on their falling edge
    G["c3_hit"] = run(HitFinderPlugin(), {"filtered_waveforms": stfc, "st_waveforms": stc}, {"height": 6.0, "width": 2})
    G["c3_hit_records"] = run(HitFinderPlugin(), {"records": r, "wave_pool_filtered": G["c3_sg"]},
                              {"use_filtered": True, "wave_source": "records", "height": 8.0, "width": 2})
    G["c3_width"] = run(WaveformWidthPlugin(), {"hit": G["c3_hit_records"], "filtered_waveforms": stfc, "st_waveforms": stc},
                        {"use_filtered": True, "sampling_rate": 0.25})
    G["c3_width_1050"] = run(WaveformWidthPlugin(), {"hit": G["c3_hit_records"], "filtered_waveforms": stfc, "st_waveforms": stc},
                             {"use_filtered": True, "sampling_rate": 0.25, "rise_low": 0.1, "rise_high": 0.5, "fall_high": 0.5, "fall_low": 0.1})
    G["c3_bf"] = run(BasicFeaturesPlugin(), {"records": r, "wave_pool": p}, {"wave_source": "records"})
    G["c3_hits_thr"] = run(ThresholdHitPlugin(), {"records": r, "wave_pool": p}, {"wave_source": "records", "threshold": 15.0})
    # ------------------------------------------------------------------ st_waveforms: dual baseline (waveforms.py:644-799)
    from waveform_analysis.core.plugins.builtin.cpu.waveforms import WaveformStruct, WaveformStructConfig

    raw = make_raw_run(3, 50, 400, seed=808)
    cols = WaveformStructConfig.default_vx2730().format_spec.columns
    arrays, upstream = [], []
    rng = np.random.default_rng(9)
    for c in range(3):
        sel = raw["channels"] == c
        arr = np.zeros((int(sel.sum()), 7 + 400), dtype=np.int64)
        arr[:, cols.board] = c % 2
        arr[:, cols.channel] = c
        arr[:, cols.timestamp] = raw["timestamps_ps"][sel] // 1000  # the VX2730 adapter scales its time stamps to ps
        arr[:, 7:] = raw["samples"][sel]
        arrays.append(arr)
        upstream.append(rng.normal(8000.0, 2.0, size=len(arr)))
    upstream[1] = upstream[1][:-3]  # wrong length: the reference falls back to NaN for that channel (:762-771)
    for k, arr in enumerate(arrays):
        G[f"st_raw{k}"] = arr
        G[f"st_up{k}"] = upstream[k]
    cfg = WaveformStructConfig.default_vx2730()
    G["st_default"] = WaveformStruct(arrays, config=cfg).structure_waveforms()
    G["st_upstream"] = WaveformStruct(arrays, config=cfg, upstream_baselines=upstream).structure_waveforms()
    G["st_window"] = WaveformStruct(arrays, config=cfg, baseline_samples=(10, 90)).structure_waveforms()
    cfg2 = WaveformStructConfig.default_vx2730()
    cfg2.wave_length = 256
    G["st_trunc"] = WaveformStruct(arrays, config=cfg2).structure_waveforms()
    bundle = rb.build_records_from_st_waveforms(G["st_upstream"], default_dt_ns=2)
    G["st_records"], G["st_pool"] = bundle.records, bundle.wave_pool
    out = os.path.join(HERE, "r2_golden.npz")
    np.savez_compressed(out, **G)
    print("wrote", out, {k: len(v) for k, v in G.items() if v.dtype.names}, os.path.getsize(out) / 1e3, "kB")


if __name__ == "__main__":
    main()
