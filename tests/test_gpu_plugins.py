"""GPU: the drop-in plugins' compute() (the reference-facing boundary) against the golden vectors
produced by the reference's own plugins with the same configs (tests/golden/make_golden.py)."""

import numpy as np
import pytest

from conftest import assert_rows_match
from fakes import Ctx

pytestmark = pytest.mark.gpu

FX_BF = ("height", "amp", "max_abs_diff")
FX_HIT = ("height", "width", "rise_time", "fall_time")


@pytest.fixture(scope="module")
def P():
    from waveformanalysis_b200 import plugins

    return plugins


def run(plugin, data, config=None):
    return plugin.compute(Ctx(config, data), "run")


def st_from_records(records, pool, L=800):
    from waveformanalysis_b200.dtypes import create_record_dtype

    st = np.zeros(len(records), dtype=create_record_dtype(L))
    for f in ("baseline", "baseline_upstream", "polarity", "timestamp", "record_id", "dt", "event_length", "board", "channel"):
        st[f] = records[f]
    st["wave"] = pool.reshape(len(records), L).view(np.int16)
    return st


def test_basic_features_plugin_records_source(P, golden):
    base = {"records": golden["records"], "wave_pool": golden["wave_pool"]}
    assert_rows_match(run(P.B200BasicFeaturesPlugin(), base, {"wave_source": "records"}), golden["bf_default"], what="bf", float_exact=FX_BF)
    out = run(P.B200BasicFeaturesPlugin(), base, {"wave_source": "records", "channel_config": {"channels": {"0:1": {"fixed_baseline": 8000.5}, "0:3": {"fixed_baseline": 7990.0}}}})
    assert_rows_match(out, golden["bf_fixed"], what="bf_fixed", float_exact=FX_BF)
    fsg = {"records": golden["filt_records"], "wave_pool": golden["filt_pool"], "wave_pool_filtered": golden["filt_sg"]}
    assert_rows_match(run(P.B200BasicFeaturesPlugin(), fsg, {"wave_source": "records", "use_filtered": True}), golden["filt_bf"], what="filt_bf")


def test_threshold_hit_plugin_records_source(P, golden):
    base = {"records": golden["records"], "wave_pool": golden["wave_pool"]}
    assert_rows_match(run(P.B200ThresholdHitPlugin(), base, {"wave_source": "records", "threshold": 15.0}), golden["hits_thr15"], what="thr15", float_exact=FX_HIT)
    cfg = {"wave_source": "records", "threshold": 12.0, "left_extension": 5, "right_extension": 0, "channel_config": {"channels": {"0:2": {"threshold": 40.0}}}}
    assert_rows_match(run(P.B200ThresholdHitPlugin(), base, cfg), golden["hits_chan"], what="chan", float_exact=FX_HIT)
    with pytest.raises(ValueError, match="Invalid channel key"):
        run(P.B200ThresholdHitPlugin(), base, {"wave_source": "records", "channel_config": {"run": {"1": {"threshold": 5.0}}}})
    assert len(run(P.B200ThresholdHitPlugin(), {"records": golden["records"][:0], "wave_pool": golden["wave_pool"][:0]}, {"wave_source": "records"})) == 0


def test_structured_waveform_sources(P, golden):
    """wave_source='auto' (the reference default): st_waveforms / filtered_waveforms rows are used in
    place as the sample pool, including genuinely negative int16 samples."""
    rec, pool = golden["records"][:300], golden["wave_pool"][: 300 * 800]
    st = st_from_records(rec, pool)
    assert_rows_match(run(P.B200BasicFeaturesPlugin(), {"st_waveforms": st}, {}), golden["st_bf_default"], what="st_bf", float_exact=FX_BF)
    st_pos = st.copy()
    st_pos["polarity"] = golden["st_polarity"]
    cfg = {"height_range": (0, None), "area_range": (10, 700), "channel_config": {"channels": {"0:2": {"fixed_baseline": 8011.25}}}}
    assert_rows_match(run(P.B200BasicFeaturesPlugin(), {"st_waveforms": st_pos}, cfg), golden["st_bf_polarity"], what="st_bf_pol", float_exact=FX_BF)
    st_neg = st_pos.copy()
    st_neg["wave"] = st_neg["wave"] - 9000
    st_neg["baseline"] = st_neg["baseline"] - 9000
    assert_rows_match(run(P.B200BasicFeaturesPlugin(), {"st_waveforms": st_neg}, {"height_range": (0, None)}), golden["st_neg_bf"], what="st_neg_bf", float_exact=FX_BF)
    assert_rows_match(run(P.B200ThresholdHitPlugin(), {"st_waveforms": st_pos}, {"threshold": 15.0}), golden["st_hits"], what="st_hits", float_exact=FX_HIT)
    assert_rows_match(run(P.B200ThresholdHitPlugin(), {"st_waveforms": st_neg}, {"threshold": 15.0, "left_extension": 4, "right_extension": 1}),
                      golden["st_neg_hits"], what="st_neg_hits", float_exact=FX_HIT)
    assert_rows_match(run(P.B200WaveformWidthIntegralPlugin(), {"st_waveforms": st_pos}, {}), golden["st_wint"], what="st_wint")
    from waveformanalysis_b200.dtypes import create_record_dtype

    stf = np.zeros(len(st), dtype=[(n, (np.float32, (800,)) if n == "wave" else create_record_dtype(800)[n]) for n in create_record_dtype(800).names])
    for f in st.dtype.names:
        if f != "wave":
            stf[f] = st_pos[f]
    stf["wave"] = golden["stf_wave"]
    assert_rows_match(run(P.B200BasicFeaturesPlugin(), {"filtered_waveforms": stf}, {"use_filtered": True}), golden["stf_bf"], what="stf_bf")


def test_wave_pool_filtered_plugin(P, golden, monkeypatch):
    fbase = {"records": golden["filt_records"], "wave_pool": golden["filt_pool"]}
    assert np.allclose(run(P.B200WavePoolFilteredPlugin(), fbase, {}), golden["filt_sg"], rtol=1e-5, atol=1e-3)
    bw = {"filter_type": "BW", "lowcut": 0.01, "highcut": 0.1, "fs": 0.5, "filter_order": 4}
    assert np.allclose(run(P.B200WavePoolFilteredPlugin(), fbase, bw), golden["filt_bw"], rtol=1e-5, atol=1e-3)
    monkeypatch.setenv("WFB_BW_EXACT", "1")  # scipy's order of operations without contraction: bit-exact
    assert np.array_equal(run(P.B200WavePoolFilteredPlugin(), fbase, bw), golden["filt_bw"])
    monkeypatch.delenv("WFB_BW_EXACT")
    mixed = {"channel_config": {"channels": {"0:1": {"filter_type": "BW", "lowcut": 0.02, "highcut": 0.2, "fs": 1.0, "filter_order": 2}}}}
    assert np.allclose(run(P.B200WavePoolFilteredPlugin(), fbase, mixed), golden["filt_mixed"], rtol=1e-5, atol=1e-3)
    with pytest.raises(ValueError):  # the reference's default BW options are invalid (highcut >= fs/2), filtering.py:98-99
        run(P.B200WavePoolFilteredPlugin(), fbase, {"filter_type": "BW"})
    with pytest.raises(ValueError):
        run(P.B200WavePoolFilteredPlugin(), fbase, {"filter_type": "XX"})


def test_width_plugins(P, golden):
    rec, pool, hits = golden["ww_records"], golden["ww_pool"], golden["ww_hit"]
    st = st_from_records(rec, pool)
    fx = ("rise_time", "fall_time", "total_width", "rise_time_samples", "fall_time_samples", "total_width_samples", "peak_height")
    assert_rows_match(run(P.B200WaveformWidthPlugin(), {"hit": hits, "st_waveforms": st}, {}), golden["ww_default"], what="ww", float_exact=fx)
    cfg = {"rise_low": 0.1, "rise_high": 0.5, "fall_high": 0.5, "fall_low": 0.1, "sampling_rate": 0.25}
    assert_rows_match(run(P.B200WaveformWidthPlugin(), {"hit": hits, "st_waveforms": st}, cfg), golden["ww_50"], what="ww50", float_exact=fx)
    base = {"records": golden["records"][:200], "wave_pool": golden["wave_pool"]}
    fxi = ("t_low", "t_high", "width", "t_low_samples", "t_high_samples", "width_samples", "q_total")
    assert_rows_match(run(P.B200WaveformWidthIntegralPlugin(), base, {"wave_source": "records"}), golden["wint_default"], what="wint", float_exact=fxi)


def test_merge_and_grouping_plugins(P, golden):
    h = golden["hits_thr15"]
    ctx = Ctx({}, {"hit_threshold": h}, plugins={"hit_merged": P.B200HitMergePlugin()})
    cl = P.B200HitMergeClustersPlugin().compute(ctx, "run")
    ctx._set_data("run", "hit_merge_clusters", cl)
    mg = P.B200HitMergePlugin().compute(ctx, "run")
    ctx._set_data("run", "hit_merged", mg)
    cp = P.B200HitMergedComponentsPlugin().compute(ctx, "run")
    ctx._set_data("run", "hit_merged_components", cp)
    assert_rows_match(cl, golden["m0_clusters"], what="clusters")
    assert_rows_match(mg, golden["m0_merged"], what="merged", float_exact=("height", "integral", "width", "rise_time", "fall_time"))
    assert_rows_match(cp, golden["m0_components"], what="components")
    # chain merging (merge_gap_ns > 0) with the max-total-width cut
    cfg = {"merge_gap_ns": 50.0, "max_total_width_ns": 400.0}
    c50 = Ctx(cfg, {"hit_threshold": h}, plugins={"hit_merged": P.B200HitMergePlugin()})
    cl50 = P.B200HitMergeClustersPlugin().compute(c50, "run")
    mg50 = P.B200HitMergePlugin().compute(c50, "run")
    assert_rows_match(cl50, golden["m50_clusters"], what="m50 clusters")
    assert_rows_match(mg50, golden["m50_merged"], what="m50 merged", float_exact=("height", "integral", "width", "rise_time", "fall_time"))
    assert len(mg50) < len(mg)
    # grouping of the chain-merged rows, some of which span two records (windows from the component hits)
    cp50 = P.B200HitMergedComponentsPlugin().compute(c50, "run")
    assert_rows_match(cp50, golden["m50_components"], what="m50 components")
    from test_oracle_golden import grouping50_cases

    for w, want in grouping50_cases(golden):
        gctx = Ctx({"time_window_ns": w}, {"hit_merged": mg50, "hit_merged_components": cp50, "hit_threshold": h})
        df = P.B200HitGroupedPlugin().compute(gctx, "run")
        assert np.array_equal(df["t_min"].to_numpy(), want["t_min"]) and np.array_equal(df["t_max"].to_numpy(), want["t_max"])
        assert np.array_equal(df["n_hits"].to_numpy(), want["n_hits"])
        assert np.array_equal(np.concatenate([np.asarray(v, np.int64) for v in df["record_ids"]]), want["record_ids"])
        assert np.array_equal(np.concatenate([np.asarray(v, np.int64) for v in df["sample_starts"]]), want["sample_starts"])
    for wname, w in (("w100", 100.0), ("w0", 0.0)):
        ctx.config = {"time_window_ns": w}
        df = P.B200HitGroupedPlugin().compute(ctx, "run")
        assert list(df.columns) == ["event_id", "t_min", "t_max", "dt/ns", "n_hits", "dt", "boards", "channels", "heights", "integrals",
                                    "timestamps", "record_ids", "sample_starts", "sample_ends"]
        assert np.array_equal(df["t_min"].to_numpy(), golden[f"hg_{wname}_t_min"])
        assert np.array_equal(df["t_max"].to_numpy(), golden[f"hg_{wname}_t_max"])
        assert np.array_equal(df["n_hits"].to_numpy(), golden[f"hg_{wname}_n_hits"])
        assert np.array_equal(df["dt/ns"].to_numpy(), golden[f"hg_{wname}_dt_ns"])
        assert np.array_equal(np.concatenate(list(df["record_ids"])), golden[f"hg_{wname}_record_ids"])
        assert np.array_equal(np.concatenate(list(df["channels"])), golden[f"hg_{wname}_channels"])
    empty = P.B200HitGroupedPlugin().compute(Ctx({}, {"hit_merged": mg[:0], "hit_merged_components": cp[:0], "hit_threshold": h[:0]}), "run")
    assert empty.empty and len(empty.columns) == 14


def test_df_events_plugin(P, golden):
    import pandas as pd

    bf = golden["bf_default"]
    df = pd.DataFrame({"timestamp": bf["timestamp"], "channel": bf["channel"], "area": bf["area"], "height": bf["height"]})
    ev = P.B200GroupedEventsPlugin().compute(Ctx({"time_window_ns": 100.0}, {"df": df}), "run")
    assert np.array_equal(ev["t_min"].to_numpy(), golden["ge_w100_nb_t_min"])
    assert np.array_equal(ev["t_max"].to_numpy(), golden["ge_w100_nb_t_max"])
    assert np.array_equal(ev["n_hits"].to_numpy(), golden["ge_w100_nb_n_hits"])
    assert np.array_equal(np.concatenate(list(ev["timestamps"])), golden["ge_w100_nb_timestamps"])


def test_records_plugins_from_raw_arrays(P, golden):
    """records / wave_pool from per-channel raw rows (VX2730 column layout: board, channel, timestamp
    at 0..2, samples from column 7), the arrays the reference's readers hand to the builder."""
    raws = []
    for c in range(4):
        sel = golden["raw_channels"] == c
        arr = np.zeros((int(sel.sum()), 7 + 800), dtype=np.int64)
        arr[:, 0], arr[:, 1], arr[:, 2] = golden["raw_boards"][sel], c, golden["raw_timestamps_ps"][sel]
        arr[:, 7:] = golden["raw_samples"][sel]
        raws.append(arr)
    ctx = Ctx({"dt": 2}, {"raw_arrays": raws})
    rec = P.B200RecordsPlugin().compute(ctx, "run")
    pool = P.B200WavePoolPlugin().compute(ctx, "run")
    want = golden["records"]
    for name in want.dtype.names:
        assert np.array_equal(rec[name], want[name], equal_nan=(want[name].dtype.kind == "f")), name
    assert np.array_equal(pool, golden["wave_pool"])


def test_records_plugins_from_v1725_files(P, tmp_path):
    """daq_adapter="v1725": the .bin files named in raw_files are indexed on the host and decoded on the
    device; rows must equal the reference's build_records_from_v1725_files output (golden)."""
    from test_oracle_golden import v1725_cases

    for tag, blobs, names, dt_ns, want_rec, want_pool in v1725_cases():
        paths = []
        for blob, name in zip(blobs, names):
            p = tmp_path / tag / name
            p.parent.mkdir(parents=True, exist_ok=True)
            p.write_bytes(blob)
            paths.append(str(p))
        # the reference groups files per channel / board and lists may repeat a path
        groups = [[paths[0]]] + [[p] for p in paths[1:]] + [[paths[0]]]
        ctx = Ctx({"daq_adapter": "v1725", "dt": dt_ns}, {"raw_files": groups})
        rec = P.B200RecordsPlugin().compute(ctx, "run")
        pool = P.B200WavePoolPlugin().compute(ctx, "run")
        assert np.array_equal(pool, want_pool), tag
        assert_rows_match(rec, want_rec, what=f"v1725 plugin {tag}", float_exact=("baseline",))


def test_hit_finder_plugin(P):
    """B200HitFinderPlugin against the reference HitFinderPlugin rows (hit_golden.npz) through the plugin
    interface: filtered_waveforms (default), st_waveforms and records sources."""
    import os

    from waveformanalysis_b200.dtypes import HIT_DTYPE, create_record_dtype

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "hit_golden.npz"), allow_pickle=False)
    rec, pool, fpool = g["records"], g["pool"], g["filtered_pool"]
    n, L = len(rec), 800
    st = st_from_records(rec, pool, L)
    stf_dtype = np.dtype([(name, (np.float32, (L,)) if name == "wave" else st.dtype.fields[name][0]) for name in st.dtype.names])
    stf = np.zeros(n, dtype=stf_dtype)
    for f in st.dtype.names:
        if f != "wave":
            stf[f] = st[f]
    stf["wave"] = fpool.reshape(n, L)
    fx = ("height", "edge_start", "edge_end")
    assert_rows_match(run(P.B200HitFinderPlugin(), {"filtered_waveforms": stf, "st_waveforms": st}, {}), g["filt_default"], what="hit default", float_exact=fx)
    assert_rows_match(run(P.B200HitFinderPlugin(), {"st_waveforms": st}, {"use_filtered": False, "height": 12.0, "width": 2}), g["st_default"],
                      what="hit st", float_exact=fx)
    assert_rows_match(run(P.B200HitFinderPlugin(), {"records": rec, "wave_pool": pool},
                          {"use_filtered": False, "wave_source": "records", "height": 12.0, "width": 2}), g["rec_default"], what="hit records", float_exact=fx)
    empty = run(P.B200HitFinderPlugin(), {"st_waveforms": st[:0]}, {"use_filtered": False})
    assert empty.dtype == HIT_DTYPE and len(empty) == 0
    with pytest.raises(ValueError):
        run(P.B200HitFinderPlugin(), {"st_waveforms": st}, {"use_filtered": False, "height_method": "area"})


def test_signal_peaks_stream_chunks(P):
    """B200SignalPeaksStreamPlugin.compute_chunk on per-channel chunks (what the reference's chunk iterator
    yields for this input) against the reference plugin's rows (hit_golden.npz stream_*)."""
    import os
    from types import SimpleNamespace

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "hit_golden.npz"), allow_pickle=False)
    rec, pool, fpool = g["records"], g["pool"], g["filtered_pool"]
    n, L = len(rec), 800
    st = st_from_records(rec, pool, L)
    stf_dtype = np.dtype([(name, (np.float32, (L,)) if name == "wave" else st.dtype.fields[name][0]) for name in st.dtype.names])
    stf = np.zeros(n, dtype=stf_dtype)
    for f in st.dtype.names:
        if f != "wave":
            stf[f] = st[f]
    stf["wave"] = fpool.reshape(n, L)
    for tag, cfg in (("stream_default", {"height": 10.0}),
                     ("stream_minmax", {"height": 10.0, "height_method": "minmax", "minmax_window_expand": 3, "width": 2}),
                     ("stream_level", {"use_derivative": False, "height": 25.0, "prominence": 4.0, "width": 3})):
        plugin = P.B200SignalPeaksStreamPlugin()
        ctx = Ctx(cfg, {})
        plugin._load_config(ctx)
        parts, bounds = [], []
        for ch in np.unique(st["channel"]):
            sel = st["channel"] == ch
            chunk = SimpleNamespace(data=st[sel], metadata={"filtered_waveforms": stf[sel], "event_offset": 0})
            out = plugin.compute_chunk(chunk, ctx, "run")
            if out is not None:
                parts.append(out.data)
                bounds.append([out.start, out.end])
        assert_rows_match(np.concatenate(parts), g[tag], what=tag, float_exact=("height", "edge_start", "edge_end"))
        # compute_chunk reports the span of its peaks; the framework widens it to the input chunk's main range
        # (checked on the CPU in test_plugin_contract.py::test_streaming_plugin_runs_inside_the_reference_framework)
        for (lo, hi), part, (mlo, mhi) in zip(bounds, parts, g[tag + "_bounds"]):
            assert lo == part["timestamp"].min() and hi == part["timestamp"].max() and mlo <= lo and hi <= mhi


# ---- the step after the path: df / df_events -> df_paired / s1_s2 ------------------------------------------


def test_df_plugin(P, golden):
    from after_cases import DF_COLUMNS, GAINS_CONFIG, check_df_columns, load_after

    A = load_after()
    rec, bf = A["df_in_records"], A["df_in_features"]

    def as_cols(df):
        cols = {c: df[c].to_numpy() for c in df.columns}
        cols["order"] = df.index.to_numpy()
        return cols

    df = run(P.B200DataFramePlugin(), {"records": rec, "basic_features": bf}, {"wave_source": "records"})
    assert tuple(df.columns) == DF_COLUMNS
    check_df_columns(as_cols(df), A, "df_plain", False)
    with pytest.warns(UserWarning, match="invalid 'gain_adc_per_pe'"):
        df_pe = run(P.B200DataFramePlugin(), {"records": rec, "basic_features": bf}, {"wave_source": "records", "gain_adc_per_pe": GAINS_CONFIG})
    assert tuple(df_pe.columns) == DF_COLUMNS + ("area_pe", "height_pe")
    check_df_columns(as_cols(df_pe), A, "df_pe", True)
    df_st = run(P.B200DataFramePlugin(), {"st_waveforms": A["df_in_st"], "basic_features": bf})
    check_df_columns(as_cols(df_st), A, "df_st", False)
    with pytest.raises(ValueError, match="basic_features length"):
        run(P.B200DataFramePlugin(), {"records": rec[:-1], "basic_features": bf}, {"wave_source": "records"})
    empty = run(P.B200DataFramePlugin(), {"records": rec[:0], "basic_features": bf[:0]}, {"wave_source": "records"})
    assert len(empty) == 0 and tuple(empty.columns) == DF_COLUMNS


def test_df_paired_chain(P):
    """df -> df_events -> df_paired through the three B200 plugins against the reference chain."""
    from after_cases import load_after

    A = load_after()
    rec, bf = A["df_in_records"], A["df_in_features"]
    df = run(P.B200DataFramePlugin(), {"records": rec, "basic_features": bf}, {"wave_source": "records"})
    w = float(A["pair_group_window_ns"])
    ev = run(P.B200GroupedEventsPlugin(), {"df": df}, {"time_window_ns": w})
    assert np.array_equal(np.concatenate(list(ev["timestamps"])), A["pair_ev_timestamps"])
    assert np.array_equal(ev["dt/ns"].to_numpy(), A["pair_ev_dt_ns"])
    for name in ("a", "b", "c"):
        nch, start = (int(v) for v in A[f"pair_{name}_nch_start"])
        for strip_attrs in (False, True):  # with and without the CSR side channel of the B200 df_events plugin
            src = ev.copy()
            if strip_attrs:
                src.attrs.clear()
            paired = P.B200PairedEventsPlugin().compute(Ctx({"n_channels": nch, "start_channel_slice": start, "time_window_ns": float(A[f"pair_{name}_tw"])},
                                                            {"df_events": src}), "run")
            assert np.array_equal(paired.index.to_numpy(), A[f"pair_{name}_index"]), name
            assert np.array_equal(paired["delta_t"].to_numpy(), A[f"pair_{name}_delta_t"]), name
            for i in range(nch):
                for kind in ("area", "height"):
                    col = f"{kind}_ch{start + i}"
                    assert str(paired[col].dtype) == str(A[f"pair_{name}_{col}_dtype"]), (name, col)
                    assert np.array_equal(paired[col].to_numpy(), A[f"pair_{name}_{col}"], equal_nan=True), (name, col)
            assert np.array_equal(paired["n_hits"].to_numpy(), A[f"pair_{name}_n_hits"])


def test_s1s2_plugin(P, golden):
    from after_cases import S1S2_CASES, load_after

    A = load_after()
    ww, feats = golden["ww_default"], A["s1s2_in_features"]
    for name, conf in S1S2_CASES.items():
        out = run(P.B200S1S2ClassifierPlugin(), {"waveform_width": ww, "basic_features": feats}, dict(conf))
        assert_rows_match(out, A[f"s1s2_{name}"], what=name, float_exact=("width_ns", "width_samples", "height", "area"))
    # sizes around the 128-row staging block, and the unaligned tail
    for n in (1, 127, 128, 129, 300):
        out = run(P.B200S1S2ClassifierPlugin(), {"waveform_width": ww[:n], "basic_features": feats}, dict(S1S2_CASES["area_prefer_s1"]))
        assert_rows_match(out, A["s1s2_area_prefer_s1"][:n], what=f"n={n}", float_exact=("width_ns", "width_samples", "height", "area"))
    assert len(run(P.B200S1S2ClassifierPlugin(), {"waveform_width": ww[:0], "basic_features": feats})) == 0
    with pytest.raises(ValueError, match="No S1/S2 criteria"):
        run(P.B200S1S2ClassifierPlugin(), {"waveform_width": ww, "basic_features": feats}, {"strict": True})
    with pytest.raises(ValueError, match="range must be a tuple"):
        run(P.B200S1S2ClassifierPlugin(), {"waveform_width": ww, "basic_features": feats}, {"s1_width_range": [1, 2]})
