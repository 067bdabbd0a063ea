"""CPU: the numpy oracle against golden vectors produced by the live reference
(tests/golden/make_golden.py).  Integer fields bit-exact; floats rel 1e-5 / abs 1e-3."""

import numpy as np
import pytest

from conftest import assert_rows_match
from oracle import np_oracle as O


def test_build_records(golden):
    rec, pool = O.build_records(golden["raw_timestamps_ps"], golden["raw_boards"], golden["raw_channels"],
                                golden["raw_samples"], dt_ns=2)
    want = golden["records"]
    for name in want.dtype.names:
        assert np.array_equal(rec[name], want[name], equal_nan=(want[name].dtype.kind == "f")), name
    assert np.array_equal(pool, golden["wave_pool"])


def test_basic_features(golden):
    rec, pool = golden["records"], golden["wave_pool"]
    assert_rows_match(O.basic_features(rec, pool), golden["bf_default"], what="bf_default",
                      float_exact=("height", "amp", "area", "max_abs_diff"))
    assert_rows_match(O.basic_features(rec, pool, height_range=(0, None), area_range=(100, -50)),
                      golden["bf_fullrange"], what="bf_fullrange", float_exact=("height", "amp", "area", "max_abs_diff"))
    for pol in ("negative", "positive"):
        r2 = rec.copy()
        r2["polarity"] = pol
        assert_rows_match(O.basic_features(r2, pool, height_range=(0, None)), golden[f"bf_{pol}"], what=pol,
                          float_exact=("height", "amp", "area", "max_abs_diff"))
    assert_rows_match(O.basic_features(rec, pool, fixed_baseline={(0, 1): 8000.5, (0, 3): 7990.0}),
                      golden["bf_fixed"], what="bf_fixed", float_exact=("height", "amp", "area", "max_abs_diff"))


def test_threshold_hits(golden):
    rec, pool = golden["records"], golden["wave_pool"]
    fx = ("height", "integral", "width", "rise_time", "fall_time")
    assert_rows_match(O.threshold_hits(rec, pool, threshold=15.0), golden["hits_thr15"], what="thr15", float_exact=fx)
    assert_rows_match(O.threshold_hits(rec, pool, threshold=12.0, thresholds={(0, 2): 40.0}, left_extension=5,
                                       right_extension=0), golden["hits_chan"], what="chan", float_exact=fx)
    rpos = rec.copy()
    rpos["polarity"] = "positive"
    assert_rows_match(O.threshold_hits(rpos, pool, threshold=-20.0), golden["hits_positive"], what="pos", float_exact=fx)


def test_ragged(golden):
    rr, rp = golden["rag_records"], golden["rag_pool"]
    assert_rows_match(O.basic_features(rr, rp, height_range=(5, -5)), golden["rag_bf"], what="rag_bf",
                      float_exact=("height", "amp", "area", "max_abs_diff"))
    assert_rows_match(O.threshold_hits(rr, rp, threshold=15.0, left_extension=3, right_extension=4),
                      golden["rag_hits"], what="rag_hits", float_exact=("height", "integral", "width"))
    assert_rows_match(O.width_integral(rr, rp), golden["rag_wint"], what="rag_wint")


def _butter(order, lo, hi, fs):
    from scipy.signal import butter

    return butter(order, [lo, hi], btype="band", output="sos", fs=fs)


def test_filters(golden):
    rs, ps = golden["filt_records"], golden["filt_pool"]
    sg = {"filter_type": "SG", "sg_window_size": 11, "sg_poly_order": 2}
    got = O.wave_pool_filtered(rs, ps, configs={}, default=sg)
    assert np.allclose(got, golden["filt_sg"], rtol=1e-5, atol=1e-3)
    # interior samples are bit-exact
    g2, w2 = got.reshape(-1, 800), golden["filt_sg"].reshape(-1, 800)
    assert np.array_equal(g2[:, 5:-5], w2[:, 5:-5])
    got = O.wave_pool_filtered(rs, ps, configs={}, default={"filter_type": "SG", "sg_window_size": 21, "sg_poly_order": 3})
    assert np.allclose(got, golden["filt_sg_21_3"], rtol=1e-5, atol=1e-3)
    bw = {"filter_type": "BW", "sos": _butter(4, 0.01, 0.1, 0.5)}
    got = O.wave_pool_filtered(rs, ps, configs={}, default=bw)
    assert np.array_equal(got, golden["filt_bw"])  # no-FMA serial restatement is bit-exact
    got = O.wave_pool_filtered(rs, ps, configs={(0, 1): {"filter_type": "BW", "sos": _butter(2, 0.02, 0.2, 1.0)}}, default=sg)
    assert np.allclose(got, golden["filt_mixed"], rtol=1e-5, atol=1e-3)


def test_zi_matches_scipy():
    from scipy.signal import sosfilt_zi

    sos = _butter(4, 0.01, 0.1, 0.5)
    assert np.allclose(O.sos_steady_state(sos), sosfilt_zi(sos), rtol=1e-12, atol=1e-13)


def test_features_on_filtered_pool(golden):
    rs = golden["filt_records"]
    fp = golden["filt_sg"]
    assert_rows_match(O.basic_features(rs, fp), golden["filt_bf"], what="filt_bf")
    assert_rows_match(O.threshold_hits(rs, fp, threshold=15.0), golden["filt_hits"], what="filt_hits")
    rn = rs.copy()
    rn["polarity"] = "negative"
    assert_rows_match(O.basic_features(rn, fp, height_range=(0, None)), golden["filt_bf_negative"], what="filt_bf_neg")


def test_width_integral(golden):
    rec, pool = golden["records"], golden["wave_pool"]
    assert_rows_match(O.width_integral(rec[:200], pool), golden["wint_default"], what="wint")
    rn = rec[:200].copy()
    rn["polarity"] = "negative"
    assert_rows_match(O.width_integral(rn, pool, q_low=0.2, q_high=0.8, dt=2.0), golden["wint_negative"], what="wint_neg")


def test_waveform_width(golden):
    rec, pool, hits = golden["ww_records"], golden["ww_pool"], golden["ww_hit"]
    waves = pool.reshape(len(rec), 800).view(np.int16)
    fx = ("rise_time", "fall_time", "total_width", "rise_time_samples", "fall_time_samples", "total_width_samples", "peak_height")
    assert len(golden["ww_default"]) > 50
    assert_rows_match(O.waveform_width(hits, rec["record_id"], waves), golden["ww_default"], what="ww", float_exact=fx)
    assert_rows_match(O.waveform_width(hits, rec["record_id"], waves, rise_high=0.5, fall_high=0.5, sampling_rate=0.25),
                      golden["ww_50"], what="ww50", float_exact=fx)
    assert_rows_match(O.waveform_width(hits, rec["record_id"], waves, interpolation=False), golden["ww_nointerp"],
                      what="ww_nointerp", float_exact=fx)
    fw = golden["ww_filtered_pool"].reshape(len(rec), 800)
    assert_rows_match(O.waveform_width(hits, rec["record_id"], fw), golden["ww_filtered"], what="ww_filt", float_exact=fx)


def test_hit_merge_and_grouping(golden):
    h = golden["hits_thr15"]
    for tag, kw in (("m0", {}), ("m50", {"merge_gap_ns": 50.0, "max_total_width_ns": 400.0})):
        cl, mg, cp = O.hit_merge(h, **kw)
        assert_rows_match(cl, golden[f"{tag}_clusters"], what=f"{tag}_clusters")
        assert_rows_match(mg, golden[f"{tag}_merged"], what=f"{tag}_merged")
        assert_rows_match(cp, golden[f"{tag}_components"], what=f"{tag}_components")
    mg = golden["m0_merged"]
    for wname, w in (("w100", 100.0), ("w0", 0.0), ("w2000", 2000.0)):
        ev = O.group_hit_windows(mg, w)
        assert np.array_equal(ev["t_min"], golden[f"hg_{wname}_t_min"])
        assert np.array_equal(ev["t_max"], golden[f"hg_{wname}_t_max"])
        assert np.array_equal(ev["n_hits"], golden[f"hg_{wname}_n_hits"])
        assert np.allclose(ev["dt_ns"], golden[f"hg_{wname}_dt_ns"], rtol=0, atol=0)
        m = ev["members"]
        assert np.array_equal(mg["record_id"][m], golden[f"hg_{wname}_record_ids"])
        assert np.array_equal(mg["timestamp"][m], golden[f"hg_{wname}_timestamps"])
        assert np.array_equal(mg["channel"][m], golden[f"hg_{wname}_channels"])


def test_group_time_window(golden):
    bf = golden["bf_default"]
    for wname, w in (("w100", 100.0), ("w30000", 30000.0)):
        ev = O.group_time_window(bf["timestamp"], bf["channel"], w)
        for flavour in ("nb", "np"):
            tag = f"ge_{wname}_{flavour}"
            assert np.array_equal(ev["t_min"], golden[f"{tag}_t_min"])
            assert np.array_equal(ev["t_max"], golden[f"{tag}_t_max"])
            assert np.array_equal(ev["n_hits"], golden[f"{tag}_n_hits"])
            assert np.array_equal(bf["timestamp"][ev["members"]], golden[f"{tag}_timestamps"])
            assert np.array_equal(bf["channel"][ev["members"]], golden[f"{tag}_channels"])


def v1725_cases():
    import os

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "v1725_golden.npz"), allow_pickle=False)
    for tag in ("a", "b", "cut"):
        k = int(g[f"{tag}_nfiles"])
        blobs = [g[f"{tag}_blob{i}"].tobytes() for i in range(k)]
        names = [str(g[f"{tag}_name{i}"]) for i in range(k)]
        yield tag, blobs, names, int(g[f"{tag}_dt_ns"]), g[f"{tag}_records"], g[f"{tag}_pool"]


def test_v1725_ingest_matches_reference():
    for tag, blobs, names, dt_ns, want_rec, want_pool in v1725_cases():
        rec, pool = O.build_records_from_v1725(blobs, names, dt_ns)
        assert np.array_equal(pool, want_pool), tag
        assert_rows_match(rec, want_rec, what=f"v1725 {tag}", float_exact=("baseline",))
        assert np.all(np.isnan(rec["baseline_upstream"]))


def hit_cases():
    """(tag, waves, meta, source, options, reference rows) for the `hit` golden file."""
    import os

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "hit_golden.npz"), allow_pickle=False)
    rec, pool, fpool = g["records"], g["pool"], g["filtered_pool"]
    n, L = len(rec), 800
    w_i16 = pool.view(np.int16).reshape(n, L)
    w_f32 = fpool.reshape(n, L)
    aos = [("filt_default", w_f32, {}),
           ("filt_lowcut", w_f32, {"height": 3.0, "prominence": 0.5, "width": 2, "distance": 6, "height_window_extension": 1}),
           ("filt_thr", w_f32, {"height": 5.0, "threshold": 0.5, "width": 1}),
           ("filt_diffheight", w_f32, {"height": 10.0, "height_method": "diff"}),
           ("st_default", w_i16, {"height": 12.0, "width": 2}),
           ("st_level", w_i16, {"use_derivative": False, "height": 20.0, "prominence": 4.0, "width": 3})]
    for tag, w, kw in aos:
        yield tag, w, rec, "aos", kw, g[tag]
    sig = -(pool.reshape(n, L).astype(np.float32) - rec["baseline"].astype(np.float32)[:, None])
    yield "rec_default", (-sig).astype(np.float64) * -1.0, rec, "records", {"height": 12.0, "width": 2}, g["rec_default"]
    prec, ppool = g["pos_records"], g["pos_pool"]
    m = len(prec)
    s = ppool.reshape(m, L).astype(np.float32) - prec["baseline"].astype(np.float32)[:, None]  # signals() negates positive pulses ...
    psig = s.astype(np.float64)                                                                 # ... and the plugin negates again
    yield "pos_level", psig, prec, "records", {"use_derivative": False, "height": 30.0, "prominence": 5.0, "width": 2}, g["pos_level"]
    yield "pos_deriv", psig, prec, "records", {"height": 8.0, "width": 2}, g["pos_deriv"]


def test_hit_find_peaks_matches_reference():
    for tag, waves, meta, source, kw, want in hit_cases():
        got = O.hit_find_peaks(list(waves), meta, source=source, **kw)
        assert_rows_match(got, want, what=f"hit {tag}", float_exact=("height", "edge_start", "edge_end"))


def stream_cases():
    import os

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "hit_golden.npz"), allow_pickle=False)
    rec, fpool = g["records"], g["filtered_pool"]
    w = fpool.reshape(len(rec), 800)
    for tag, kw in (("stream_default", {"height": 10.0}),
                    ("stream_minmax", {"height": 10.0, "height_method": "minmax", "minmax_window_expand": 3, "width": 2}),
                    ("stream_level", {"use_derivative": False, "height": 25.0, "prominence": 4.0, "width": 3})):
        yield tag, rec, w, kw, g[tag], g[tag + "_bounds"]


def test_stream_find_peaks_matches_reference():
    """The streaming plugin emits one chunk per channel here (80 records < chunk_size): rows in channel order."""
    for tag, rec, w, kw, want, _bounds in stream_cases():
        parts = []
        for ch in np.unique(rec["channel"]):
            sel = rec["channel"] == ch
            parts.append(O.stream_find_peaks(list(w[sel]), rec[sel], **kw))
        assert_rows_match(np.concatenate(parts), want, what=tag, float_exact=("height", "edge_start", "edge_end"))


def ragged_parts_case():
    import os

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "v1725_golden.npz"), allow_pickle=False)
    blocks = [g[f"rag_samples{c}"] for c in range(3)]
    ts = np.concatenate([g[f"rag_ts{c}"] for c in range(3)])
    boards = np.concatenate([np.full(len(b), c % 2, dtype=np.int16) for c, b in enumerate(blocks)])
    chans = np.concatenate([np.full(len(b), c, dtype=np.int16) for c, b in enumerate(blocks)])
    return ts, boards, chans, blocks, g["rag_records"], g["rag_pool"]


def test_records_from_parts_of_different_widths():
    ts, boards, chans, blocks, want_rec, want_pool = ragged_parts_case()
    rec, pool = O.build_records_ragged(ts, boards, chans, blocks, dt_ns=2)
    assert np.array_equal(pool, want_pool)
    assert_rows_match(rec, want_rec, what="ragged parts", float_exact=("baseline",))


def grouping50_cases(golden):
    import os

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "grouping_golden.npz"), allow_pickle=False)
    for wname, w in (("w100", 100.0), ("w0", 0.0), ("w5000", 5000.0)):
        yield w, {k[len(f"hg50_{wname}_"):]: g[k] for k in g.files if k.startswith(f"hg50_{wname}_")}


def check_grouping50(ev, mg, want):
    m = ev["members"]
    assert np.array_equal(ev["t_min"], want["t_min"]) and np.array_equal(ev["t_max"], want["t_max"])
    assert np.array_equal(ev["n_hits"], want["n_hits"])
    assert np.array_equal(mg["record_id"][m], want["record_ids"]) and np.array_equal(mg["timestamp"][m], want["timestamps"])
    assert np.array_equal(mg["channel"][m].astype(np.int64), want["channels"])
    assert np.array_equal(mg["sample_start"][m].astype(np.int64), want["sample_starts"])


def test_grouping_of_cross_record_merged_hits(golden):
    mg, cp, h = golden["m50_merged"], golden["m50_components"], golden["hits_thr15"]
    assert (mg["sample_start"] < 0).any()
    for w, want in grouping50_cases(golden):
        check_grouping50(O.group_hit_windows(mg, w, cp, h), mg, want)
    with pytest.raises(ValueError):
        O.group_hit_windows(mg, 100.0)


# ---- the step after the path: df / df_paired / s1_s2 (tests/golden/make_golden_after.py) -------------------


def test_df_columns_oracle():
    from after_cases import GAINS_RESOLVED, check_df_columns, load_after

    A = load_after()
    rec, bf = A["df_in_records"], A["df_in_features"]
    check_df_columns(O.df_columns(bf, rec["record_id"]), A, "df_plain", False)
    check_df_columns(O.df_columns(bf, rec["record_id"], GAINS_RESOLVED), A, "df_pe", True)
    assert np.isnan(A["df_pe_area_pe"]).any() and not np.isnan(A["df_pe_area_pe"]).all()
    st_feat = bf.copy()
    st_feat["board"] = 0
    check_df_columns(O.df_columns(st_feat, None), A, "df_st", False)


def test_s1s2_oracle(golden):
    from after_cases import S1S2_CASES, load_after

    A = load_after()
    for name, conf in S1S2_CASES.items():
        got = O.s1s2_classify(golden["ww_default"], A["s1s2_in_features"], **conf)
        assert_rows_match(got, A[f"s1s2_{name}"], what=name, float_exact=("width_ns", "width_samples", "height", "area"))


def test_pair_events_oracle():
    from after_cases import check_pair, load_after

    A = load_after()
    off = A["pair_ev_offsets"]
    for name in ("a", "b", "c"):
        nch = int(A[f"pair_{name}_nch_start"][0])
        out = O.pair_events(off, A["pair_ev_timestamps"], A["pair_ev_areas"], A["pair_ev_heights"], A["pair_ev_dt_ns"],
                            float(A[f"pair_{name}_tw"]), nch)
        check_pair(out, off, A, name)
