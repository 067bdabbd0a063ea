"""GPU parity of the remaining hot-path rows through the C-ABI: records builder (K1),
wave_pool_filtered (SG / BW), waveform_width, waveform_width_integral, hit merge ordering and
event grouping (K4), against the live-reference golden vectors and the numpy oracle."""

import numpy as np
import pytest

import known_answers as K
from conftest import assert_rows_match

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from waveformanalysis_b200 import ops

    return ops


def test_build_records_golden(ops, golden):
    rec, pool = ops.build_records(golden["raw_timestamps_ps"], golden["raw_boards"], golden["raw_channels"], golden["raw_samples"], dt_ns=2)
    want = golden["records"]
    for name in want.dtype.names:
        assert np.array_equal(rec[name], want[name], equal_nan=(want[name].dtype.kind == "f")), name
    assert np.array_equal(pool, golden["wave_pool"])


def test_build_records_ties_and_odd_length(ops):
    from oracle import np_oracle as O

    rng = np.random.default_rng(3)
    n, L = 5000, 123  # odd length: scalar copy path; heavy timestamp ties exercise the tie-break keys
    ts = rng.integers(0, 40, size=n).astype(np.int64) * 2000 - 20000
    boards = rng.integers(-1, 3, size=n).astype(np.int16)
    chans = rng.integers(0, 6, size=n).astype(np.int16)
    samples = rng.integers(0, 16384, size=(n, L)).astype(np.int16)
    rec, pool = ops.build_records(ts, boards, chans, samples, dt_ns=4, baseline_window=(3, 200), epoch_ns=1_700_000_000)
    wrec, wpool = O.build_records(ts, boards, chans, samples, dt_ns=4, baseline_window=(3, 200), epoch_ns=1_700_000_000)
    for name in wrec.dtype.names:
        assert np.array_equal(rec[name], wrec[name], equal_nan=(wrec[name].dtype.kind == "f")), name
    assert np.array_equal(pool, wpool)
    base = rng.normal(8000, 5, size=n)
    rec, _ = ops.build_records(ts, boards, chans, samples, dt_ns=4, baselines=base)
    wrec, _ = O.build_records(ts, boards, chans, samples, dt_ns=4, baselines=base)
    assert np.array_equal(rec["baseline"], wrec["baseline"])


def test_dual_baseline_known_answer(ops):
    samples, want = K.dual_baseline_case()
    n = len(samples)
    rec, _ = ops.build_records(np.arange(n) * 1000, np.zeros(n), np.zeros(n), samples, dt_ns=2)
    assert np.array_equal(rec["baseline"], want)
    assert np.isnan(rec["baseline_upstream"]).all()


def test_sort_pairs(ops):
    rng = np.random.default_rng(0)
    keys = rng.integers(-2**62, 2**62, size=100_003)
    keys[::7] = keys[0]
    k, v = ops.sort_pairs(keys, np.arange(len(keys)))
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(v, order) and np.array_equal(k, keys[order])
    # passes over digits that are the same in every key are skipped: all keys equal (no pass at all), one / two /
    # three differing digits (odd and even numbers of passes end in the output buffers), the sign bit only
    cases = [np.full(5000, 123456789, dtype=np.int64), rng.integers(0, 256, 70_001), rng.integers(0, 1 << 16, 70_001) << 24,
             (rng.integers(0, 1 << 24, 33_333) << 8) + (7 << 40), rng.integers(-1, 1, 4097) * (1 << 62), np.array([5], dtype=np.int64),
             1_700_000_000_000_000_000 + rng.integers(0, 1 << 37, 200_000)]
    for keys in cases:
        keys = np.asarray(keys, dtype=np.int64)
        k, v = ops.sort_pairs(keys, np.arange(len(keys))[::-1].copy())
        order = np.argsort(keys, kind="stable")
        assert np.array_equal(k, keys[order]) and np.array_equal(v, np.arange(len(keys))[::-1][order])


def _butter(order, lo, hi, fs):
    from scipy.signal import butter

    return butter(order, [lo, hi], btype="band", output="sos", fs=fs)


def test_filters_golden(ops, golden, monkeypatch):
    rs, ps = golden["filt_records"], golden["filt_pool"]
    sg = {"filter_type": "SG", "sg_window_size": 11, "sg_poly_order": 2}
    got = ops.filter_pool(rs, ps, configs={}, default=sg)
    assert np.allclose(got, golden["filt_sg"], rtol=1e-5, atol=1e-3)
    g2, w2 = got.reshape(-1, 800), golden["filt_sg"].reshape(-1, 800)
    assert np.array_equal(g2[:, 5:-5], w2[:, 5:-5])  # interior bit-exact with scipy
    got = ops.filter_pool(rs, ps, configs={}, default={"filter_type": "SG", "sg_window_size": 21, "sg_poly_order": 3})
    assert np.allclose(got, golden["filt_sg_21_3"], rtol=1e-5, atol=1e-3)
    # Butterworth: the default kernel uses fused multiply-adds and parks the forward pass as float32 (parity bar for
    # floats: rel 1e-5 / abs 1e-3); WFB_BW_EXACT=1 runs scipy's order of operations without contraction: bit-exact
    got = ops.filter_pool(rs, ps, configs={}, default={"filter_type": "BW", "sos": _butter(4, 0.01, 0.1, 0.5)})
    assert np.allclose(got, golden["filt_bw"], rtol=1e-5, atol=1e-3)
    assert np.abs(got - golden["filt_bw"]).max() < 2e-4
    monkeypatch.setenv("WFB_BW_EXACT", "1")
    got = ops.filter_pool(rs, ps, configs={}, default={"filter_type": "BW", "sos": _butter(4, 0.01, 0.1, 0.5)})
    assert np.array_equal(got, golden["filt_bw"])
    monkeypatch.delenv("WFB_BW_EXACT")
    got = ops.filter_pool(rs, ps, configs={(0, 1): {"filter_type": "BW", "sos": _butter(2, 0.02, 0.2, 1.0)}}, default=sg)
    assert np.allclose(got, golden["filt_mixed"], rtol=1e-5, atol=1e-3)


def test_filters_edge_cases(ops, golden):
    """reference tests/test_wave_pool_filtered_plugin.py:46-180: BW on a record shorter than the pad
    length is a no-op; SG window larger than the record shrinks; window <= poly is the identity."""
    from oracle import np_oracle as O

    rr, rp = golden["rag_records"], golden["rag_pool"]
    sg = {"filter_type": "SG", "sg_window_size": 11, "sg_poly_order": 2}
    bw = {"filter_type": "BW", "sos": _butter(4, 0.01, 0.1, 0.5)}
    for cfg in (sg, bw):
        got = ops.filter_pool(rr, rp, configs={}, default=cfg)
        want = O.wave_pool_filtered(rr, rp, configs={}, default=cfg)
        assert np.allclose(got, want, rtol=1e-5, atol=1e-3)
    short = rr["event_length"] <= 27
    assert short.any()
    got = ops.filter_pool(rr, rp, configs={}, default=bw)
    for i in np.flatnonzero(short & (rr["event_length"] > 0))[:20]:
        o, L = int(rr["wave_offset"][i]), int(rr["event_length"][i])
        assert np.array_equal(got[o : o + L], rp[o : o + L].astype(np.float32))


def test_waveform_width_golden(ops, golden):
    rec, pool, hits = golden["ww_records"], golden["ww_pool"], golden["ww_hit"]
    waves = pool.reshape(len(rec), 800).view(np.int16)
    fx = ("rise_time", "fall_time", "total_width", "rise_time_samples", "fall_time_samples", "total_width_samples", "peak_height")
    assert_rows_match(ops.waveform_width(hits, rec["record_id"], waves), golden["ww_default"], what="ww", float_exact=fx)
    assert_rows_match(ops.waveform_width(hits, rec["record_id"], waves, rise_high=0.5, fall_high=0.5, sampling_rate=0.25),
                      golden["ww_50"], what="ww50", float_exact=fx)
    assert_rows_match(ops.waveform_width(hits, rec["record_id"], waves, interpolation=False), golden["ww_nointerp"], what="ww_nointerp", float_exact=fx)
    fw = golden["ww_filtered_pool"].reshape(len(rec), 800)
    assert_rows_match(ops.waveform_width(hits, rec["record_id"], fw), golden["ww_filtered"], what="ww_filt")


def test_waveform_width_drops_and_unknown_records(ops, golden):
    rec, pool, hits = golden["ww_records"], golden["ww_pool"], golden["ww_hit"].copy()
    waves = pool.reshape(len(rec), 800).view(np.int16)
    hits["record_id"][::5] = 10**9  # no such waveform -> row dropped (waveform_width.py:166-167)
    hits["position"][1::5] = 900    # position >= len -> dropped (:244-245)
    from oracle import np_oracle as O

    want = O.waveform_width(hits, rec["record_id"], waves)
    assert 0 < len(want) < len(hits)
    assert_rows_match(ops.waveform_width(hits, rec["record_id"], waves), want, what="ww_drop")


def test_width_integral_golden(ops, golden):
    rec, pool = golden["records"], golden["wave_pool"]
    fx = ("t_low", "t_high", "width", "t_low_samples", "t_high_samples", "width_samples", "q_total")
    assert_rows_match(ops.width_integral(rec[:200], pool), golden["wint_default"], what="wint", float_exact=fx)
    rn = rec[:200].copy()
    rn["polarity"] = "negative"
    assert_rows_match(ops.width_integral(rn, pool, q_low=0.2, q_high=0.8, dt=2.0), golden["wint_negative"], what="wint_neg", float_exact=fx)
    rr, rp = golden["rag_records"], golden["rag_pool"]
    assert_rows_match(ops.width_integral(rr, rp), golden["rag_wint"], what="rag_wint", float_exact=fx)
    with pytest.raises(ValueError):
        ops.width_integral(rec[:4], pool, q_low=0.9, q_high=0.1)


def test_hit_merge_default_golden(ops, golden):
    cl, mg, cp = ops.hit_merge_default(golden["hits_thr15"])
    assert_rows_match(cl, golden["m0_clusters"], what="clusters")
    assert_rows_match(mg, golden["m0_merged"], what="merged", float_exact=("height", "integral", "width", "rise_time", "fall_time"))
    assert_rows_match(cp, golden["m0_components"], what="components")


FX_MERGED = ("height", "integral", "width", "rise_time", "fall_time")


def test_hit_merge_chain_golden(ops, golden):
    cl, mg, cp = ops.hit_merge(golden["hits_thr15"])
    assert_rows_match(cl, golden["m0_clusters"], what="clusters")
    assert_rows_match(mg, golden["m0_merged"], what="merged", float_exact=FX_MERGED)
    assert_rows_match(cp, golden["m0_components"], what="components")
    cl, mg, cp = ops.hit_merge(golden["hits_thr15"], merge_gap_ns=50.0, max_total_width_ns=400.0)
    assert_rows_match(cl, golden["m50_clusters"], what="m50 clusters")
    assert_rows_match(mg, golden["m50_merged"], what="m50 merged", float_exact=FX_MERGED)
    assert_rows_match(cp, golden["m50_components"], what="m50 components")


@pytest.mark.parametrize("gap,maxw", [(20.0, 10000.0), (200.0, 300.0), (1e6, 1e9), (5.0, 0.0)])
def test_hit_merge_chain_vs_oracle(ops, gap, maxw):
    """Noise-level threshold: long chains, clusters across records of a channel, equal heights, the width cut
    inside a gap-free piece, and one cluster per channel-dt run (huge gap)."""
    from oracle import np_oracle as O
    from waveformanalysis_b200 import engine
    from waveformanalysis_b200.synth import make_raw_run, records_from_raw

    raw = make_raw_run(6, 400, 500, seed=11)
    rec, pool = records_from_raw(raw)
    rec["dt"][rec["channel"] == 2] = 4  # a second sampling interval
    hits = engine.process_host(rec, pool, features=False, threshold=4.0)["hits"]
    hits["dt"][::97] = 8  # dt changes inside a channel break chains
    assert len(hits) > 20 * len(rec)
    got = ops.hit_merge(hits, merge_gap_ns=gap, max_total_width_ns=maxw)
    want = O.hit_merge(hits, merge_gap_ns=gap, max_total_width_ns=maxw)
    assert_rows_match(got[0], want[0], what="clusters")
    assert_rows_match(got[1], want[1], what="merged", float_exact=FX_MERGED)
    assert_rows_match(got[2], want[2], what="components")
    assert len(want[1]) < len(hits) or maxw == 0.0


def test_group_hit_windows_golden(ops, golden):
    mg = golden["m0_merged"]
    for wname, w in (("w100", 100.0), ("w0", 0.0), ("w2000", 2000.0)):
        ev = ops.group_hit_windows(mg, w)
        assert np.array_equal(ev["t_min"], golden[f"hg_{wname}_t_min"])
        assert np.array_equal(ev["t_max"], golden[f"hg_{wname}_t_max"])
        assert np.array_equal(ev["n_hits"], golden[f"hg_{wname}_n_hits"])
        assert np.array_equal(ev["dt_ns"], golden[f"hg_{wname}_dt_ns"])
        m = ev["members"]
        assert np.array_equal(mg["record_id"][m], golden[f"hg_{wname}_record_ids"])
        assert np.array_equal(mg["timestamp"][m], golden[f"hg_{wname}_timestamps"])
        assert np.array_equal(mg["channel"][m], golden[f"hg_{wname}_channels"])
    h, w, want = K.grouping_case()
    ev = ops.group_hit_windows(h, w)
    assert ev["t_min"].tolist() == want["t_min"] and ev["t_max"].tolist() == want["t_max"]
    assert ev["dt_ns"].tolist() == want["dt_ns"] and ev["n_hits"].tolist() == want["n_hits"]


def test_group_time_window_golden(ops, golden):
    bf = golden["bf_default"]
    for wname, w in (("w100", 100.0), ("w30000", 30000.0)):
        ev = ops.group_time_window(bf["timestamp"], bf["channel"], w)
        tag = f"ge_{wname}_nb"
        assert np.array_equal(ev["t_min"], golden[f"{tag}_t_min"])
        assert np.array_equal(ev["t_max"], golden[f"{tag}_t_max"])
        assert np.array_equal(ev["n_hits"], golden[f"{tag}_n_hits"])
        assert np.array_equal(bf["timestamp"][ev["members"]], golden[f"{tag}_timestamps"])
        assert np.array_equal(bf["channel"][ev["members"]], golden[f"{tag}_channels"])


def test_grouping_large_vs_oracle(ops):
    """Bigger seeded case: dense coincidences so that chains and anchored windows span many hits."""
    from oracle import np_oracle as O
    from waveformanalysis_b200 import engine
    from waveformanalysis_b200.synth import make_raw_run, records_from_raw

    raw = make_raw_run(8, 4000, 256, seed=11, coincidence_fraction=0.6)
    rec, pool = records_from_raw(raw)
    hits = engine.process_host(rec, pool, features=False, threshold=15.0)["hits"]
    assert len(hits) > 20000
    cl, mg, cp = ops.hit_merge_default(hits)
    wcl, wmg, wcp = O.hit_merge(hits)
    assert np.array_equal(cl, wcl) and np.array_equal(mg, wmg) and np.array_equal(cp, wcp)
    for w in (0.0, 100.0, 5000.0):
        got, want = ops.group_hit_windows(mg, w), O.group_hit_windows(mg, w)
        for k in ("t_min", "t_max", "n_hits", "offsets", "members", "event_of_hit"):
            assert np.array_equal(got[k], want[k]), (w, k)
    feats = engine.process_host(rec, pool, hits=False)["features"]
    for w in (100.0, 20000.0):
        got, want = ops.group_time_window(feats["timestamp"], feats["channel"], w), O.group_time_window(feats["timestamp"], feats["channel"], w)
        for k in ("t_min", "t_max", "n_hits", "offsets", "members"):
            assert np.array_equal(got[k], want[k]), (w, k)


def test_v1725_ingest_golden(ops):
    from test_oracle_golden import v1725_cases

    for tag, blobs, names, dt_ns, want_rec, want_pool in v1725_cases():
        rec, pool = ops.build_records_from_v1725(blobs, names, dt_ns)
        assert np.array_equal(pool, want_pool), tag
        assert_rows_match(rec, want_rec, what=f"v1725 {tag}", float_exact=("baseline",))


def test_v1725_ingest_large_vs_oracle(ops):
    """Several files, many timestamp ties across files and channels, ragged lengths; the records feed the
    fused kernel unchanged."""
    from oracle import np_oracle as O
    from waveformanalysis_b200 import engine
    from waveformanalysis_b200.synth import make_v1725_blob

    blobs = [make_v1725_blob(n_events=700, n_channels=16, seed=20 + k, lengths=(16, 600), tie_every=2, t0=3 * k) for k in range(4)]
    names = [f"run_b{k % 3}_seg{k}.bin" for k in range(4)]
    rec, pool = ops.build_records_from_v1725(blobs, names, 4)
    want_rec, want_pool = O.build_records_from_v1725(blobs, names, 4)
    assert len(rec) > 20000 and np.array_equal(pool, want_pool)
    assert_rows_match(rec, want_rec, what="v1725 large", float_exact=("baseline",))
    out = engine.process_host(rec, pool, threshold=25.0, signed_samples=False)
    assert_rows_match(out["features"], O.basic_features(rec, pool), what="v1725 features", float_exact=("height", "amp", "max_abs_diff"))
    assert_rows_match(out["hits"], O.threshold_hits(rec, pool, threshold=25.0), what="v1725 hits", float_exact=("height", "width", "rise_time", "fall_time"))
    assert ops.build_records_from_v1725([], [], 4)[0].shape == (0,)


def test_hit_find_peaks_golden(ops):
    """`hit` rows against the reference HitFinderPlugin (tests/golden/hit_golden.npz): filtered float32 rows,
    raw int16 rows, records source with unknown / positive polarity, every option the plugin passes."""
    import os

    from waveformanalysis_b200.dtypes import create_record_dtype

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "hit_golden.npz"), allow_pickle=False)
    rec, pool, fpool = g["records"], g["pool"], g["filtered_pool"]
    n, L = len(rec), 800
    st = np.zeros(n, dtype=create_record_dtype(L))
    for f in ("baseline", "baseline_upstream", "polarity", "timestamp", "record_id", "dt", "event_length", "board", "channel"):
        st[f] = rec[f]
    st["wave"] = pool.reshape(n, L).view(np.int16)
    stf_dtype = np.dtype([(name, (np.float32, (L,)) if name == "wave" else st.dtype.fields[name][0]) for name in st.dtype.names])
    stf = np.zeros(n, dtype=stf_dtype)
    for f in st.dtype.names:
        if f != "wave":
            stf[f] = st[f]
    stf["wave"] = fpool.reshape(n, L)
    fx = ("height", "edge_start", "edge_end")
    cases = [("filt_default", stf, {}),
             ("filt_lowcut", stf, {"height": 3.0, "prominence": 0.5, "width": 2, "distance": 6, "height_window_extension": 1}),
             ("filt_thr", stf, {"height": 5.0, "threshold": 0.5, "width": 1}),
             ("filt_diffheight", stf, {"height": 10.0, "height_method": "diff"}),
             ("st_default", st, {"height": 12.0, "width": 2}),
             ("st_level", st, {"use_derivative": False, "height": 20.0, "prominence": 4.0, "width": 3})]
    for tag, data, kw in cases:
        assert_rows_match(ops.find_peaks_waveforms(data, **kw), g[tag], what=f"hit {tag}", float_exact=fx)
    assert_rows_match(ops.find_peaks_records(rec, pool, height=12.0, width=2), g["rec_default"], what="hit rec_default", float_exact=fx)
    prec, ppool = g["pos_records"], g["pos_pool"]
    assert_rows_match(ops.find_peaks_records(prec, ppool, use_derivative=False, height=30.0, prominence=5.0, width=2), g["pos_level"],
                      what="hit pos_level", float_exact=fx)
    assert_rows_match(ops.find_peaks_records(prec, ppool, height=8.0, width=2), g["pos_deriv"], what="hit pos_deriv", float_exact=fx)


def test_hit_find_peaks_vs_oracle_ragged_and_plateaus(ops):
    """Ragged records, integer plateaus, a minimal distance > 2 on float data, truncated rows."""
    from oracle import np_oracle as O
    from waveformanalysis_b200.synth import make_ragged_records

    rec, pool = make_ragged_records(300, seed=5, min_len=1, max_len=500)
    waves = []
    for o, l, b, pol in zip(rec["wave_offset"], rec["event_length"], rec["baseline"], rec["polarity"]):
        s = pool[int(o):int(o) + int(l)].astype(np.float32) - np.float32(b)  # RecordsView.signals before polarity normalisation
        waves.append((s if pol == "positive" else -s).astype(np.float64))    # the plugin's -signals()
    for kw in ({"height": 6.0, "prominence": 2.0, "width": 1}, {"use_derivative": False, "height": 10.0, "prominence": 3.0, "width": 2}):
        want = O.hit_find_peaks(waves, rec, source="records", **kw)
        assert len(want) > 50
        assert_rows_match(ops.find_peaks_records(rec, pool, **kw), want, what="ragged hit", float_exact=("height", "edge_start", "edge_end"))


def test_build_records_ragged_golden(ops):
    """Parts of different waveform widths (one narrower than the baseline window) against the reference's
    per-part builder + merge."""
    from test_oracle_golden import ragged_parts_case

    ts, boards, chans, blocks, want_rec, want_pool = ragged_parts_case()
    rec, pool = ops.build_records_ragged(ts, boards, chans, blocks, dt_ns=2)
    assert np.array_equal(pool, want_pool)
    assert_rows_match(rec, want_rec, what="ragged parts", float_exact=("baseline",))


def test_group_hit_windows_cross_record_clusters(ops, golden):
    """hit_merged rows merged across records (sample window -1): windows from the component hits, then the
    device grouping; against the reference HitGroupedPlugin (grouping_golden.npz)."""
    from test_oracle_golden import check_grouping50, grouping50_cases

    mg, cp, h = golden["m50_merged"], golden["m50_components"], golden["hits_thr15"]
    for w, want in grouping50_cases(golden):
        check_grouping50(ops.group_hit_windows(mg, w, cp, h), mg, want)
    with pytest.raises(ValueError):
        ops.group_hit_windows(mg, 100.0)
