"""GPU parity: fused baseline -> threshold hits -> basic_features through the C-ABI
(wfb_process_host and wfb_features_hits) against the live-reference golden vectors and the
numpy oracle.  Integer fields bit-exact; float fields rel 1e-5 / abs 1e-3 (BASELINE.json)."""

import numpy as np
import pytest

import known_answers as K
from conftest import assert_rows_match

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from waveformanalysis_b200 import engine

    return engine


def both_paths(eng, rec, pool, **kw):
    """Run through the host pipeline and through the device-resident path; both must agree."""
    a = eng.process_host(rec, pool, **kw)
    kw2 = {k: v for k, v in kw.items() if k not in ("thresholds", "fixed_baselines", "chunk_records")}
    rules = eng.make_rules(kw.get("thresholds"), kw.get("fixed_baselines"))
    b = eng.DeviceRun.from_host(rec, pool).run_to_host(rules=rules, **kw2)
    if kw.get("features", True):
        assert np.array_equal(a["features"].view(np.uint8), b["features"].view(np.uint8))
    if kw.get("hits", True):
        assert np.array_equal(a["hits"].view(np.uint8), b["hits"].view(np.uint8))
    return a


FX_BF = ("height", "amp", "max_abs_diff")


def test_golden_basic_features(eng, golden):
    rec, pool = golden["records"], golden["wave_pool"]
    out = both_paths(eng, rec, pool, hits=False)
    assert_rows_match(out["features"], golden["bf_default"], what="bf_default", float_exact=FX_BF)
    out = both_paths(eng, rec, pool, hits=False, height_range=(0, None), area_range=(100, -50))
    assert_rows_match(out["features"], golden["bf_fullrange"], what="bf_fullrange", float_exact=FX_BF)
    for pol in ("negative", "positive"):
        r2 = rec.copy()
        r2["polarity"] = pol
        out = both_paths(eng, r2, pool, hits=False, height_range=(0, None))
        assert_rows_match(out["features"], golden[f"bf_{pol}"], what=pol, float_exact=FX_BF + ("area",))
    out = both_paths(eng, rec, pool, hits=False, fixed_baselines={(0, 1): 8000.5, (0, 3): 7990.0})
    assert_rows_match(out["features"], golden["bf_fixed"], what="bf_fixed", float_exact=FX_BF)


FX_HIT = ("height", "width", "rise_time", "fall_time")


def test_golden_threshold_hits(eng, golden):
    rec, pool = golden["records"], golden["wave_pool"]
    out = both_paths(eng, rec, pool, features=False, threshold=15.0)
    assert_rows_match(out["hits"], golden["hits_thr15"], what="thr15", float_exact=FX_HIT)
    out = both_paths(eng, rec, pool, features=False, threshold=12.0, thresholds={(0, 2): 40.0}, left_extension=5, right_extension=0)
    assert_rows_match(out["hits"], golden["hits_chan"], what="chan", float_exact=FX_HIT)
    rpos = rec.copy()
    rpos["polarity"] = "positive"
    out = both_paths(eng, rpos, pool, features=False, threshold=-20.0)
    assert_rows_match(out["hits"], golden["hits_positive"], what="pos", float_exact=FX_HIT)


def test_golden_fused_both(eng, golden):
    rec, pool = golden["records"], golden["wave_pool"]
    out = both_paths(eng, rec, pool, threshold=15.0, want_counts=True)
    assert_rows_match(out["features"], golden["bf_default"], what="bf", float_exact=FX_BF)
    assert_rows_match(out["hits"], golden["hits_thr15"], what="hits", float_exact=FX_HIT)
    want_counts = np.bincount(golden["hits_thr15"]["record_id"], minlength=len(rec))
    assert np.array_equal(out["counts"], want_counts)


def test_golden_ragged(eng, golden):
    rr, rp = golden["rag_records"], golden["rag_pool"]
    out = both_paths(eng, rr, rp, height_range=(5, -5), threshold=15.0, left_extension=3, right_extension=4)
    assert_rows_match(out["features"], golden["rag_bf"], what="rag_bf", float_exact=FX_BF)
    assert_rows_match(out["hits"], golden["rag_hits"], what="rag_hits", float_exact=("height", "width"))


def test_golden_filtered_pool(eng, golden):
    rs, fp = golden["filt_records"], golden["filt_sg"]
    out = both_paths(eng, rs, fp, threshold=15.0)
    assert_rows_match(out["features"], golden["filt_bf"], what="filt_bf")
    assert_rows_match(out["hits"], golden["filt_hits"], what="filt_hits")
    rn = rs.copy()
    rn["polarity"] = "negative"
    out = both_paths(eng, rn, fp, hits=False, height_range=(0, None))
    assert_rows_match(out["features"], golden["filt_bf_negative"], what="filt_bf_neg")


def test_known_answers(eng):
    r, pool, cfg, want = K.bf_records_view()
    out = eng.process_host(r, pool, hits=False, **cfg)["features"]
    assert np.isclose(out["height"][0], want["height0"]) and np.isclose(out["amp"][0], want["amp0"])
    assert np.isclose(out["max_abs_diff"][0], want["max_abs_diff0"]) and out["board"].tolist() == want["boards"]
    r, pool, cfg, want = K.bf_fixed_baseline()
    out = eng.process_host(r, pool, hits=False, height_range=cfg["height_range"], area_range=cfg["area_range"],
                           fixed_baselines=cfg["fixed_baseline"])["features"]
    assert np.isclose(out["height"][0], want["height0"]) and np.isclose(out["area"][0], want["area0"])
    r, pool, cfg, want = K.bf_filtered_pool()
    out = eng.process_host(r, pool, hits=False, **cfg)["features"]
    for k, v in want.items():
        np.testing.assert_allclose(out[k], v)
    for case in (K.hits_two_regions(), K.hits_extension(), K.hits_records_view(), K.hits_rise_fall(False), K.hits_rise_fall(True)):
        r, pool, cfg, want = case
        out = eng.process_host(r, pool, features=False, **cfg)["hits"]
        for k, v in want.items():
            np.testing.assert_array_equal(out[k], np.asarray(v, dtype=out[k].dtype))


def test_empty_inputs(eng):
    from waveformanalysis_b200.dtypes import RECORDS_DTYPE

    out = eng.process_host(np.zeros(0, RECORDS_DTYPE), np.zeros(0, np.uint16))
    assert len(out["features"]) == 0 and len(out["hits"]) == 0
    rec = np.zeros(3, RECORDS_DTYPE)  # zero-length records
    rec["record_id"] = np.arange(3)
    rec["dt"] = 2
    out = eng.process_host(rec, np.zeros(0, np.uint16))
    assert len(out["features"]) == 3 and len(out["hits"]) == 0
    assert np.all(out["features"]["height"] == 0) and np.all(out["features"]["event_index"] == np.arange(3))


def test_out_of_bounds_records_raise(eng):
    from waveformanalysis_b200.dtypes import RECORDS_DTYPE

    rec = np.zeros(2, RECORDS_DTYPE)
    rec["event_length"] = 16
    rec["wave_offset"] = [0, 8]
    rec["dt"] = 2
    with pytest.raises(ValueError, match="outside wave_pool"):
        eng.process_host(rec, np.zeros(20, np.uint16))
    with pytest.raises(ValueError, match="outside wave_pool"):
        run = eng.DeviceRun.from_host(rec, np.zeros(20, np.uint16))
        run.run_to_host()


@pytest.mark.parametrize("n_samples", [800, 250, 1031])
def test_random_vs_oracle(eng, n_samples):
    from oracle import np_oracle as O
    from waveformanalysis_b200.synth import make_raw_run, records_from_raw

    raw = make_raw_run(16, 512, n_samples, seed=1234 + n_samples)
    rec, pool = records_from_raw(raw)
    rec["polarity"][::7] = "negative"
    rec["polarity"][3::11] = "positive"
    kw = dict(height_range=(40, 90), area_range=(0, None), threshold=15.0)
    want_f = O.basic_features(rec, pool, height_range=kw["height_range"], area_range=kw["area_range"])
    want_h = O.threshold_hits(rec, pool, threshold=15.0)
    # small chunks exercise the chunked pipeline, the carried hit offsets and pool_base handling
    out = both_paths(eng, rec, pool, chunk_records=1000, **kw)
    assert_rows_match(out["features"], want_f, what="features", float_exact=FX_BF)
    assert_rows_match(out["hits"], want_h, what="hits", float_exact=FX_HIT)
    assert len(want_h) > 1000


def test_many_hits_per_record_and_small_capacity(eng):
    """Noise-level threshold: dozens of hits per record overflow the per-warp staging area (the
    kernel re-scans those records) and the first output buffer (the host retries)."""
    from oracle import np_oracle as O
    from waveformanalysis_b200.synth import make_raw_run, records_from_raw

    raw = make_raw_run(4, 300, 800, seed=5)
    rec, pool = records_from_raw(raw)
    want = O.threshold_hits(rec, pool, threshold=2.0, left_extension=1, right_extension=7)
    assert len(want) > 40 * len(rec)
    out = eng.process_host(rec, pool, features=False, threshold=2.0, left_extension=1, right_extension=7, hit_cap=100)
    assert_rows_match(out["hits"], want, what="dense hits", float_exact=FX_HIT)
    run = eng.DeviceRun.from_host(rec, pool)
    got = run.run_to_host(features=False, threshold=2.0, left_extension=1, right_extension=7, hit_cap=64)
    assert np.array_equal(got["hits"].view(np.uint8), out["hits"].view(np.uint8))


def test_device_synth_matches_oracle(eng):
    """The on-device generator used by bench.py: copy a slice back and check it with the oracle."""
    from oracle import np_oracle as O

    run = eng.DeviceRun.synth(20000, 800, 16, seed=99, with_rows=True)
    rec, pool = run.records_to_host(), run.pool_to_host()
    assert np.all(np.diff(rec["timestamp"]) > 0) and np.array_equal(rec["wave_offset"], np.arange(20000) * 800)
    assert np.array_equal(rec["baseline"], pool.reshape(-1, 800)[:, :40].astype(np.float64).mean(axis=1))
    got = run.run_to_host(threshold=15.0)
    assert_rows_match(got["features"], O.basic_features(rec, pool), what="synth features", float_exact=FX_BF)
    assert_rows_match(got["hits"], O.threshold_hits(rec, pool, threshold=15.0), what="synth hits", float_exact=FX_HIT)
    # linearity / idempotence properties usable at full bench size: same input -> identical bytes
    again = run.run_to_host(threshold=15.0)
    assert np.array_equal(got["hits"].view(np.uint8), again["hits"].view(np.uint8))


@pytest.mark.parametrize("variant", ["global", "staged", "auto"])
def test_kernel_variants_agree(eng, golden, variant, monkeypatch):
    """All data-movement variants of the fused kernel (TMA-staged slot ring,
    and the global-memory accessor used for very long records) produce identical bytes."""
    monkeypatch.setenv("WFB_FUSED_VARIANT", variant)
    rec, pool = golden["records"], golden["wave_pool"]
    out = eng.DeviceRun.from_host(rec, pool).run_to_host(threshold=15.0)
    assert_rows_match(out["features"], golden["bf_default"], what="bf", float_exact=FX_BF)
    assert_rows_match(out["hits"], golden["hits_thr15"], what="hits", float_exact=FX_HIT)
    rr, rp = golden["rag_records"], golden["rag_pool"]
    out = eng.DeviceRun.from_host(rr, rp).run_to_host(height_range=(5, -5), threshold=15.0, left_extension=3, right_extension=4)
    assert_rows_match(out["features"], golden["rag_bf"], what="rag_bf", float_exact=FX_BF)
    assert_rows_match(out["hits"], golden["rag_hits"], what="rag_hits", float_exact=("height", "width"))
    fp = golden["filt_sg"]
    out = eng.DeviceRun.from_host(golden["filt_records"], fp).run_to_host(threshold=15.0)
    assert_rows_match(out["features"], golden["filt_bf"], what="filt_bf")
    assert_rows_match(out["hits"], golden["filt_hits"], what="filt_hits")


def test_long_records_use_global_path(eng):
    from oracle import np_oracle as O
    from waveformanalysis_b200.synth import make_raw_run, records_from_raw

    raw = make_raw_run(2, 6, 120_000, seed=3)  # 240 KB per record: does not fit a shared-memory slot
    rec, pool = records_from_raw(raw)
    out = eng.DeviceRun.from_host(rec, pool).run_to_host(threshold=15.0, height_range=(0, None))
    assert_rows_match(out["features"], O.basic_features(rec, pool, height_range=(0, None)), what="long bf", float_exact=FX_BF)
    assert_rows_match(out["hits"], O.threshold_hits(rec, pool, threshold=15.0), what="long hits", float_exact=FX_HIT)


@pytest.mark.parametrize("knobs", [{"WFB_LPR_SC": "4"}, {"WFB_LPR_SC": "8"}, {"WFB_LPR_SC": "12"}, {"WFB_LPR_SC": "16"}, {"WFB_LPR_SC": "28"},
                                   {"WFB_LPR_POOL": "24"}, {"WFB_LPR_POOL": "0", "WFB_LPR_SC": "4"}, {"WFB_LPR_NO_TMAP": "1"},
                                   {"WFB_LPR_NO_TMAP": "1", "WFB_LPR_SC": "16", "WFB_LPR_POOL": "7"}, {"WFB_FUSED_VARIANT": "staged", "WFB_FUSED_SLOTS": "2"},
                                   {"WFB_FUSED_VARIANT": "staged", "WFB_FUSED_SLOTS": "3"}, {"WFB_LPR_IMPL": "chunk"},
                                   {"WFB_LPR_IMPL": "chunk", "WFB_LPR_SC": "8", "WFB_LPR_POOL": "5"}])
def test_lane_per_record_kernel_knobs(eng, golden, knobs, monkeypatch):
    """The lane-per-record kernel with other segment lengths, a tiny / absent per-warp hit pool (every
    record goes through the overflow re-stream with the direct row sink) and without the 2-D tensor
    map (per-lane bulk copies) must give the same rows."""
    from oracle import np_oracle as O
    from waveformanalysis_b200.synth import make_raw_run, records_from_raw

    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    rec, pool = golden["records"], golden["wave_pool"]
    out = eng.DeviceRun.from_host(rec, pool).run_to_host(threshold=15.0)
    assert_rows_match(out["features"], golden["bf_default"], what="bf", float_exact=FX_BF)
    assert_rows_match(out["hits"], golden["hits_thr15"], what="hits", float_exact=FX_HIT)
    rpos = rec.copy()
    rpos["polarity"] = "positive"
    out = eng.DeviceRun.from_host(rpos, pool).run_to_host(features=False, threshold=-20.0)
    assert_rows_match(out["hits"], golden["hits_positive"], what="pos", float_exact=FX_HIT)
    rr, rp = golden["rag_records"], golden["rag_pool"]
    out = eng.DeviceRun.from_host(rr, rp).run_to_host(height_range=(5, -5), threshold=15.0, left_extension=3, right_extension=4)
    assert_rows_match(out["features"], golden["rag_bf"], what="rag_bf", float_exact=FX_BF)
    assert_rows_match(out["hits"], golden["rag_hits"], what="rag_hits", float_exact=("height", "width"))
    # fixed-length run from the device generator (the tensor-map layout), maximum supported extensions
    raw = make_raw_run(8, 700, 800, seed=77)
    r2, p2 = records_from_raw(raw)
    want = O.threshold_hits(r2, p2, threshold=6.0, left_extension=8, right_extension=8)
    got = eng.DeviceRun.from_host(r2, p2).run_to_host(features=False, threshold=6.0, left_extension=8, right_extension=8)
    assert_rows_match(got["hits"], want, what="ext8", float_exact=FX_HIT)
    # every extension pair the block-item variant takes (0..2), dense flicker (threshold inside the noise)
    for le, re_ in ((0, 0), (1, 2), (2, 1), (0, 2), (2, 0)):
        want = O.threshold_hits(r2, p2, threshold=4.0, left_extension=le, right_extension=re_)
        got = eng.DeviceRun.from_host(r2, p2).run_to_host(features=False, threshold=4.0, left_extension=le, right_extension=re_)
        assert_rows_match(got["hits"], want, what=f"ext {le},{re_}", float_exact=FX_HIT)


def test_full_size_properties(eng):
    """BASELINE configs[1] size (16 M records x 800 samples on one GPU): properties that do not need the
    oracle at full size - hit rows ordered by (record, start), per-record counts sum to the total,
    a slice re-processed on its own gives the same bytes, a second pass is idempotent - plus the
    oracle on a slice copied back to the host."""
    import torch

    from oracle import np_oracle as O
    from waveformanalysis_b200.dtypes import BASIC_FEATURES_DTYPE, THRESHOLD_HIT_DTYPE

    free, _ = torch.cuda.mem_get_info()
    n = 16_000_000 if free > 60e9 else 2_000_000
    L = 800
    run = eng.DeviceRun.synth(n, L, 16, seed=2025, with_rows=False)
    res = run.features_hits(threshold=15.0, hit_cap=1024, want_counts=True)
    total = int(res["total"].item())
    assert 2 * n < total < 20 * n
    counts = res["counts"]
    assert int(counts.sum().item()) == total
    out = {"features": torch.empty(n * 36, dtype=torch.uint8, device="cuda"),
           "hits": torch.empty((total + 16) * 60, dtype=torch.uint8, device="cuda"),
           "total": torch.zeros(1, dtype=torch.int64, device="cuda")}
    run.features_hits(threshold=15.0, hit_cap=total + 16, out=out)
    run.check()
    assert int(out["total"].item()) == total
    hits = out["hits"][: total * 60].view(total, 60)
    rid = hits[:, 52:60].contiguous().view(torch.int64).view(-1)
    es = hits[:, 16:20].contiguous().view(torch.int32).view(-1)
    ee = hits[:, 20:24].contiguous().view(torch.int32).view(-1)
    pos = hits[:, 0:8].contiguous().view(torch.int64).view(-1)
    key = rid * 65536 + es.to(torch.int64)
    assert bool((key[1:] >= key[:-1]).all()), "hit rows are not ordered by (record, start)"
    assert bool(((pos >= es) & (pos < ee) & (es >= 0) & (ee <= L)).all())
    assert torch.equal(torch.bincount(rid, minlength=n).to(counts.dtype), counts)
    feats = out["features"].view(n, 36)
    ev = feats[:, 28:36].contiguous().view(torch.int64).view(-1)
    assert torch.equal(ev, torch.arange(n, device="cuda"))
    # idempotence
    first_hits = out["hits"][: total * 60].clone()
    first_feat = out["features"].clone()
    run.features_hits(threshold=15.0, hit_cap=total + 16, out=out)
    assert torch.equal(out["hits"][: total * 60], first_hits) and torch.equal(out["features"], first_feat)
    # a slice in the middle, processed on its own and checked with the oracle on the host
    lo, m = n // 2 + 12345, 4096
    sub = eng.DeviceRun(run.meta[lo * 48:(lo + m) * 48].clone(), run.pool, m, 0, L)
    got = sub.run_to_host(threshold=15.0)
    f_full = first_feat.view(n, 36)[lo:lo + m].cpu().numpy().view(BASIC_FEATURES_DTYPE).reshape(-1)
    gf = got["features"].copy()
    gf["event_index"] += lo  # row index is relative to the run that was processed
    assert np.array_equal(gf.view(np.uint8), f_full.view(np.uint8))
    h0 = int(counts[:lo].sum().item())
    h1 = h0 + int(counts[lo:lo + m].sum().item())
    h_full = first_hits.view(total, 60)[h0:h1].cpu().numpy().view(THRESHOLD_HIT_DTYPE).reshape(-1)
    assert np.array_equal(got["hits"].view(np.uint8), h_full.view(np.uint8))
    pool_h = run.pool[lo * L:(lo + m) * L].cpu().numpy().view(np.uint16)
    from waveformanalysis_b200.dtypes import RECORDS_DTYPE

    rec = np.zeros(m, RECORDS_DTYPE)
    meta = run.meta[lo * 48:(lo + m) * 48].cpu().numpy()
    rec["timestamp"] = meta.view(np.int64).reshape(m, 6)[:, 0]
    rec["baseline"] = meta.view(np.float64).reshape(m, 6)[:, 1]
    rec["wave_offset"] = np.arange(m) * L
    rec["event_length"] = L
    rec["dt"] = meta.view(np.int32).reshape(m, 12)[:, 7]
    rec["board"] = meta.view(np.int16).reshape(m, 24)[:, 16]
    rec["channel"] = meta.view(np.int16).reshape(m, 24)[:, 17]
    rec["record_id"] = meta.view(np.int64).reshape(m, 6)[:, 5]
    rec["polarity"] = "unknown"
    want_f = O.basic_features(rec, pool_h)
    want_f["event_index"] += lo
    assert_rows_match(f_full, want_f, what="full-size slice features", float_exact=FX_BF)
    assert_rows_match(h_full, O.threshold_hits(rec, pool_h, threshold=15.0), what="full-size slice hits", float_exact=FX_HIT)


def test_host_pipeline_keeps_error_flags_of_early_chunks(eng, golden):
    """More chunks than pipeline slots: a record of the FIRST chunk that points outside the pool must still fail the
    call (the per-chunk flag lives in a workspace that later chunks clear)."""
    rec, pool = golden["records"].copy(), golden["wave_pool"]
    rec["wave_offset"][3] = len(pool) + 5000
    with pytest.raises(Exception, match="outside wave_pool bounds"):
        eng.process_host(rec, pool, threshold=15.0, chunk_records=64)


def test_host_pipeline_accepts_reordered_records(eng, golden):
    """RecordsView resolves every record on its own (records_view.py:47-56), so records need not be in wave_offset
    order: the host pipeline then takes the pool range of a chunk over all of its records."""
    from oracle import np_oracle as O

    rec, pool = golden["records"], golden["wave_pool"]
    rng = np.random.default_rng(3)
    perm = rng.permutation(len(rec))
    shuffled = rec[perm].copy()
    out = eng.process_host(shuffled, pool, threshold=15.0, chunk_records=100)
    assert_rows_match(out["features"], O.basic_features(shuffled, pool), what="shuffled features", float_exact=FX_BF)
    assert_rows_match(out["hits"], O.threshold_hits(shuffled, pool, threshold=15.0), what="shuffled hits", float_exact=FX_HIT)


@pytest.mark.parametrize("n_samples,variant,impl", [(800, "auto", "lane"), (250, "auto", "lane"), (1031, "auto", "lane"), (64, "auto", "lane"),
                                                    (800, "auto", "warp"), (250, "auto", "warp"), (1031, "global", "warp"), (64, "auto", "warp")])
def test_float32_pool_vs_oracle(eng, n_samples, variant, impl, monkeypatch):
    """float32 pools (wave_pool_filtered): the kernel compares raw samples with a per-record float32 bound instead of
    evaluating b - x >= thr in float64 per sample.  Samples are planted on the bound and one float32 step to either side
    of it, for all four polarity modes, fractional baselines and ragged record starts."""
    from oracle import np_oracle as O
    from waveformanalysis_b200.synth import make_raw_run, records_from_raw

    monkeypatch.setenv("WFB_FUSED_VARIANT", variant)
    monkeypatch.setenv("WFB_F32_IMPL", impl)  # lane: lane-per-record kernel (fused_f32.cuh), warp: warp-per-record kernel
    rng = np.random.default_rng(n_samples)
    raw = make_raw_run(8, 300, n_samples, seed=99 + n_samples)
    rec, pool = records_from_raw(raw)
    rec["polarity"][::5] = "negative"
    rec["polarity"][2::7] = "positive"
    rec["baseline"] += rng.uniform(-0.5, 0.5, len(rec))
    poolf = pool.astype(np.float32) + rng.normal(0, 0.3, len(pool)).astype(np.float32)
    thr = 15.0
    positive = np.asarray(rec["polarity"]) == "positive"
    for i in range(len(rec)):  # samples at the float32 neighbours of b -+ thr
        o, b = int(rec["wave_offset"][i]), float(rec["baseline"][i])
        x0 = np.float32(b + thr) if positive[i] else np.float32(b - thr)
        around = [np.nextafter(x0, np.float32(-np.inf)), x0, np.nextafter(x0, np.float32(np.inf))]
        for k, pos in enumerate(rng.choice(n_samples - 4, size=6, replace=False)):
            poolf[o + 2 + pos] = around[k % 3]
    # a ragged view: records that start off the 16-byte grid
    sub = rec[3:-3:2].copy()
    sub["wave_offset"] += 3
    sub["event_length"] -= 5
    for r_, tag in ((rec, "fixed"), (sub, "ragged")):
        want_f = O.basic_features(r_, poolf, height_range=(40, 90), area_range=(0, None))
        want_h = O.threshold_hits(r_, poolf, threshold=thr)
        out = both_paths(eng, r_, poolf, chunk_records=700, height_range=(40, 90), area_range=(0, None), threshold=thr)
        assert_rows_match(out["features"], want_f, what=f"{tag} features", float_exact=FX_BF)
        assert_rows_match(out["hits"], want_h, what=f"{tag} hits", float_exact=("height",))
        assert len(want_h) > 300
    # height range that cuts through 8-sample chunks, area range that ends inside the record
    want_f = O.basic_features(rec, poolf, height_range=(13, n_samples - 9), area_range=(5, -3))
    out = eng.process_host(rec, poolf, hits=False, height_range=(13, n_samples - 9), area_range=(5, -3))
    assert_rows_match(out["features"], want_f, what="cut ranges", float_exact=FX_BF)


@pytest.mark.parametrize("ext,knobs", [((2, 2), {}), ((0, 0), {}), ((1, 2), {}), ((2, 1), {}), ((0, 2), {}), ((2, 0), {}),
                                       ((2, 2), {"WFB_LPR_POOL": "24"}), ((2, 2), {"WFB_LPR_SC": "4"}), ((1, 1), {"WFB_LPR_SC": "16", "WFB_LPR_NO_TMAP": "1"}),
                                       ((3, 2), {})])
def test_float32_lane_kernel_edges(eng, ext, knobs, monkeypatch):
    """The register-resident run / tail state of the float32 lane-per-record kernel (fused_f32.cuh): ragged records with
    runs at both record ends (padding up to the run-wide width joins the window), runs one sample apart (overlapping
    windows), single-sample runs, every extension pair up to two samples, the pool-overflow path (rows straight out),
    other segment lengths, the per-lane bulk-copy ring.  (3, 2) is outside the lane kernel's range: the warp-per-record
    kernel takes it.  Both kernels must agree with the oracle."""
    from oracle import np_oracle as O
    from waveformanalysis_b200.dtypes import RECORDS_DTYPE

    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    rng = np.random.default_rng(sum(ext) * 7 + len(knobs))
    n, Lmax = 700, 331
    lens = rng.integers(1, Lmax + 1, n)
    lens[:40] = Lmax
    lens[40:60] = rng.integers(1, 6, 20)
    rec = np.zeros(n, dtype=RECORDS_DTYPE)
    rec["record_id"] = np.arange(n)
    rec["timestamp"] = np.arange(n) * 1_000_000
    rec["dt"] = 2
    rec["channel"] = np.arange(n) % 5
    rec["event_length"] = lens
    gaps = rng.integers(0, 7, n)  # records start anywhere relative to the 16-byte grid
    rec["wave_offset"] = np.cumsum(lens + gaps) - lens
    rec["baseline"] = 1000.0 + rng.uniform(-0.5, 0.5, n)
    pol = np.array(["unknown", "negative", "positive", "raw_positive"])[np.arange(n) % 4]
    rec["polarity"] = pol
    pool = (1000.0 + rng.normal(0, 2.0, int(rec["wave_offset"][-1] + lens[-1] + 8))).astype(np.float32)
    positive = pol == "positive"
    for i in range(n):
        o, L = int(rec["wave_offset"][i]), int(lens[i])
        sign = 1.0 if positive[i] else -1.0
        w = pool[o:o + L]
        kind = i % 6
        if kind == 0 and L > 12:      # a run that reaches the record end
            w[-rng.integers(1, 6):] += sign * 40
        elif kind == 1 and L > 12:    # a run at the record start
            w[: rng.integers(1, 6)] += sign * 40
        elif kind == 2 and L > 30:    # runs one and two samples apart, single-sample runs
            w[10:14] += sign * 40
            w[15] += sign * 40
            w[18:20] += sign * 40
            w[21] += sign * 40
        elif kind == 3 and L > 8:     # everything above threshold
            w += sign * 40
        elif kind == 4 and L > 40:    # a run ending one / two samples before the record end
            w[L - 9:L - 1 - (i % 2)] += sign * 40
    thr = 9.0
    want_h = O.threshold_hits(rec, pool, threshold=thr, left_extension=ext[0], right_extension=ext[1])
    want_f = O.basic_features(rec, pool, height_range=(3, 60), area_range=(0, None))
    assert len(want_h) > 600
    for impl in ("lane", "warp"):
        monkeypatch.setenv("WFB_F32_IMPL", impl)
        out = both_paths(eng, rec, pool, chunk_records=300, height_range=(3, 60), area_range=(0, None), threshold=thr,
                         left_extension=ext[0], right_extension=ext[1])
        assert_rows_match(out["hits"], want_h, what=f"{impl} hits ext={ext}", float_exact=("height",))
        assert_rows_match(out["features"], want_f, what=f"{impl} features", float_exact=FX_BF)
