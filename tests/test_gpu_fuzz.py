"""Randomised parity of the fused kernel's hit path against the numpy oracle on adversarial small runs:
runs that start / end exactly on 8-sample chunk boundaries, records that are entirely above threshold,
alternating samples, values exactly at the threshold, misaligned ragged records, every extension pair,
mixed polarities, negative / zero / unreachable thresholds, tiny hit pools and other segment lengths."""

import numpy as np
import pytest

from conftest import assert_rows_match

pytestmark = pytest.mark.gpu

FX_HIT = ("height", "width", "rise_time", "fall_time")
FX_BF = ("height", "amp", "max_abs_diff")


def adversarial_run(seed: int):
    from waveformanalysis_b200.dtypes import RECORDS_DTYPE

    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 90))
    lens = rng.choice([0, 1, 7, 8, 9, 15, 16, 17, 31, 32, 33, 64, 65, 100, 160, 257, 513], size=n).astype(np.int32)
    if rng.random() < 0.5:
        lens[:] = int(rng.choice([8, 16, 24, 40, 64, 96, 128, 320]))  # fixed-length pool: the tensor-map path
    offs = np.zeros(n, dtype=np.int64)
    gap = int(rng.integers(0, 4)) if lens.min() != lens.max() else 0
    offs[1:] = np.cumsum(lens[:-1] + gap)
    total = int(offs[-1] + lens[-1]) + gap
    base = 1000 if rng.random() < 0.8 else int(rng.integers(0, 60000))
    pool = np.full(total, base, dtype=np.int64)
    thr = float(rng.choice([0.0, 1.0, 3.0, 15.0, -2.0, 1e9, 2.5]))
    amp = int(max(abs(thr), 1)) + int(rng.integers(0, 4))
    rec = np.zeros(n, dtype=RECORDS_DTYPE)
    for i in range(n):
        L, o = int(lens[i]), int(offs[i])
        if L == 0:
            continue
        kind = int(rng.integers(0, 9))
        w = np.full(L, base, dtype=np.int64)
        sign = -1 if rng.random() < 0.6 else 1
        if kind == 0:      # everything above threshold
            w += sign * amp * 3
        elif kind == 1:    # alternating samples
            w[::2] += sign * amp * 2
        elif kind == 2:    # runs starting / ending on chunk boundaries (in pool coordinates)
            for _ in range(int(rng.integers(1, 4))):
                a = int(rng.integers(0, L))
                a -= (o + a) % 8
                b = a + 8 * int(rng.integers(1, 5))
                w[max(a, 0):min(max(b, 0), L)] += sign * amp * 2
        elif kind == 3:    # values exactly at / just below the threshold
            w += sign * rng.integers(0, 2, size=L) * int(round(abs(thr)))
        elif kind == 4:    # random walk
            w += np.cumsum(rng.integers(-3, 4, size=L))
        elif kind == 7:    # runs starting / ending on or next to 32-sample block boundaries (in pool coordinates)
            for _ in range(int(rng.integers(1, 4))):
                a = int(rng.integers(0, L))
                a -= (o + a) % 32
                b = a + 32 * int(rng.integers(1, 4))
                a += int(rng.integers(-2, 3))
                b += int(rng.integers(-2, 3))
                w[max(a, 0):min(max(b, 0), L)] += sign * amp * 2
        elif kind == 8:    # a long run with a distinct minimum somewhere inside, noise on top
            a, b = sorted(int(x) for x in rng.integers(0, L + 1, size=2))
            w[a:b] += sign * amp * 3
            if b > a:
                w[int(rng.integers(a, b))] += sign * amp * 5
            w += sign * rng.integers(0, 2, size=L)
        elif kind == 5:    # one long pulse with a flickering tail
            a = int(rng.integers(0, L))
            w[a:] += sign * (amp * 4 * np.exp(-np.arange(L - a) / max(L / 6, 1))).astype(np.int64)
            w += rng.integers(-2, 3, size=L)
        pool[o:o + L] = w
    pool = np.clip(pool, 0, 65535).astype(np.uint16)
    rec["wave_offset"] = offs
    rec["event_length"] = lens
    rec["timestamp"] = np.cumsum(rng.integers(1, 10**6, size=n)) * 1000
    rec["dt"] = rng.choice([1, 2, 4], size=n)
    rec["board"] = rng.integers(0, 2, size=n)
    rec["channel"] = rng.integers(0, 3, size=n)
    rec["record_id"] = np.arange(n)
    rec["baseline"] = base + rng.choice([0.0, 0.25, -0.5, 0.999], size=n)
    rec["baseline_upstream"] = np.nan
    rec["polarity"] = rng.choice(["unknown", "positive", "negative"], size=n)
    kw = dict(threshold=thr, left_extension=int(rng.integers(0, 9)), right_extension=int(rng.integers(0, 9)))
    return rec, pool, kw


@pytest.mark.parametrize("block", range(12))
def test_fused_hits_adversarial(block, monkeypatch):
    from oracle import np_oracle as O
    from waveformanalysis_b200 import engine

    if block % 4 == 1:
        monkeypatch.setenv("WFB_LPR_POOL", "3")
    if block % 4 == 2:
        monkeypatch.setenv("WFB_LPR_SC", "4")
    if block % 4 == 3:
        monkeypatch.setenv("WFB_LPR_NO_TMAP", "1")
    for seed in range(block * 40, block * 40 + 40):
        rec, pool, kw = adversarial_run(seed)
        if block >= 4:  # extensions of at most two samples: the block-item variant of the lane-per-record kernel
            kw["left_extension"] %= 3
            kw["right_extension"] %= 3
        if block >= 10:
            monkeypatch.setenv("WFB_LPR_IMPL", "chunk")
        want_h = O.threshold_hits(rec, pool, **kw)
        want_f = O.basic_features(rec, pool, height_range=(2, -1), area_range=(0, None))
        got = engine.DeviceRun.from_host(rec, pool).run_to_host(height_range=(2, -1), area_range=(0, None), **kw)
        assert_rows_match(got["hits"], want_h, what=f"seed {seed} hits {kw}", float_exact=FX_HIT)
        assert_rows_match(got["features"], want_f, what=f"seed {seed} features", float_exact=FX_BF)


def test_find_peaks_adversarial():
    """`hit` on plateaus, monotone ramps, peaks at the record edges, equal neighbours and tiny records, all
    three sources of the plugin, against the oracle's find_peaks restatement (itself pinned to live scipy)."""
    from oracle import np_oracle as O
    from waveformanalysis_b200 import ops
    from waveformanalysis_b200.dtypes import RECORDS_DTYPE, create_record_dtype

    rng = np.random.default_rng(123)
    fx = ("height", "edge_start", "edge_end")
    for trial in range(12):
        n, L = int(rng.integers(1, 40)), int(rng.choice([3, 4, 8, 17, 64, 200]))
        w = np.full((n, L), 2000, dtype=np.int64)
        for i in range(n):
            kind = int(rng.integers(0, 5))
            if kind == 0:
                w[i] += np.cumsum(rng.integers(-4, 5, size=L))
            elif kind == 1:
                w[i] -= (np.arange(L) // 3) * 7          # descending staircase: plateaus in the derivative
            elif kind == 2:
                w[i, : L // 2] -= 300                      # step at the start / middle
            elif kind == 3:
                w[i] -= (100 * np.exp(-((np.arange(L) - L / 2) ** 2) / max(L / 3, 1))).astype(np.int64)
            w[i] += rng.integers(-1, 2, size=L)
        w = np.clip(w, 0, 16383)
        st = np.zeros(n, dtype=create_record_dtype(L))
        st["wave"] = w.astype(np.int16)
        st["timestamp"] = np.arange(n) * 10**6
        st["dt"] = 2
        st["channel"] = rng.integers(0, 3, size=n)
        st["record_id"] = np.arange(n)
        st["event_length"] = L
        st["baseline"] = 2000.25
        kw = dict(height=float(rng.choice([1.0, 3.0, 10.0])), prominence=float(rng.choice([0.5, 2.0])), width=float(rng.choice([1, 2])),
                  use_derivative=bool(trial % 2), distance=2)
        want = O.hit_find_peaks(list(st["wave"]), st, source="aos", **kw)
        assert_rows_match(ops.find_peaks_waveforms(st, **kw), want, what=f"aos trial {trial}", float_exact=fx)
        rec = np.zeros(n, dtype=RECORDS_DTYPE)
        for f in ("timestamp", "dt", "channel", "record_id", "event_length", "baseline"):
            rec[f] = st[f]
        rec["wave_offset"] = np.arange(n) * L
        rec["polarity"] = rng.choice(["unknown", "positive"], size=n)
        pool = w.astype(np.uint16).reshape(-1)
        sig = pool.reshape(n, L).astype(np.float32) - rec["baseline"].astype(np.float32)[:, None]
        waves = [(s if p == "positive" else -s).astype(np.float64) for s, p in zip(sig, rec["polarity"])]
        want = O.hit_find_peaks(waves, rec, source="records", **kw)
        assert_rows_match(ops.find_peaks_records(rec, pool, **kw), want, what=f"records trial {trial}", float_exact=fx)


def test_fused_hits_signed_rows_adversarial():
    """int16 structured rows used in place (offset-binary path of the kernel) with negative sample values."""
    from oracle import np_oracle as O
    from waveformanalysis_b200 import engine
    from waveformanalysis_b200.aos import structured_as_records
    from waveformanalysis_b200.dtypes import create_record_dtype

    rng = np.random.default_rng(77)
    for trial in range(10):
        n, L = int(rng.integers(1, 70)), int(rng.choice([8, 16, 40, 96]))
        st = np.zeros(n, dtype=create_record_dtype(L))
        base = int(rng.choice([-3000, 0, 12, 9000]))
        w = base + np.cumsum(rng.integers(-6, 7, size=(n, L)), axis=1)
        w[:, L // 3: L // 3 + int(rng.integers(1, L // 2))] -= int(rng.integers(5, 60))
        st["wave"] = np.clip(w, -32768, 32767).astype(np.int16)
        st["timestamp"] = np.arange(n) * 10**6
        st["dt"] = 4
        st["channel"] = rng.integers(0, 4, size=n)
        st["record_id"] = np.arange(n)
        st["event_length"] = L
        st["baseline"] = base + rng.choice([0.0, 0.5, -0.125], size=n)
        st["polarity"] = "unknown"
        rec, pool, signed = structured_as_records(st)
        assert signed
        kw = dict(threshold=float(rng.choice([2.0, 7.0, 20.0])), left_extension=int(rng.integers(0, 9)), right_extension=int(rng.integers(0, 9)))
        # oracle on the true (signed) sample values: shift samples and baseline into the uint16 range
        shift = 32768
        rec_u = np.zeros(n, dtype=rec.dtype)
        for f in rec.dtype.names:
            rec_u[f] = rec[f]
        rec_u["wave_offset"] = np.arange(n) * L
        rec_u["baseline"] = rec["baseline"] + shift
        pool_u = (st["wave"].astype(np.int64) + shift).astype(np.uint16).reshape(-1)
        want = O.threshold_hits(rec_u, pool_u, **kw)
        want["height"] = want["height"]  # heights are differences: the shift cancels
        got = engine.process_host(rec, pool, features=False, signed_samples=True, **kw)["hits"]
        assert_rows_match(got, want, what=f"signed trial {trial}", float_exact=("width", "rise_time", "fall_time"))


def test_after_path_ops_fuzz():
    """df_columns / s1s2_classify / pair_events against the oracle on seeded random rows (ties in the
    timestamps included: both sides sort stably)."""
    from oracle import np_oracle as O
    from waveformanalysis_b200 import ops
    from waveformanalysis_b200.dtypes import BASIC_FEATURES_DTYPE, WAVEFORM_WIDTH_DTYPE

    rng = np.random.default_rng(99)
    for n in (1, 257, 5000, 70001):
        bf = np.zeros(n, dtype=BASIC_FEATURES_DTYPE)
        bf["timestamp"] = rng.integers(-5, max(n // 3, 2), n) * 1000 - (1 << 40) * (rng.random(n) < 0.1)
        for f in ("height", "amp", "area", "max_abs_diff"):
            bf[f] = rng.normal(0, 1e3, n).astype(np.float32)
        bf["area"][rng.random(n) < 0.02] = np.nan
        bf["board"], bf["channel"] = rng.integers(0, 3, n), rng.integers(0, 5, n)
        rid = rng.permutation(n).astype(np.int64)
        gains = {(0, 0): 3.5, (1, 4): 0.25, (2, 2): 1e-3, (7, 7): 2.0}
        for g, r in ((None, None), (gains, rid), ({}, rid)):
            got, want = ops.df_columns(bf, r, g), O.df_columns(bf, r, g)
            assert set(got) == set(want)
            for k in want:
                assert got[k].dtype == want[k].dtype, k
                assert np.array_equal(got[k], want[k], equal_nan=want[k].dtype.kind == "f"), (n, k)
        # s1_s2
        m = max(n // 2, 1)
        ww = np.zeros(m, dtype=WAVEFORM_WIDTH_DTYPE)
        ww["total_width"] = rng.uniform(0, 500, m).astype(np.float32)
        ww["total_width"][rng.random(m) < 0.05] = np.nan
        ww["total_width_samples"] = ww["total_width"] / 2
        ww["record_id"] = rng.integers(-3, n + 3, m)
        ww["timestamp"], ww["peak_position"] = rng.integers(0, 1 << 50, m), rng.integers(0, 800, m)
        ww["board"], ww["channel"] = rng.integers(0, 3, m), rng.integers(0, 5, m)
        conf = dict(s1_width_range=(None, 250.0), s2_width_range=(200.0, None), s1_area_range=(-500.0, 500.0), s2_height_range=(0.0, None),
                    conflict_policy=("unknown", "prefer_s1", "prefer_s2")[n % 3], width_unit=("ns", "samples")[n % 2])
        got, want = ops.s1s2_classify(ww, bf, **conf), O.s1s2_classify(ww, bf, **conf)
        assert got.tobytes() == want.tobytes() or all(np.array_equal(got[f], want[f], equal_nan=True) for f in want.dtype.names)
        # df_paired
        lens = rng.integers(1, 6, m)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        tot = int(off[-1])
        ts = rng.integers(0, 1 << 45, tot)
        ar, he = rng.normal(0, 100, tot).astype(np.float32), rng.normal(0, 10, tot).astype(np.float32)
        dt = rng.uniform(0, 200, m)
        dt[rng.random(m) < 0.05] = np.nan
        for nch in (0, 1, 4):
            got, want = ops.pair_events(off, ts, ar, he, dt, 100.0, nch), O.pair_events(off, ts, ar, he, dt, 100.0, nch)
            for k in want:
                assert np.array_equal(got[k], want[k], equal_nan=True), (n, nch, k)


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_float32_lane_and_warp_kernels_agree_on_random_pools(seed, monkeypatch):
    """Differential fuzz of the two float32 kernels (fused_f32.cuh lane-per-record / fused_features_hits.cu warp-per-record)
    and the oracle: random lengths, offsets, polarities, thresholds (per channel too), extensions, samples that sit on
    the threshold bound, NaN / inf samples inside and outside runs."""
    from oracle import np_oracle as O
    from waveformanalysis_b200 import engine
    from waveformanalysis_b200.dtypes import RECORDS_DTYPE

    rng = np.random.default_rng(1000 + seed)
    n = 1500
    Lmax = int(rng.choice([37, 128, 515, 800]))
    lens = rng.integers(1, Lmax + 1, n)
    lens[rng.random(n) < 0.5] = Lmax
    rec = np.zeros(n, dtype=RECORDS_DTYPE)
    rec["record_id"] = np.arange(n)
    rec["timestamp"] = np.cumsum(rng.integers(1, 5000, n)) * 1000
    rec["dt"] = rng.choice([1, 2, 4], n)
    rec["board"] = rng.integers(0, 2, n)
    rec["channel"] = rng.integers(0, 4, n)
    rec["event_length"] = lens
    rec["wave_offset"] = np.cumsum(lens + rng.integers(0, 5, n)) - lens
    rec["baseline"] = rng.choice([0.0, 12.5, 1000.0, -40.25, 16000.7], n) + rng.uniform(-1, 1, n)
    rec["polarity"] = rng.choice(["unknown", "negative", "positive"], n)
    total = int(rec["wave_offset"][-1] + lens[-1] + 8)
    pool = np.zeros(total, dtype=np.float32)
    thr = float(rng.choice([3.0, 9.5, 25.0]))
    positive = np.asarray(rec["polarity"]) == "positive"
    for i in range(n):
        o, L, b = int(rec["wave_offset"][i]), int(lens[i]), float(rec["baseline"][i])
        w = (b + rng.normal(0, thr / 3.0, L)).astype(np.float32)
        sign = 1.0 if positive[i] else -1.0
        for _ in range(int(rng.integers(0, 4))):
            a0 = int(rng.integers(0, L))
            a1 = min(L, a0 + int(rng.integers(1, 40)))
            w[a0:a1] += np.float32(sign * thr * rng.uniform(1.0, 6.0))
        x0 = np.float32(b + sign * thr)
        for k in rng.integers(0, L, size=3):  # on the bound and its float32 neighbours
            w[k] = [np.nextafter(x0, np.float32(-np.inf)), x0, np.nextafter(x0, np.float32(np.inf))][int(k) % 3]
        if i % 37 == 0 and L > 4:
            w[int(rng.integers(0, L))] = np.float32(sign * np.inf)  # a sample far above threshold
        if i % 41 == 0 and L > 4:
            w[int(rng.integers(0, L))] = np.float32(-sign * np.inf)  # far below: never part of a run
        pool[o:o + L] = w
    ext = (int(rng.integers(0, 3)), int(rng.integers(0, 3)))
    thrs = {(0, 1): thr * 2.0, (1, 3): thr * 0.5}
    want_h = O.threshold_hits(rec, pool, threshold=thr, thresholds=thrs, left_extension=ext[0], right_extension=ext[1])
    want_f = O.basic_features(rec, pool, height_range=(2, 30), area_range=(1, -1))
    outs = {}
    for impl in ("lane", "warp"):
        monkeypatch.setenv("WFB_F32_IMPL", impl)
        outs[impl] = engine.process_host(rec, pool, threshold=thr, thresholds=thrs, left_extension=ext[0], right_extension=ext[1],
                                         height_range=(2, 30), area_range=(1, -1), chunk_records=400)
        finite = np.isfinite(want_h["height"]) & np.isfinite(want_h["integral"])
        for f in want_h.dtype.names:
            g, w_ = outs[impl]["hits"][f], want_h[f]
            if w_.dtype.kind in "iu":
                assert np.array_equal(g, w_), (impl, f)
            else:
                assert np.allclose(g[finite], w_[finite], rtol=1e-5, atol=1e-3), (impl, f)
        ok = np.isfinite(want_f["area"]) & np.isfinite(want_f["height"]) & np.isfinite(want_f["max_abs_diff"])
        for f in ("height", "amp", "area", "max_abs_diff"):
            assert np.allclose(outs[impl]["features"][f][ok], want_f[f][ok], rtol=1e-5, atol=1e-3), (impl, f)
    assert len(want_h) > 1000
