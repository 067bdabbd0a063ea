"""CPU: numpy oracle against the reference tests' hand-computed values (tests/known_answers.py)."""

import numpy as np

import known_answers as K
from oracle import np_oracle as O


def test_bf_records_view():
    r, pool, cfg, want = K.bf_records_view()
    out = O.basic_features(r, pool, **cfg)
    assert np.isclose(out["height"][0], want["height0"])
    assert np.isclose(out["amp"][0], want["amp0"])
    assert np.isclose(out["max_abs_diff"][0], want["max_abs_diff0"])
    assert out["board"].tolist() == want["boards"]


def test_bf_fixed_baseline():
    r, pool, cfg, want = K.bf_fixed_baseline()
    out = O.basic_features(r, pool, **cfg)
    assert np.isclose(out["height"][0], want["height0"])
    assert np.isclose(out["area"][0], want["area0"])


def test_bf_filtered_pool():
    r, pool, cfg, want = K.bf_filtered_pool()
    out = O.basic_features(r, pool, **cfg)
    for k, v in want.items():
        np.testing.assert_allclose(out[k], v)


def _check_hits(case):
    r, pool, cfg, want = case
    out = O.threshold_hits(r, pool, **cfg)
    for k, v in want.items():
        np.testing.assert_array_equal(out[k], np.asarray(v, dtype=out[k].dtype))


def test_hits_known_answers():
    _check_hits(K.hits_two_regions())
    _check_hits(K.hits_extension())
    _check_hits(K.hits_records_view())
    _check_hits(K.hits_rise_fall(False))
    _check_hits(K.hits_rise_fall(True))


def test_dual_baseline():
    samples, want = K.dual_baseline_case()
    got = O.baseline_mean(samples, 0, 40)
    assert np.array_equal(got, want)
    assert np.all(np.abs(got - 100) < 1)
    assert np.isnan(O.baseline_mean(samples, 5, 5)).all()


def test_grouping_known_answer():
    h, w, want = K.grouping_case()
    ev = O.group_hit_windows(h, w)
    assert ev["t_min"].tolist() == want["t_min"]
    assert ev["t_max"].tolist() == want["t_max"]
    assert ev["dt_ns"].tolist() == want["dt_ns"]
    assert ev["n_hits"].tolist() == want["n_hits"]


def test_find_peaks_restatement_matches_live_scipy():
    """The oracle's find_peaks restatement against scipy.signal.find_peaks itself (the un-vendored
    dependency that holds the arithmetic of the reference's `hit` plugin): random walks with plateaus
    and ties, every condition the plugin passes."""
    from scipy.signal import find_peaks

    from oracle import np_oracle as O

    rng = np.random.default_rng(3)
    for trial in range(60):
        n = int(rng.integers(3, 400))
        x = np.cumsum(rng.normal(0, 3, n))
        if trial % 3 == 0:
            x = np.round(x)  # plateaus and equal heights
        kw = dict(height=float(rng.uniform(-5, 5)), prominence=float(rng.uniform(0.2, 6)), width=float(rng.uniform(0.5, 6)),
                  distance=int(rng.integers(1, 12)) if trial % 2 else 2)
        if trial % 5 == 0:
            kw["threshold"] = float(rng.uniform(0, 1.5))
        if trial % 3 == 0 and kw["distance"] > 2:
            kw["distance"] = 2  # equal heights: scipy's priority order among ties is not defined
        want_p, props = find_peaks(x, **kw)
        got_p, lips, rips, proms = O.find_peaks_1d(x, **kw)
        assert np.array_equal(got_p, want_p), (trial, kw)
        assert np.array_equal(lips, props["left_ips"]) and np.array_equal(rips, props["right_ips"]), trial
        assert np.array_equal(proms, props["prominences"]), trial
