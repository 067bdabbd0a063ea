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
