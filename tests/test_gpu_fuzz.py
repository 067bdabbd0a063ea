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
    lens = rng.choice([0, 1, 7, 8, 9, 15, 16, 17, 31, 32, 33, 64, 100, 257], size=n).astype(np.int32)
    if rng.random() < 0.5:
        lens[:] = int(rng.choice([8, 16, 24, 40, 64, 96]))  # fixed-length pool: the tensor-map path
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
        kind = int(rng.integers(0, 7))
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


@pytest.mark.parametrize("block", range(8))
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
        want_h = O.threshold_hits(rec, pool, **kw)
        want_f = O.basic_features(rec, pool, height_range=(2, -1), area_range=(0, None))
        got = engine.DeviceRun.from_host(rec, pool).run_to_host(height_range=(2, -1), area_range=(0, None), **kw)
        assert_rows_match(got["hits"], want_h, what=f"seed {seed} hits {kw}", float_exact=FX_HIT)
        assert_rows_match(got["features"], want_f, what=f"seed {seed} features", float_exact=FX_BF)
