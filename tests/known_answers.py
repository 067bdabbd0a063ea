"""Known-answer vectors taken from the reference's OWN tests (hand-computed values there).

Each case: inputs rebuilt exactly as the cited reference test builds them, and the values the
reference test asserts.  Used by tests/test_oracle_known_answers.py (CPU, numpy oracle) and by
tests/test_gpu_known_answers.py (CUDA path through the C-ABI).
"""

import numpy as np

from waveformanalysis_b200.dtypes import RECORDS_DTYPE


def _records(n):
    r = np.zeros(n, dtype=RECORDS_DTYPE)
    r["polarity"] = ""  # np.zeros default in the reference tests (not 'unknown')
    r["baseline_upstream"] = 0.0
    return r


def bf_records_view():
    """tests/basic_features_helpers.py:48-60 + tests/test_basic_features_records.py:28-42."""
    r = _records(2)
    r["timestamp"] = [10, 20]
    r["board"] = [3, 4]
    r["channel"] = [0, 1]
    r["record_id"] = [10, 11]
    r["baseline"] = [100.0, 100.0]
    r["wave_offset"] = [0, 4]
    r["event_length"] = [4, 4]
    pool = np.array([90, 80, 95, 90, 90, 85, 90, 90], dtype=np.uint16)
    cfg = dict(height_range=(0, 4), area_range=(0, 4))
    want = dict(height0=20.0, amp0=15.0, max_abs_diff0=15.0, boards=[3, 4])
    return r, pool, cfg, want


def bf_fixed_baseline():
    """tests/test_basic_features_records.py:44-61: fixed_baseline 95 on channel 3:0 ->
    height 0, area -25."""
    r, _, cfg, _ = bf_records_view()
    pool = np.array([95, 100, 115, 95, 90, 85, 90, 90], dtype=np.uint16)
    return r, pool, dict(cfg, fixed_baseline={(3, 0): 95.0}), dict(height0=0.0, area0=-25.0)


def bf_filtered_pool():
    """tests/basic_features_helpers.py:63-77 + tests/test_basic_features_records.py:72-90."""
    r, _, cfg, _ = bf_records_view()
    pool = np.full(8, 95, dtype=np.float32)
    return r, pool, cfg, dict(height=[5.0, 5.0], amp=[0.0, 0.0], area=[20.0, 20.0], max_abs_diff=[0.0, 0.0])


def _hit_record(wave, *, board=0, channel=0, ts=1_000_000, dt=2, event_length=None):
    r = _records(1)
    r["baseline"] = 100.0
    r["timestamp"] = ts
    r["board"] = board
    r["channel"] = channel
    r["dt"] = dt
    r["event_length"] = len(wave) if event_length is None else event_length
    r["wave_offset"] = 0
    return r, np.asarray(wave, dtype=np.uint16)


def hits_two_regions():
    """tests/plugins/test_threshold_hit_plugin.py:99-123: edges [5,15]/[8,18], width 3."""
    w = np.full(32, 100)
    w[5:8] = 80
    w[15:18] = 70
    r, p = _hit_record(w)
    return r, p, dict(threshold=10.0, left_extension=0, right_extension=0), dict(
        edge_start=[5, 15], edge_end=[8, 18], width=[3.0, 3.0], dt=[2, 2])


def hits_extension():
    """tests/plugins/test_threshold_hit_plugin.py:158-178: edges 8..15, width 7."""
    w = np.full(32, 100)
    w[10:12] = 80
    r, p = _hit_record(w)
    return r, p, dict(threshold=10.0, left_extension=2, right_extension=3), dict(
        edge_start=[8], edge_end=[15], width=[7.0])


def hits_records_view():
    """tests/plugins/test_threshold_hit_plugin.py:31-42, 210-233: board 5, channel 2, edges [2,6)."""
    r, p = _hit_record([100, 100, 80, 80, 80, 80, 100, 100], board=5, channel=2, ts=123_456)
    return r, p, dict(threshold=10.0, left_extension=0, right_extension=0), dict(
        edge_start=[2], edge_end=[6], board=[5], channel=[2], record_id=[0])


def hits_rise_fall(ext=False):
    """tests/plugins/test_threshold_hit_plugin.py:352-395: rise = fall = 4 ns with and without
    extensions; with (2,3) extensions edges are [2,12)."""
    w = np.full(16, 100)
    w[4:9] = [80, 70, 60, 70, 80]
    r, p = _hit_record(w)
    if ext:
        return r, p, dict(threshold=10.0, left_extension=2, right_extension=3), dict(
            edge_start=[2], edge_end=[12], rise_time=[4.0], fall_time=[4.0])
    return r, p, dict(threshold=10.0, left_extension=0, right_extension=0), dict(
        rise_time=[4.0], fall_time=[4.0])


def dual_baseline_case():
    """tests/test_dual_baseline.py:60-95: baseline ~= 100 from the first 40 samples."""
    rng = np.random.default_rng(0)
    samples = (100 + rng.integers(-1, 2, size=(6, 120))).astype(np.int16)
    return samples, samples[:, :40].astype(np.float64).mean(axis=1)


def grouping_case():
    """tests/plugins/test_hit_grouped_plugin.py:76-127: two hits on channels 0/1 whose absolute
    windows overlap (time_window_ns=0) -> one event, t_min=96000, t_max=110000, dt/ns=14."""
    from waveformanalysis_b200.dtypes import THRESHOLD_HIT_DTYPE

    h = np.zeros(2, dtype=THRESHOLD_HIT_DTYPE)
    h["position"] = [10, 14]
    h["height"] = [20.0, 25.0]
    h["integral"] = [30.0, 40.0]
    h["edge_start"] = [8, 13]
    h["edge_end"] = [12, 16]
    h["timestamp"] = [100_000, 106_000]
    h["dt"] = 2
    h["board"] = 0
    h["channel"] = [0, 1]
    h["record_id"] = [0, 1]
    return h, 0.0, dict(t_min=[96_000], t_max=[110_000], dt_ns=[14.0], n_hits=[2])
