"""Shared cases of the df / df_paired / s1_s2 tests (the configs tests/golden/make_golden_after.py ran through
the live reference)."""

import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
AFTER = os.path.join(HERE, "golden", "after_golden.npz")

GAINS_CONFIG = {"0:0": 12.5, "0:1": 13.2, "0:2": {"gain_adc_per_pe": 7.0}, "0:3": -1.0}
# what resolve_channel_value_map keeps of it (channel.py:595-611): mappings are unwrapped, gain <= 0 dropped
GAINS_RESOLVED = {(0, 0): 12.5, (0, 1): 13.2, (0, 2): 7.0}

S1S2_CASES = {
    "none": {},
    "width": {"s1_width_range": (None, 420.0), "s2_width_range": (420.0, None)},
    "samples_conflict": {"width_unit": "samples", "s1_width_range": (75.0, 230.0), "s2_width_range": (200.0, 600.0),
                         "s1_height_range": (6.0, None), "conflict_policy": "prefer_s2"},
    "area_prefer_s1": {"s1_area_range": (0.0, 25000.0), "s2_area_range": (15000.0, None), "s2_height_range": (None, 7.5),
                       "conflict_policy": "prefer_s1"},
    "conflict_unknown": {"s1_width_range": (0.0, 500.0), "s2_width_range": (400.0, 1200.0)},
    "only_s2_height": {"s2_height_range": (5.0, 8.0)},
}

DF_COLUMNS = ("timestamp", "record_id", "area", "height", "amp", "max_abs_diff", "board", "channel")


def load_after():
    return np.load(AFTER, allow_pickle=False)


def check_df_columns(cols: dict, A, prefix: str, with_pe: bool):
    assert np.array_equal(cols["order"], A[f"{prefix}_index"]), prefix
    names = DF_COLUMNS + (("area_pe", "height_pe") if with_pe else ())
    for c in names:
        want = A[f"{prefix}_{c}"]
        got = np.asarray(cols[c])
        assert got.dtype == want.dtype, (prefix, c, got.dtype, want.dtype)
        assert np.array_equal(got, want, equal_nan=want.dtype.kind == "f"), (prefix, c)


def check_pair(out: dict, offsets, A, name: str):
    """out = keep / delta_t / area_ch / height_ch for ALL events; golden = the kept rows of the reference."""
    nch, start = (int(v) for v in A[f"pair_{name}_nch_start"])
    keep = out["keep"]
    assert np.array_equal(np.flatnonzero(keep), A[f"pair_{name}_index"]), name
    assert np.array_equal(out["delta_t"][keep], A[f"pair_{name}_delta_t"]), name
    for i in range(nch):
        for kind, arr in (("area", out["area_ch"]), ("height", out["height_ch"])):
            want = A[f"pair_{name}_{kind}_ch{start + i}"]
            assert np.array_equal(arr[keep, i].astype(want.dtype), want, equal_nan=True), (name, kind, i)
