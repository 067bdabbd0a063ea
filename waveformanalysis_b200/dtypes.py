"""Packed numpy structured dtypes shared with the reference (layout contract of the C-ABI).

Every dtype here is byte-for-byte the one the reference defines; the CUDA kernels
read/write these packed rows directly (all fields sit on 2- or 4-byte boundaries).

Reference definitions:
  RECORDS_DTYPE / create_record_dtype  waveform_analysis/core/processing/dtypes.py:36-64, 80-100
  BASIC_FEATURES_DTYPE                 core/plugins/builtin/cpu/basic_features.py:29-40
  THRESHOLD_HIT_DTYPE                  core/plugins/builtin/cpu/hit_finder.py:33-49
  WAVEFORM_WIDTH_DTYPE                 core/plugins/builtin/cpu/waveform_width.py:22-37
  WAVEFORM_WIDTH_INTEGRAL_DTYPE        core/plugins/builtin/cpu/waveform_width_integral.py:25-39
  HIT_MERGED*_DTYPE                    core/plugins/builtin/cpu/hit_merge.py:17-49
  HIT_DTYPE                            core/plugins/builtin/cpu/peak_finding.py:30-43
"""

from __future__ import annotations

import numpy as np

RECORDS_DTYPE = np.dtype(
    [
        ("timestamp", "i8"),
        ("pid", "i4"),
        ("board", "i2"),
        ("channel", "i2"),
        ("baseline", "f8"),
        ("baseline_upstream", "f8"),
        ("polarity", "U8"),
        ("record_id", "i8"),
        ("dt", "i4"),
        ("trigger_type", "i2"),
        ("flags", "u4"),
        ("wave_offset", "i8"),
        ("event_length", "i4"),
        ("time", "i8"),
    ]
)
assert RECORDS_DTYPE.itemsize == 102


def create_record_dtype(wave_length: int) -> np.dtype:
    """ST_WAVEFORM_DTYPE with a ``wave`` field of ``wave_length`` int16 samples (76 + 2L bytes)."""
    return np.dtype(
        [
            ("baseline", "f8"),
            ("baseline_upstream", "f8"),
            ("polarity", "U8"),
            ("timestamp", "i8"),
            ("record_id", "i8"),
            ("dt", "i4"),
            ("event_length", "i4"),
            ("board", "i2"),
            ("channel", "i2"),
            ("wave", "i2", (int(wave_length),)),
        ]
    )


BASIC_FEATURES_DTYPE = np.dtype(
    [
        ("height", "f4"),
        ("amp", "f4"),
        ("area", "f4"),
        ("max_abs_diff", "f4"),
        ("timestamp", "i8"),
        ("board", "i2"),
        ("channel", "i2"),
        ("event_index", "i8"),
    ]
)
assert BASIC_FEATURES_DTYPE.itemsize == 36

THRESHOLD_HIT_DTYPE = np.dtype(
    [
        ("position", "i8"),
        ("height", "f4"),
        ("integral", "f4"),
        ("edge_start", "i4"),
        ("edge_end", "i4"),
        ("width", "f4"),
        ("dt", "i4"),
        ("rise_time", "f4"),
        ("fall_time", "f4"),
        ("timestamp", "i8"),
        ("board", "i2"),
        ("channel", "i2"),
        ("record_id", "i8"),
    ]
)
assert THRESHOLD_HIT_DTYPE.itemsize == 60

HIT_DTYPE = np.dtype(
    [
        ("position", "i8"),
        ("height", "f4"),
        ("integral", "f4"),
        ("edge_start", "f4"),
        ("edge_end", "f4"),
        ("dt", "i4"),
        ("timestamp", "i8"),
        ("board", "i2"),
        ("channel", "i2"),
        ("record_id", "i8"),
    ]
)

WAVEFORM_WIDTH_DTYPE = np.dtype(
    [
        ("rise_time", "f4"),
        ("fall_time", "f4"),
        ("total_width", "f4"),
        ("rise_time_samples", "f4"),
        ("fall_time_samples", "f4"),
        ("total_width_samples", "f4"),
        ("peak_position", "i8"),
        ("peak_height", "f4"),
        ("timestamp", "i8"),
        ("board", "i2"),
        ("channel", "i2"),
        ("record_id", "i8"),
    ]
)
assert WAVEFORM_WIDTH_DTYPE.itemsize == 56

WAVEFORM_WIDTH_INTEGRAL_DTYPE = np.dtype(
    [
        ("t_low", "f4"),
        ("t_high", "f4"),
        ("width", "f4"),
        ("t_low_samples", "f4"),
        ("t_high_samples", "f4"),
        ("width_samples", "f4"),
        ("q_total", "f8"),
        ("timestamp", "i8"),
        ("board", "i2"),
        ("channel", "i2"),
        ("event_index", "i8"),
    ]
)
assert WAVEFORM_WIDTH_INTEGRAL_DTYPE.itemsize == 52

HIT_MERGED_DTYPE = np.dtype(
    [
        ("position", "i8"),
        ("height", "f4"),
        ("integral", "f4"),
        ("sample_start", "i4"),
        ("sample_end", "i4"),
        ("width", "f4"),
        ("dt", "i4"),
        ("rise_time", "f4"),
        ("fall_time", "f4"),
        ("timestamp", "i8"),
        ("board", "i2"),
        ("channel", "i2"),
        ("record_id", "i8"),
        ("component_offset", "i8"),
        ("component_count", "i4"),
    ]
)
assert HIT_MERGED_DTYPE.itemsize == 72

HIT_MERGED_COMPONENTS_DTYPE = np.dtype([("merged_index", "i8"), ("hit_index", "i8")])
HIT_MERGE_CLUSTERS_DTYPE = np.dtype([("cluster_index", "i8"), ("hit_index", "i8")])

# polarity codes used in the device-side record metadata (wfb_rec_meta.polarity)
POLARITY_UNKNOWN = 0
POLARITY_POSITIVE = 1
POLARITY_NEGATIVE = 2


def polarity_codes(polarity: np.ndarray) -> np.ndarray:
    """Map the reference's 'positive' | 'negative' | anything-else strings to device codes."""
    pol = np.asarray(polarity)
    out = np.zeros(pol.shape, dtype=np.uint8)
    out[pol == "positive"] = POLARITY_POSITIVE
    out[pol == "negative"] = POLARITY_NEGATIVE
    return out


# core/plugins/builtin/cpu/s1_s2_classifier.py:29-42 (packed, 45 bytes)
S1_S2_CLASSIFIER_DTYPE = np.dtype(
    [
        ("label", "i1"),
        ("width_ns", "f4"),
        ("width_samples", "f4"),
        ("height", "f4"),
        ("area", "f4"),
        ("timestamp", "i8"),
        ("board", "i2"),
        ("channel", "i2"),
        ("record_id", "i8"),
        ("peak_position", "i8"),
    ]
)
LABEL_UNKNOWN, LABEL_S1, LABEL_S2 = 0, 1, 2
