"""Structured waveform arrays (``st_waveforms`` / ``filtered_waveforms``, AoS rows of
``create_record_dtype(L)``: 76 header bytes + L samples, core/processing/dtypes.py:36-64) viewed
as records + a sample pool WITHOUT repacking: the AoS buffer itself is the pool, every row's
``wave`` field is a (misaligned) record inside it."""

from __future__ import annotations

import numpy as np

from .dtypes import RECORDS_DTYPE

RAW_POSITIVE = "rawpos"  # internal polarity tag: raw (unsigned-signal) arithmetic with positive pulses


def structured_as_records(data: np.ndarray, *, explicit_dt: int | None = None, raw_polarity: bool = False,
                          return_clamp: bool = False):
    """Return (records, pool, signed) for a structured waveform array; with ``return_clamp`` a fourth value: None, or
    the int32 lengths the edges of hit rows are clamped to when ``event_length`` differs from the row width.

    ``raw_polarity``: basic_features' st branch never uses the float32 signal path; a 'positive'
    polarity selects max-baseline / sum(wave-baseline) in float64 (basic_features.py:241-262).
    """
    names = data.dtype.names or ()
    if "wave" not in names:
        raise ValueError("waveform source is missing the 'wave' field")
    n = len(data)
    wave_dt, wave_off = data.dtype.fields["wave"][0], data.dtype.fields["wave"][1]
    base, shape = wave_dt.subdtype if wave_dt.subdtype else (wave_dt, (1,))
    L = int(shape[0])
    if base == np.int16:
        elem, pool_dtype, signed = 2, np.uint16, True
    elif base == np.uint16:
        elem, pool_dtype, signed = 2, np.uint16, False
    elif base == np.float32:
        elem, pool_dtype, signed = 4, np.float32, False
    else:
        raise ValueError(f"unsupported wave sample dtype {base}")
    itemsize = data.dtype.itemsize
    data = np.ascontiguousarray(data)
    if itemsize % elem or wave_off % elem:
        # unusual layout: fall back to a packed copy of the samples
        pool = np.ascontiguousarray(data["wave"]).reshape(-1).view(pool_dtype)
        offsets = np.arange(n, dtype=np.int64) * L
    else:
        pool = data.view(np.uint8).reshape(-1).view(pool_dtype)
        offsets = (np.arange(n, dtype=np.int64) * itemsize + wave_off) // elem
    # hit_threshold clamps hit edges to event_length (hit_finder.py:388-391); the other consumers
    # always use the whole row (basic_features.py:203, 225)
    clamp = None
    if return_clamp and "event_length" in names and n and np.any(data["event_length"] != L):
        clamp = np.maximum(data["event_length"].astype(np.int64), 0).astype(np.int32)
    rec = np.zeros(n, dtype=RECORDS_DTYPE)
    rec["baseline_upstream"] = np.nan
    rec["timestamp"] = data["timestamp"] if "timestamp" in names else 0
    if "baseline" in names:
        rec["baseline"] = data["baseline"]
    else:
        rec["baseline"] = data["wave"].mean(axis=1, dtype=np.float64) if n else 0.0  # hit_finder.py:188-192
    rec["board"] = data["board"] if "board" in names else 0
    rec["channel"] = data["channel"] if "channel" in names else 0
    rec["record_id"] = data["record_id"] if "record_id" in names else np.arange(n, dtype=np.int64)
    if "polarity" in names:
        pol = np.asarray(data["polarity"]).astype("U8")
        if raw_polarity:
            rec["polarity"] = np.where(pol == "positive", RAW_POSITIVE, "unknown")
        else:
            rec["polarity"] = pol
    else:
        rec["polarity"] = "unknown"
    if "dt" in names:
        rec["dt"] = data["dt"]
    elif explicit_dt is not None:
        rec["dt"] = int(explicit_dt)
    else:
        rec["dt"] = 1
    rec["wave_offset"] = offsets
    rec["event_length"] = L
    rec["time"] = rec["timestamp"] // 1000
    if return_clamp:
        return rec, pool, signed, clamp
    return rec, pool, signed
