"""Seeded synthetic digitizer runs: VX2730-like (dt=2 ns) or V1725-like (dt=4 ns).

Host (numpy) generator used by the tests, the golden-vector script and the CPU legs of
``bench.py``.  Full-size bench inputs are generated on the device by ``wfb_synth_fill``
(csrc/synth.cuh) from a counter-based hash; the two generators are independent (different
random streams) - parity is always checked on identical arrays, never across generators.
"""

from __future__ import annotations

import numpy as np

from .dtypes import RECORDS_DTYPE


def make_raw_run(
    n_channels: int,
    n_per_channel: int,
    n_samples: int,
    *,
    seed: int = 1234,
    dt_ns: int = 2,
    positive_pulses: bool = False,
    n_boards: int = 1,
    coincidence_fraction: float = 0.3,
    pulse_range=(20.0, 500.0),
) -> dict:
    """Return dict(timestamps_ps, boards, channels, samples[int16 (n, L)]) in per-channel order
    (channel-major, time-ascending inside a channel), i.e. the order the reference's per-channel
    files are read in."""
    rng = np.random.default_rng(seed)
    L = int(n_samples)
    n = n_channels * n_per_channel
    quant = 1000 * dt_ns
    ts = np.empty((n_channels, n_per_channel), dtype=np.int64)
    base_q = None
    for c in range(n_channels):
        gaps = rng.exponential(25e6 / quant, size=n_per_channel)  # mean 25 us in ticks
        q = np.floor(np.cumsum(gaps)).astype(np.int64) + 1000
        if c == 0:
            base_q = q
        elif coincidence_fraction > 0:
            pick = rng.random(n_per_channel) < coincidence_fraction
            jitter = rng.integers(-40_000 // quant, 40_000 // quant + 1, size=n_per_channel)
            q = np.where(pick, base_q + jitter, q)
            q.sort()
        q = np.maximum.accumulate(q - np.arange(n_per_channel)) + np.arange(n_per_channel)
        ts[c] = q * quant + c  # +c ps keeps timestamps globally unique
    chans = np.repeat(np.arange(n_channels, dtype=np.int16), n_per_channel)
    boards = (chans % n_boards).astype(np.int16) if n_boards > 1 else np.zeros(n, dtype=np.int16)
    baseline = 8000.0 + 10.0 * chans.astype(np.float64)
    wave = rng.normal(0.0, 3.0, size=(n, L)) + baseline[:, None]
    t = np.arange(L, dtype=np.float64)[None, :]
    n_pulses = rng.integers(1, 4, size=n)
    lo = 100 if L > 400 else max(L // 8, 1)
    hi = L - 200 if L > 400 else max(L // 2, lo + 1)
    sign = 1.0 if positive_pulses else -1.0
    for k in range(3):
        active = n_pulses > k
        amp = rng.uniform(pulse_range[0], pulse_range[1], size=n) * active
        start = rng.integers(lo, hi, size=n).astype(np.float64)
        flat = rng.random(n) < 0.10
        tt = t - start[:, None]
        s1 = np.where(tt >= 0, (1.0 - np.exp(-np.maximum(tt, 0) / 4.0)) * np.exp(-np.maximum(tt, 0) / 30.0), 0.0)
        s2 = ((tt >= 0) & (tt < 200)).astype(np.float64)
        shape = np.where(flat[:, None], s2, s1 * 1.35)
        wave += sign * amp[:, None] * shape
    samples = np.clip(np.rint(wave), 0, 16383).astype(np.int16)
    return dict(
        timestamps_ps=ts.reshape(-1),
        boards=boards,
        channels=chans,
        samples=samples,
        dt_ns=int(dt_ns),
    )


def records_from_raw(raw: dict, *, baseline_window=(0, 40), polarity: str | None = None):
    """Plain-numpy construction of (records, wave_pool) in the reference's global order, for
    feeding plugins in tests without going through the K1 kernel.  (The oracle has its own
    independent restatement in ``oracle.np_oracle.build_records``.)"""
    samples = raw["samples"]
    n, L = samples.shape
    rec = np.zeros(n, dtype=RECORDS_DTYPE)
    rec["timestamp"] = raw["timestamps_ps"]
    rec["board"] = raw["boards"]
    rec["channel"] = raw["channels"]
    s, e = baseline_window
    rec["baseline"] = samples[:, s:e].astype(np.float64).sum(axis=1) / float(e - s)
    rec["baseline_upstream"] = np.nan
    rec["polarity"] = polarity or "unknown"
    rec["dt"] = raw["dt_ns"]
    rec["event_length"] = L
    rec["time"] = rec["timestamp"] // 1000
    order = np.lexsort((np.arange(n), rec["channel"], rec["board"], rec["pid"], rec["timestamp"]))
    rec = rec[order]
    rec["record_id"] = np.arange(n)
    rec["wave_offset"] = np.arange(n, dtype=np.int64) * L
    pool = samples[order].view(np.uint16).reshape(-1).copy()
    return rec, pool


def make_ragged_records(n: int, *, seed: int = 7, min_len: int = 1, max_len: int = 700, dt_ns: int = 2):
    """Variable-length records (V1725-style bundles) for edge-case parity tests."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(min_len, max_len + 1, size=n).astype(np.int32)
    lens[rng.random(n) < 0.05] = 0
    offs = np.zeros(n, dtype=np.int64)
    offs[1:] = np.cumsum(lens[:-1])
    total = int(lens.sum())
    pool = np.clip(np.rint(rng.normal(8000, 4, size=total)), 0, 16383).astype(np.uint16)
    rec = np.zeros(n, dtype=RECORDS_DTYPE)
    for i in range(n):
        Li = int(lens[i])
        if Li > 30 and rng.random() < 0.8:
            s = int(rng.integers(0, Li - 5))
            w = int(rng.integers(2, 60))
            a = int(rng.integers(20, 900))
            seg = pool[offs[i] + s : offs[i] + min(Li, s + w)].astype(np.int64) - a
            pool[offs[i] + s : offs[i] + min(Li, s + w)] = np.clip(seg, 0, 16383)
    rec["timestamp"] = np.cumsum(rng.integers(1, 50_000, size=n)).astype(np.int64) * 1000 * dt_ns
    rec["board"] = rng.integers(0, 2, size=n)
    rec["channel"] = rng.integers(0, 5, size=n)
    rec["baseline"] = 8000.0 + rng.normal(0, 0.7, size=n)
    rec["baseline_upstream"] = np.nan
    pol = rng.integers(0, 3, size=n)
    rec["polarity"] = np.array(["unknown", "positive", "negative"])[pol]
    rec["record_id"] = np.arange(n)
    rec["dt"] = dt_ns
    rec["wave_offset"] = offs
    rec["event_length"] = lens
    rec["time"] = rec["timestamp"] // 1000
    return rec, pool


def make_v1725_blob(*, n_events: int, n_channels: int = 16, seed: int = 0, lengths=(64, 64), tie_every: int = 0, t0: int = 0) -> bytes:
    """A synthetic CAEN V1725 DAW_DEMO .bin stream (layout: utils/formats/v1725.py:62-114 of the
    reference): per event a 16-byte header whose bytes 4 and 11 hold the channel mask, then per fired
    channel a 12-byte header (size in 32-bit words incl. the header: 22 bits; bit 6 of byte 3 = trunc;
    48-bit sample-index timestamp; 16-bit baseline) followed by (size - 3) * 4 bytes of int16 samples.
    ``lengths`` = (min, max) samples per waveform (rounded to even); ``tie_every`` repeats timestamps."""
    rng = np.random.default_rng(seed)
    out = bytearray()
    t = int(t0)
    for ev in range(n_events):
        mask = int(rng.integers(1, 1 << n_channels))
        hdr = bytearray(16)
        hdr[0:4] = int(rng.integers(0, 2**31)).to_bytes(4, "little")  # bytes the reader ignores
        hdr[4] = mask & 0xFF
        hdr[11] = (mask >> 8) & 0xFF
        out += hdr
        if not (tie_every and ev % tie_every == 0):
            t += int(rng.integers(1, 5000))
        for ch in range(16):
            if not (mask >> ch) & 1:
                continue
            ns = int(rng.integers(lengths[0], lengths[1] + 1))
            ns += ns & 1
            size = 3 + ns // 2
            ch_hdr = bytearray(12)
            ch_hdr[0] = size & 0xFF
            ch_hdr[1] = (size >> 8) & 0xFF
            ch_hdr[2] = (size >> 16) & 0x3F
            ch_hdr[3] = (0x40 if rng.random() < 0.2 else 0) | int(rng.integers(0, 64))
            ts = t + (0 if (tie_every and ev % tie_every == 0) else int(rng.integers(0, 3)))
            ch_hdr[4:10] = int(ts).to_bytes(6, "little")
            ch_hdr[10:12] = int(rng.integers(0, 16384)).to_bytes(2, "little")
            out += ch_hdr
            wave = (8000 + rng.normal(0, 30, ns)).astype(np.int16)
            wave[:: max(ns // 5, 1)] -= 9000  # some genuinely negative int16 samples
            out += wave.tobytes()
    return bytes(out)
