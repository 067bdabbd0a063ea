"""numpy restatement of the reference hot path (TEST INFRASTRUCTURE - see oracle/__init__.py).

Each function cites the reference file:line it restates (paths relative to
``/root/reference/waveform_analysis``).  The restatements are written for clarity and for
vectorised execution over many records; they are not copies of the reference loops.
"""

from __future__ import annotations

import os

import numpy as np

from waveformanalysis_b200.dtypes import (
    BASIC_FEATURES_DTYPE,
    HIT_MERGE_CLUSTERS_DTYPE,
    HIT_MERGED_COMPONENTS_DTYPE,
    HIT_MERGED_DTYPE,
    RECORDS_DTYPE,
    THRESHOLD_HIT_DTYPE,
    WAVEFORM_WIDTH_DTYPE,
    WAVEFORM_WIDTH_INTEGRAL_DTYPE,
)

# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------


def resolve_slice(start, end, length: int) -> tuple[int, int]:
    """Python ``seq[start:end]`` bounds for a sequence of ``length`` (basic_features.py:155-156)."""
    lo, hi, _ = slice(start, end).indices(int(length))
    return lo, max(lo, hi)


def _is_fixed_contiguous(records: np.ndarray) -> int | None:
    """Return L when every record has event_length L and wave_offset == i*L, else None."""
    if len(records) == 0:
        return None
    ln = records["event_length"]
    L = int(ln[0])
    if L <= 0 or np.any(ln != L):
        return None
    off = records["wave_offset"]
    if np.array_equal(off, np.arange(len(records), dtype=np.int64) * L):
        return L
    return None


def _record_wave(records, pool, i):
    o = int(records["wave_offset"][i])
    return pool[o : o + int(records["event_length"][i])]


# --------------------------------------------------------------------------------------
# K1: records builder (core/processing/records_builder.py)
# --------------------------------------------------------------------------------------


def baseline_mean(samples: np.ndarray, start: int, end: int) -> np.ndarray:
    """records.baseline = mean(float64(samples[:, start:end])) (records_builder.py:243-257).

    ``samples`` is the (n, L) sample matrix (header columns already stripped), so the VX2730
    default window raw[:, 7:47] is ``start=0, end=40``.  Empty window -> NaN.
    """
    n, L = samples.shape
    end = min(int(end), L)
    if end <= start:
        return np.full(n, np.nan, dtype=np.float64)
    return np.mean(samples[:, start:end].astype(float), axis=1)


def build_records(
    timestamps_ps: np.ndarray,
    boards: np.ndarray,
    channels: np.ndarray,
    samples: np.ndarray,
    *,
    dt_ns: int,
    baseline_window: tuple[int, int] = (0, 40),
    baselines: np.ndarray | None = None,
    epoch_ns: int | None = None,
    flags: np.ndarray | None = None,
) -> tuple[np.ndarray, np.ndarray]:
    """Fixed-length records + wave_pool in the reference's global order.

    Restates ``_build_records_part_from_raw_array`` (records_builder.py:212-302) followed by
    ``_merge_records_part_refs`` (:341-426): the merged order is ascending
    (timestamp, pid=0, board, channel, input order), which is what one global
    ``lexsort((seq, channel, board, pid, timestamp))`` (:115-120) gives when the inputs are
    concatenated in part order.  ``record_id`` = arange after the merge (:425), ``wave_offset``
    = record_id * L (:300, 408), ``time`` = timestamp // 1000 (+ epoch) (:279, 507-510),
    wave_pool = int16 samples reinterpreted as uint16 (:108-112, 299).
    """
    n, L = samples.shape
    rec = np.zeros(n, dtype=RECORDS_DTYPE)
    rec["timestamp"] = np.asarray(timestamps_ps, dtype=np.int64)
    rec["pid"] = 0
    rec["board"] = np.asarray(boards, dtype=np.int16)
    rec["channel"] = np.asarray(channels, dtype=np.int16)
    if baselines is None:
        rec["baseline"] = baseline_mean(samples, baseline_window[0], baseline_window[1])
    else:
        rec["baseline"] = np.asarray(baselines, dtype=np.float64)
    rec["baseline_upstream"] = np.nan
    rec["polarity"] = "unknown"
    rec["dt"] = np.int32(dt_ns)
    rec["trigger_type"] = 0
    rec["flags"] = 0 if flags is None else np.asarray(flags, dtype=np.uint32)
    rec["event_length"] = np.int32(L)
    rec["time"] = rec["timestamp"] // 1000
    if epoch_ns is not None:
        rec["time"] += np.int64(epoch_ns)
    seq = np.arange(n, dtype=np.int64)
    order = np.lexsort((seq, rec["channel"], rec["board"], rec["pid"], rec["timestamp"]))
    rec = rec[order]
    rec["record_id"] = np.arange(n, dtype=np.int64)
    rec["wave_offset"] = np.arange(n, dtype=np.int64) * L
    pool = np.ascontiguousarray(samples[order]).astype(np.uint16, copy=False).reshape(-1)
    return rec, pool


# --------------------------------------------------------------------------------------
# basic_features (core/plugins/builtin/cpu/basic_features.py:108-195, records branch)
# --------------------------------------------------------------------------------------


def basic_features(
    records: np.ndarray,
    pool: np.ndarray,
    *,
    height_range=(40, 90),
    area_range=(0, None),
    fixed_baseline: dict | None = None,
) -> np.ndarray:
    """Restates the records branch of BasicFeaturesPlugin.compute.

    ``fixed_baseline`` maps (board, channel) -> float override (the resolved
    ``channel_config`` value, basic_features.py:134-146).  ``pool`` is the uint16 wave_pool or
    the float32 wave_pool_filtered.
    """
    n = len(records)
    out = np.zeros(n, dtype=BASIC_FEATURES_DTYPE)
    if n == 0:
        return out
    base = records["baseline"].astype(np.float64).copy()
    if fixed_baseline:
        for (b, c), v in fixed_baseline.items():
            if v is None:
                continue
            sel = (records["board"] == b) & (records["channel"] == c)
            base[sel] = float(v)
    pol = np.asarray(records["polarity"])
    known = (pol == "positive") | (pol == "negative")
    positive = pol == "positive"

    out["timestamp"] = records["timestamp"]
    out["board"] = records["board"]
    out["channel"] = records["channel"]
    out["event_index"] = np.arange(n, dtype=np.int64)

    L = _is_fixed_contiguous(records)
    if L is not None:
        W = pool[: n * L].reshape(n, L)
        p0, p1 = resolve_slice(height_range[0], height_range[1], L)
        c0, c1 = resolve_slice(area_range[0], area_range[1], L)
        _basic_block(out, W, base, known, positive, p0, p1, c0, c1)
        return out
    for i in range(n):
        w = _record_wave(records, pool, i)
        Li = len(w)
        p0, p1 = resolve_slice(height_range[0], height_range[1], Li)
        c0, c1 = resolve_slice(area_range[0], area_range[1], Li)
        _basic_block(
            out[i : i + 1], w.reshape(1, Li), base[i : i + 1], known[i : i + 1],
            positive[i : i + 1], p0, p1, c0, c1,
        )
    return out


def _basic_block(out, W, base, known, positive, p0, p1, c0, c1):
    n, L = W.shape
    height = np.zeros(n, dtype=np.float64)
    amp = np.zeros(n, dtype=np.float64)
    area = np.zeros(n, dtype=np.float64)
    unk = ~known
    if np.any(unk):
        Wu = W[unk]
        bu = base[unk]
        if p1 > p0:
            wmin = Wu[:, p0:p1].min(axis=1).astype(np.float64)
            wmax = Wu[:, p0:p1].max(axis=1).astype(np.float64)
            height[unk] = bu - wmin  # basic_features.py:174 (unknown -> "negative")
            amp[unk] = wmax - wmin  # :175
        if c1 > c0:
            area[unk] = np.sum(bu[:, None] - Wu[:, c0:c1].astype(np.float64), axis=1)  # :180-185
    if np.any(known):
        Wk = W[known]
        b32 = base[known].astype(np.float32)
        s = Wk.astype(np.float32) - b32[:, None]  # records_view.py:94-96
        flip = ~positive[known]  # negative: s = -(w-b); positive: -(-(w-b)) = w-b (:98-99 + bf:151)
        s[flip] = -s[flip]
        if p1 > p0:
            smax = s[:, p0:p1].max(axis=1).astype(np.float64)
            smin = s[:, p0:p1].min(axis=1).astype(np.float64)
            height[known] = smax  # :166
            amp[known] = smax - smin  # :167
        if c1 > c0:
            area[known] = np.sum(s[:, c0:c1].astype(np.float64), axis=1)  # :178
    out["height"] = height
    out["amp"] = amp
    out["area"] = area
    if L > 1:
        d = np.abs(np.diff(W.astype(np.float64), axis=1))  # :187-189
        out["max_abs_diff"] = d.max(axis=1)


# --------------------------------------------------------------------------------------
# hit_threshold (core/plugins/builtin/cpu/hit_finder.py:122-255, 329-413)
# --------------------------------------------------------------------------------------


def threshold_hits(
    records: np.ndarray,
    pool: np.ndarray,
    *,
    threshold: float = 10.0,
    thresholds: dict | None = None,
    left_extension: int = 2,
    right_extension: int = 2,
    block: int = 4096,
) -> np.ndarray:
    """Restates ThresholdHitPlugin.compute (records branch) + _build_hits_from_signal_matrix.

    ``thresholds`` maps (board, channel) -> per-channel threshold override
    (hit_finder.py:288-327).  Rows come out record-major, then by start sample (:352-355).
    """
    n = len(records)
    if n == 0:
        return np.zeros(0, dtype=THRESHOLD_HIT_DTYPE)
    left = max(0, int(left_extension))
    right = max(0, int(right_extension))
    lens = records["event_length"].astype(np.int64)
    Lmax = int(lens.max())  # padded matrix width (records_view.py:179-189, hit_finder.py:364)
    thr = np.full(n, float(threshold), dtype=np.float64)
    if thresholds:
        for (b, c), v in thresholds.items():
            sel = (records["board"] == b) & (records["channel"] == c)
            thr[sel] = float(v)
    positive = np.asarray(records["polarity"]) == "positive"  # hit_finder.py:323-325
    base = records["baseline"].astype(np.float64)
    fixedL = _is_fixed_contiguous(records)
    rows = []
    for r0 in range(0, n, block):
        r1 = min(n, r0 + block)
        m = r1 - r0
        W = np.zeros((m, Lmax), dtype=np.float64)  # padding is 0.0 (records_view.py:189)
        if fixedL is not None:
            W[:, :] = pool[r0 * fixedL : r1 * fixedL].reshape(m, fixedL)
        else:
            for k in range(m):
                w = _record_wave(records, pool, r0 + k)
                W[k, : len(w)] = w
        b = base[r0:r1, None]
        sig = np.where(positive[r0:r1, None], W - b, b - W)  # :240-241
        valid = np.arange(Lmax)[None, :] < lens[r0:r1, None]
        msk = (sig >= thr[r0:r1, None]) & valid  # :346-348
        edge = np.zeros((m, Lmax + 2), dtype=np.int8)
        edge[:, 1:-1] = msk
        flips = np.flatnonzero(edge[:, 1:] != edge[:, :-1])  # alternating start, end per row
        if flips.size == 0:
            continue
        rr, cc = np.divmod(flips, Lmax + 1)
        starts, ends, rws = cc[0::2], cc[1::2], rr[0::2]
        for s, e, k in zip(starts.tolist(), ends.tolist(), rws.tolist()):
            i = r0 + k
            a0 = max(0, s - left)
            a1 = min(Lmax, e + right)  # :369-370
            if a1 <= a0:
                continue
            seg = sig[k, a0:a1]
            rel = int(np.argmax(seg))  # first maximum (:378)
            p = a0 + rel
            dt_ns = int(records["dt"][i])
            ts = np.int64(records["timestamp"][i]) + p * (float(dt_ns) * 1e3)  # f64 (:383-386)
            rl = max(int(lens[i]), 0)
            es = min(max(a0, 0), rl)
            ee = max(min(max(a1, 0), rl), es)  # :388-391
            rows.append(
                (
                    p,
                    float(seg[rel]),
                    float(np.sum(np.maximum(seg, 0.0))),
                    es,
                    ee,
                    float(ee - es),
                    dt_ns,
                    float(max(p - s, 0) * dt_ns),
                    float(max((e - 1) - p, 0) * dt_ns),
                    int(ts),
                    int(records["board"][i]),
                    int(records["channel"][i]),
                    int(records["record_id"][i]),
                )
            )
    if not rows:
        return np.zeros(0, dtype=THRESHOLD_HIT_DTYPE)
    return np.array(rows, dtype=THRESHOLD_HIT_DTYPE)


# --------------------------------------------------------------------------------------
# wave_pool_filtered (core/plugins/builtin/cpu/filtering.py:181-241, records.py:368-438)
# --------------------------------------------------------------------------------------


def sg_coefficients(window: int, poly: int) -> np.ndarray:
    """Savitzky-Golay smoothing (deriv 0) FIR taps.

    scipy.signal.savgol_coeffs (scipy 1.18.1, signal/_savitzky_golay.py:14-138, an
    un-vendored dependency: pyproject.toml:31 ``scipy>=1.7.0``) solves the least-squares
    system A c = e0 with A[k, j] = x_j**k for x = halflen .. -halflen; restated here.
    """
    h = window // 2
    x = np.arange(-h, window - h, dtype=float)[::-1]
    A = x ** np.arange(poly + 1).reshape(-1, 1)
    y = np.zeros(poly + 1)
    y[0] = 1.0
    c, *_ = np.linalg.lstsq(A, y, rcond=None)
    return c


def sg_edge_matrix(window: int, poly: int) -> np.ndarray:
    """(window x window) projector P = A pinv(A): the degree-``poly`` LSQ fit over ``window``
    samples evaluated at each of the window positions (savgol mode='interp' edge rule,
    scipy/signal/_savitzky_golay.py:141-186)."""
    A = np.vander(np.arange(window, dtype=float), poly + 1)
    return A @ np.linalg.pinv(A)


def effective_sg_window(n_samples: int, sg_window: int, poly: int) -> int | None:
    """filtering.py:181-195: window = min(sg_window, L) forced odd; None (identity) if <= poly."""
    w = min(int(sg_window), int(n_samples))
    if w % 2 == 0:
        w -= 1
    if w <= int(poly):
        return None
    return w


def sg_filter_rows(X: np.ndarray, sg_window: int, poly: int) -> np.ndarray:
    """SG 'interp' filter of every row of float32 X (n, L) -> float32 (filtering.py:226-241).

    Interior: f64-accumulated correlation with the symmetric taps, stored f32 (bit-exact with
    scipy on interior samples).  Edges: closed-form projector in f64 - scipy runs
    that fit through polyfit on f32 data, so the 2*halflen edge samples agree to rel ~1e-6,
    not bit-exactly.
    """
    X = np.asarray(X, dtype=np.float32)
    n, L = X.shape
    w = effective_sg_window(L, sg_window, poly)
    if w is None:
        return X.copy()
    h = w // 2
    c = sg_coefficients(w, poly)
    Xd = X.astype(np.float64)
    out = np.empty((n, L), dtype=np.float64)
    acc = np.zeros((n, L - 2 * h), dtype=np.float64)
    for j in range(w):
        acc += c[w - 1 - j] * Xd[:, j : j + L - 2 * h]
    out[:, h : L - h] = acc
    P = sg_edge_matrix(w, poly)
    out[:, :h] = Xd[:, :w] @ P[:h].T
    out[:, L - h :] = Xd[:, L - w :] @ P[w - h :].T
    return out.astype(np.float32)


def bw_padlen(sos: np.ndarray) -> int:
    """filtering.py:198-203 (= scipy sosfiltfilt default padlen)."""
    ns = int(sos.shape[0])
    return 3 * (2 * ns + 1 - min(int((sos[:, 2] == 0).sum()), int((sos[:, 5] == 0).sum())))


def sos_steady_state(sos: np.ndarray) -> np.ndarray:
    """Step-response steady state of each DF2T biquad, scaled by the DC gain of the sections
    before it.  Restates scipy.signal.sosfilt_zi / lfilter_zi as of scipy 1.18.1 (the version
    in the build image; the reference leaves scipy unpinned, pyproject.toml:31):
    y_inf = sum(b)/sum(a); zi = reversed cumulative sum of (b - y_inf*a), first entry dropped.
    """
    ns = sos.shape[0]
    zi = np.zeros((ns, 2))
    scale = 1.0
    for s in range(ns):
        b = sos[s, :3]
        a = sos[s, 3:]
        if a[0] != 1:
            b, a = b / a[0], a / a[0]
        y_inf = np.sum(b) / np.sum(a)
        v = b - y_inf * a
        zi[s, 1] = scale * v[2]
        zi[s, 0] = scale * (v[2] + v[1])
        scale *= np.sum(sos[s, :3]) / np.sum(sos[s, 3:])
    return zi


def bw_filter_rows(X: np.ndarray, sos: np.ndarray, zi: np.ndarray | None = None) -> np.ndarray:
    """Zero-phase cascade of biquads on every row of float32 X, float64 inside, f32 out.

    Restates scipy.signal.sosfiltfilt as called at filtering.py:218-224: odd extension by
    ``padlen`` samples, DF2T forward pass started from zi*x0, reversed pass started from
    zi*y_last, extension dropped.  Separate multiply and add (no FMA) reproduces scipy
    bit-for-bit.  Rows of length <= padlen are returned unfiltered (:221-222).
    """
    X = np.asarray(X, dtype=np.float32)
    n, L = X.shape
    edge = bw_padlen(sos)
    if L <= edge:
        return X.copy()
    if zi is None:
        zi = sos_steady_state(sos)
    x = X.astype(np.float64)
    left = 2.0 * x[:, :1] - x[:, edge:0:-1]
    right = 2.0 * x[:, -1:] - x[:, -2 : -edge - 2 : -1]
    ext = np.concatenate([left, x, right], axis=1)

    def cascade(sig, x0):
        y = sig
        for s in range(sos.shape[0]):
            b0, b1, b2, _a0, a1, a2 = sos[s]
            z0 = zi[s, 0] * x0
            z1 = zi[s, 1] * x0
            o = np.empty_like(y)
            for t in range(y.shape[1]):
                xt = y[:, t]
                yt = b0 * xt + z0
                z0 = (b1 * xt - a1 * yt) + z1
                z1 = b2 * xt - a2 * yt
                o[:, t] = yt
            y = o
        return y

    # scipy: zi is scaled by the *input* edge value x0 for every section
    fwd = cascade(ext, ext[:, 0])
    rev = cascade(fwd[:, ::-1], fwd[:, -1])
    y = rev[:, ::-1][:, edge:-edge]
    return y.astype(np.float32)


def wave_pool_filtered(records, pool, *, configs: dict, default: dict) -> np.ndarray:
    """Per-(board,channel) filter of every record (records.py:368-438).

    ``default`` / ``configs[(board, channel)]`` are dicts with ``filter_type`` 'SG'
    (``sg_window_size``, ``sg_poly_order``) or 'BW' (``sos`` ndarray).
    """
    out = np.zeros(len(pool), dtype=np.float32)
    if len(records) == 0 or len(pool) == 0:
        return out
    keys = np.stack([records["board"].astype(np.int64), records["channel"].astype(np.int64)], 1)
    for key in {tuple(k) for k in keys.tolist()}:
        cfg = configs.get(key, default)
        idx = np.flatnonzero((keys[:, 0] == key[0]) & (keys[:, 1] == key[1]))
        lens = records["event_length"][idx]
        for Lk in np.unique(lens):
            if Lk <= 0:
                continue
            sub = idx[lens == Lk]
            offs = records["wave_offset"][sub].astype(np.int64)
            gather = offs[:, None] + np.arange(int(Lk))[None, :]
            X = pool[gather].astype(np.float32)
            if cfg["filter_type"] == "BW":
                Y = bw_filter_rows(X, np.asarray(cfg["sos"], dtype=np.float64))
            else:
                Y = sg_filter_rows(X, cfg["sg_window_size"], cfg["sg_poly_order"])
            out[gather] = Y
    return out


# --------------------------------------------------------------------------------------
# waveform_width (core/plugins/builtin/cpu/waveform_width.py:97-194, 205-374)
# --------------------------------------------------------------------------------------


def _crossing(y: np.ndarray, thr, rising: bool, interpolation: bool):
    if len(y) == 0:
        return None
    idx = np.flatnonzero(y >= thr) if rising else np.flatnonzero(y <= thr)
    if idx.size == 0:
        return None
    i = int(idx[0])
    if not interpolation or i == 0:
        return float(i)
    y0, y1 = y[i - 1], y[i]
    if abs(y1 - y0) < 1e-10:
        return float(i)
    return float(i - 1) + (thr - y0) / (y1 - y0)


def waveform_width(
    hits: np.ndarray,
    wave_record_ids: np.ndarray,
    waves: np.ndarray,
    *,
    sampling_rate: float | None = None,
    rise_low=0.1,
    rise_high=0.9,
    fall_high=0.9,
    fall_low=0.1,
    interpolation=True,
) -> np.ndarray:
    """Restates WaveformWidthPlugin.compute for ``hits`` (HIT_DTYPE-like: position, timestamp,
    board, channel, record_id) over AoS waves (n, L) int16 (st_waveforms) or float32
    (filtered_waveforms); ``wave_record_ids[i]`` is the record_id of row i (first match wins,
    waveform_width.py:164-168).  Arithmetic dtype follows numpy promotion exactly as the
    reference: int16 rows -> float64, float32 rows -> float32 (:240-241)."""
    if sampling_rate is None:
        sampling_rate = 0.5
    rows = []
    first_row: dict[int, int] = {}
    for i, rid in enumerate(np.asarray(wave_record_ids).tolist()):
        first_row.setdefault(int(rid), i)
    for hrow in hits:
        rid = int(hrow["record_id"])
        if rid not in first_row:
            continue
        w = waves[first_row[rid]]
        pos = hrow["position"]
        bl = np.mean(w[:50])
        wc = w - bl
        if pos >= len(wc):
            continue
        pv = wc[pos]
        if pv <= 0:
            continue
        left, right = wc[:pos], wc[pos:]
        rl = _crossing(left, pv * rise_low, True, interpolation)
        rh = _crossing(left, pv * rise_high, True, interpolation)
        rts = (rh - rl) if (rl is not None and rh is not None) else 0.0
        fh = _crossing(right, pv * fall_high, False, interpolation)
        fl = _crossing(right, pv * fall_low, False, interpolation)
        if fh is not None and fl is not None:
            fh += pos
            fl += pos
            fts = fl - fh
        else:
            fts = 0.0
        tws = (fl - rl) if (rl is not None and fl is not None) else 0.0
        rows.append(
            (
                float(rts / sampling_rate),
                float(fts / sampling_rate),
                float(tws / sampling_rate),
                float(rts),
                float(fts),
                float(tws),
                int(pos),
                float(pv),
                int(hrow["timestamp"]),
                int(hrow["board"]) if "board" in hrow.dtype.names else 0,
                int(hrow["channel"]),
                rid,
            )
        )
    if not rows:
        return np.zeros(0, dtype=WAVEFORM_WIDTH_DTYPE)
    return np.array(rows, dtype=WAVEFORM_WIDTH_DTYPE)


# --------------------------------------------------------------------------------------
# waveform_width_integral (core/plugins/builtin/cpu/waveform_width_integral.py:166-231)
# --------------------------------------------------------------------------------------


def width_integral(
    records: np.ndarray,
    pool: np.ndarray,
    *,
    q_low=0.10,
    q_high=0.90,
    sampling_rate=0.5,
    dt=None,
) -> np.ndarray:
    """Records branch: cumulative-charge quantile indices per record."""
    if dt is None:
        dt = 1.0 / float(sampling_rate)
    n = len(records)
    out = np.zeros(n, dtype=WAVEFORM_WIDTH_INTEGRAL_DTYPE)
    for i in range(n):
        w = _record_wave(records, pool, i)
        b = float(records["baseline"][i])
        pol = str(records["polarity"][i])
        if pol in ("positive", "negative"):
            s = w.astype(np.float32) - np.float32(b)  # records_view.signals
            if pol == "positive":
                s = -s
            sig = -s.astype(np.float64)  # :185
        else:
            raw = w.astype(np.float64) - b
            sig = raw if pol == "positive" else -raw  # :187-191
        x = np.maximum(sig, 0.0)
        q = float(np.sum(x))
        if q <= 0 or not np.isfinite(q):
            lo = hi = 0
        else:
            cs = np.cumsum(x)
            lo = int(np.searchsorted(cs, q_low * q, side="left"))
            hi = int(np.searchsorted(cs, q_high * q, side="left"))
        ws = float(max(hi - lo, 0))
        out[i] = (
            float(lo) * dt, float(hi) * dt, ws * dt, float(lo), float(hi), ws, q,
            int(records["timestamp"][i]), int(records["board"][i]), int(records["channel"][i]), i,
        )
    return out


# --------------------------------------------------------------------------------------
# hit_merge (core/plugins/builtin/cpu/hit_merge.py:115-181, 256-322, 353-534)
# --------------------------------------------------------------------------------------


def hit_abs_windows(hits: np.ndarray, start_name="edge_start", end_name="edge_end"):
    """abs_start/abs_end in ps as float64 (hit_merge.py:75-81, event_grouping.py:365-367)."""
    ts = hits["timestamp"].astype(np.float64)
    pos = hits["position"].astype(np.float64)
    dt_ps = hits["dt"].astype(np.float64) * 1e3
    a0 = ts + (hits[start_name].astype(np.float64) - pos) * dt_ps
    a1 = ts + (hits[end_name].astype(np.float64) - pos) * dt_ps
    return a0, a1


def hit_merge(hits: np.ndarray, *, merge_gap_ns=0.0, max_total_width_ns=10000.0):
    """Returns (hit_merge_clusters, hit_merged, hit_merged_components).

    Per hardware channel in ascending (board, channel): stable sort by abs_start, chain while
    merge_gap_ns > 0, same dt, gap <= merge_gap, total width <= max width (hit_merge.py:137-179).
    Cluster row = the single hit, or for multi-hit clusters the anchor = highest hit (earliest
    timestamp among equal heights), summed integral, sample window only when all members share
    a record (:256-322).
    """
    nh = len(hits)
    if nh == 0:
        return (
            np.zeros(0, dtype=HIT_MERGE_CLUSTERS_DTYPE),
            np.zeros(0, dtype=HIT_MERGED_DTYPE),
            np.zeros(0, dtype=HIT_MERGED_COMPONENTS_DTYPE),
        )
    a0, a1 = hit_abs_windows(hits)
    key = hits["board"].astype(np.int64) * 65536 + (hits["channel"].astype(np.int64) & 0xFFFF)
    # board is the primary key, signed ordering as in the reference's structured sort
    order = np.lexsort((np.arange(nh), a0, hits["channel"].astype(np.int32), hits["board"].astype(np.int32)))
    del key
    gap_ps = merge_gap_ns * 1e3
    maxw_ps = max_total_width_ns * 1e3
    cluster_of = np.zeros(nh, dtype=np.int64)  # in sorted order
    cid = -1
    prev_b = prev_c = None
    for k, i in enumerate(order.tolist()):
        b, c = int(hits["board"][i]), int(hits["channel"][i])
        new = True
        if (b, c) == (prev_b, prev_c) and merge_gap_ns > 0:
            nxt_end = max(c_end, a1[i])
            if hits["dt"][i] == last_dt and (a0[i] - c_end) <= gap_ps and (nxt_end - c_start) <= maxw_ps:
                new = False
                c_end = nxt_end
        if new:
            cid += 1
            c_start, c_end = a0[i], a1[i]
        last_dt = hits["dt"][i]
        prev_b, prev_c = b, c
        cluster_of[k] = cid
    clusters = np.zeros(nh, dtype=HIT_MERGE_CLUSTERS_DTYPE)
    clusters["cluster_index"] = cluster_of
    clusters["hit_index"] = order
    ncl = cid + 1
    merged = np.zeros(ncl, dtype=HIT_MERGED_DTYPE)
    bounds = np.flatnonzero(np.r_[True, cluster_of[1:] != cluster_of[:-1], True])
    for ci in range(ncl):
        s, e = int(bounds[ci]), int(bounds[ci + 1])
        members = order[s:e]
        if e - s == 1:
            h = hits[members[0]]
            merged[ci] = (
                h["position"], h["height"], h["integral"], h["edge_start"], h["edge_end"], h["width"],
                h["dt"], h["rise_time"], h["fall_time"], h["timestamp"], h["board"], h["channel"],
                h["record_id"], s, 1,
            )
            continue
        hh = hits[members]
        hts = hh["height"].astype(np.float64)
        cand = np.flatnonzero(hts == hts.max())
        anchor = cand[0] if len(cand) == 1 else cand[np.argmin(hh["timestamp"][cand])]
        a = hh[anchor]
        if len(set(hh["record_id"].tolist())) == 1:
            ss, se = int(hh["edge_start"].min()), int(hh["edge_end"].max())
        else:
            ss, se = -1, -1
        width = float(max(se - ss, 0.0))
        if ss < 0 or se < 0:
            width = -1.0
        integ = float(np.sum([float(v) for v in hh["integral"]]))
        merged[ci] = (
            a["position"], float(hts.max()), integ, ss, se, width, a["dt"], a["rise_time"],
            a["fall_time"], a["timestamp"], a["board"], a["channel"], a["record_id"], s, e - s,
        )
    comps = np.zeros(nh, dtype=HIT_MERGED_COMPONENTS_DTYPE)
    comps["merged_index"] = cluster_of
    comps["hit_index"] = order
    return clusters, merged, comps


# --------------------------------------------------------------------------------------
# event grouping (core/processing/event_grouping.py)
# --------------------------------------------------------------------------------------


def merged_abs_windows(hits, component_rows=None, component_hits=None):
    """Absolute windows of hit_merged rows; rows merged across records carry sample_start/end = -1 and take
    min / max of their component hits' windows (event_grouping.py:365-414)."""
    names = hits.dtype.names
    sn, en = ("sample_start", "sample_end") if "sample_start" in names else ("edge_start", "edge_end")
    a0, a1 = hit_abs_windows(hits, sn, en)
    bad = np.flatnonzero((hits[sn] < 0) | (hits[en] < 0))
    if len(bad):
        if component_rows is None or component_hits is None:
            raise ValueError("component_rows and component_hits are required when hit windows contain invalid edges")
        c0, c1 = hit_abs_windows(component_hits, "edge_start", "edge_end")
        idx = component_rows["hit_index"].astype(np.int64)
        for m in bad.tolist():
            o, c = int(hits["component_offset"][m]), int(hits["component_count"][m])
            if c <= 0:
                raise ValueError(f"missing hit_merged_components rows for hit_merged index {m}")
            a0[m] = c0[idx[o:o + c]].min()
            a1[m] = c1[idx[o:o + c]].max()
    return a0, a1


def group_hit_windows(hits: np.ndarray, time_window_ns: float, component_rows=None, component_hits=None) -> dict:
    """Restates group_hit_windows (event_grouping.py:287-471) as a prefix-max scan; rows without a valid
    sample window use their components (merged_abs_windows).

    Returns dict with per-event arrays (event_id, t_min, t_max, dt_ns, n_hits), ``offsets``
    (n_events+1) and ``members`` (hit indices, event-major, each event ordered by
    (board, channel, dt, abs_start, timestamp, record_id), :423-432), plus ``event_of_hit``.
    """
    names = hits.dtype.names
    sn, en = ("sample_start", "sample_end") if "sample_start" in names else ("edge_start", "edge_end")
    nh = len(hits)
    if nh == 0:
        z = np.zeros(0, dtype=np.int64)
        return dict(event_id=z, t_min=z, t_max=z, dt_ns=np.zeros(0), n_hits=z, offsets=np.zeros(1, np.int64), members=z, event_of_hit=z)
    a0, a1 = merged_abs_windows(hits, component_rows, component_hits)
    ts = hits["timestamp"].astype(np.int64)
    dtv = hits["dt"].astype(np.int32)
    rid = hits["record_id"].astype(np.int64)
    order = np.lexsort((rid, ts, dtv, a0))  # :418
    gap = time_window_ns * 1e3
    pm = np.maximum.accumulate(a1[order])
    new = np.ones(nh, dtype=bool)
    new[1:] = a0[order][1:] > pm[:-1] + gap  # :462
    ev_sorted = np.cumsum(new) - 1
    event_of_hit = np.empty(nh, dtype=np.int64)
    event_of_hit[order] = ev_sorted
    n_ev = int(ev_sorted[-1]) + 1
    morder = np.lexsort((rid, ts, a0, dtv, hits["channel"].astype(np.int16), hits["board"].astype(np.int16), event_of_hit))
    counts = np.bincount(event_of_hit, minlength=n_ev)
    offsets = np.zeros(n_ev + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    tmin = np.full(n_ev, np.inf)
    tmax = np.full(n_ev, -np.inf)
    np.minimum.at(tmin, event_of_hit, a0)
    np.maximum.at(tmax, event_of_hit, a1)
    t_min = tmin.astype(np.int64)  # int() truncation (:434-435)
    t_max = tmax.astype(np.int64)
    return dict(
        event_id=np.arange(n_ev, dtype=np.int64),
        t_min=t_min,
        t_max=t_max,
        dt_ns=(t_max - t_min) / 1e3,
        n_hits=counts.astype(np.int64),
        offsets=offsets,
        members=morder.astype(np.int64),
        event_of_hit=event_of_hit,
    )


def group_time_window(timestamps: np.ndarray, channels: np.ndarray, time_window_ns: float) -> dict:
    """Restates group_multi_channel_hits (event_grouping.py:99-283, 477-510).

    Input must have unique timestamps (the reference sorts with an unstable quicksort, SURVEY
    'unstable sorts').  Anchored windows: next anchor = first ts > ts[anchor] + W (float64
    compare).  Members ordered by channel; t_min / t_max are the timestamps of the lowest /
    highest channel member (:250-257).
    """
    ts_in = np.asarray(timestamps, dtype=np.int64)
    n = len(ts_in)
    order = np.argsort(ts_in, kind="stable")
    ts = ts_in[order]
    ch = np.asarray(channels)[order]
    W = time_window_ns * 1e3
    bounds = [0]
    cur = 0
    while cur < n:
        cur = int(np.searchsorted(ts.astype(np.float64), float(ts[cur]) + W, side="right"))
        bounds.append(cur)
    bounds = np.asarray(bounds, dtype=np.int64)
    n_ev = len(bounds) - 1
    members = np.empty(n, dtype=np.int64)
    t_min = np.zeros(n_ev, dtype=np.int64)
    t_max = np.zeros(n_ev, dtype=np.int64)
    for e in range(n_ev):
        s, t = int(bounds[e]), int(bounds[e + 1])
        o = np.argsort(ch[s:t], kind="stable")
        members[s:t] = order[s:t][o]
        t_min[e] = ts[s:t][o][0]
        t_max[e] = ts[s:t][o][-1]
    return dict(
        event_id=np.arange(n_ev, dtype=np.int64),
        t_min=t_min,
        t_max=t_max,
        dt_ns=(t_max - t_min) / 1e3,
        n_hits=np.diff(bounds),
        offsets=bounds,
        members=members,
    )


# --------------------------------------------------------------------------------------------
# CAEN V1725 DAW_DEMO binary ingest
# --------------------------------------------------------------------------------------------
def v1725_board_from_name(name: str) -> int:
    """Board id from the file name, 0 when absent (utils/formats/v1725.py:62-67)."""
    import re

    m = re.search(r"_b(\d+)", os.path.basename(str(name)), flags=re.IGNORECASE)
    return int(m.group(1)) if m else 0


def v1725_scan(blob) -> dict:
    """Walk the event / channel header chain of one .bin stream (utils/formats/v1725.py:69-114): per
    event a 16-byte header (channel mask = byte 4 | byte 11 << 8), per fired channel (ascending) a
    12-byte header - size in 32-bit words = low 22 bits of bytes 0..2, trunc = bit 6 of byte 3,
    timestamp = bytes 4..9, baseline = bytes 10..11 - then (size - 3) * 4 payload bytes.  Parsing stops
    at the first short read, exactly like the reference's reader."""
    b = bytes(blob)
    pos, n = 0, len(b)
    off, ns, ch, ts, bl, tr = [], [], [], [], [], []
    while True:
        if n - pos <= 0:
            break
        if n - pos < 16:
            break
        hdr = b[pos:pos + 16]
        pos += 16
        mask = hdr[4] + (hdr[11] << 8)
        stop = False
        for c in range(16):
            if not (mask >> c) & 1:
                continue
            if n - pos < 12:
                pos = n  # the reader has consumed the partial header
                stop = True
                break
            h = b[pos:pos + 12]
            pos += 12
            size = int.from_bytes(h[:3], "little") & ((1 << 22) - 1)
            sig = (size - 3) << 2
            if sig < 0:
                raise ValueError("V1725 channel header with size < 3 words")
            if n - pos < sig:
                pos = n
                stop = True
                break
            off.append(pos)
            ns.append(sig // 2)
            ch.append(c)
            ts.append(int.from_bytes(h[4:10], "little"))
            bl.append(int.from_bytes(h[10:12], "little"))
            tr.append((h[3] >> 6) & 1)
            pos += sig
        if stop:
            continue  # the next header read hits EOF
    return dict(payload_offset=np.array(off, dtype=np.int64), n_samples=np.array(ns, dtype=np.int32),
                channel=np.array(ch, dtype=np.int16), timestamp=np.array(ts, dtype=np.int64),
                baseline=np.array(bl, dtype=np.uint16), trunc=np.array(tr, dtype=np.uint8))


def build_records_from_v1725(blobs, names, dt_ns: int):
    """records + wave_pool from V1725 streams (core/processing/records_builder.py:164-209, 798-830):
    timestamp_ps = sample-index timestamp * dt_ns * 1000 (formats/base.py:184-186), baseline = header
    field, flags = trunc, samples reinterpreted as uint16 (:108-112), time = timestamp_ps // 1000; rows
    ordered by lexsort((seq, channel, board, pid, timestamp)) over all files, which equals the reference's
    per-file sort followed by its k-way merge (tie-break part index, then row)."""
    cols = {k: [] for k in ("board", "channel", "timestamp", "baseline", "flags", "n_samples")}
    waves = []
    for blob, name in zip(blobs, names):
        s = v1725_scan(blob)
        raw = np.frombuffer(bytes(blob), dtype=np.uint8)
        k = len(s["channel"])
        cols["board"].append(np.full(k, v1725_board_from_name(name), dtype=np.int16))
        cols["channel"].append(s["channel"])
        cols["timestamp"].append(s["timestamp"] * np.int64(int(dt_ns) * 1000))
        cols["baseline"].append(s["baseline"].astype(np.float64))
        cols["flags"].append(s["trunc"].astype(np.uint32))
        cols["n_samples"].append(s["n_samples"])
        for o, m in zip(s["payload_offset"].tolist(), s["n_samples"].tolist()):
            waves.append(raw[o:o + 2 * m].view(np.uint16))
    c = {k: (np.concatenate(v) if v else np.zeros(0)) for k, v in cols.items()}
    n = len(waves)
    rec = np.zeros(n, dtype=RECORDS_DTYPE)
    if n == 0:
        return rec, np.zeros(0, dtype=np.uint16)
    order = np.lexsort((np.arange(n), c["channel"], c["board"], np.zeros(n, np.int32), c["timestamp"]))
    rec["timestamp"] = c["timestamp"][order]
    rec["board"] = c["board"][order]
    rec["channel"] = c["channel"][order]
    rec["baseline"] = c["baseline"][order]
    rec["baseline_upstream"] = np.nan
    rec["polarity"] = "unknown"
    rec["dt"] = dt_ns
    rec["flags"] = c["flags"][order]
    rec["event_length"] = c["n_samples"][order]
    rec["time"] = rec["timestamp"] // 1000
    lens = rec["event_length"].astype(np.int64)
    rec["wave_offset"] = np.cumsum(lens) - lens
    rec["record_id"] = np.arange(n)
    pool = np.concatenate([waves[i] for i in order.tolist()]) if n else np.zeros(0, dtype=np.uint16)
    return rec, pool.astype(np.uint16, copy=False)


# --------------------------------------------------------------------------------------------
# `hit` = scipy.signal.find_peaks per record.  scipy is an un-vendored dependency of the
# reference (pyproject.toml: scipy>=1.7.0; 1.18.1 in the build container); the algorithm below restates
# scipy/signal/_peak_finding.py (find_peaks) and _peak_finding_utils.pyx (_local_maxima_1d,
# _select_by_peak_distance, _peak_prominences, _peak_widths) and is pinned against live scipy in
# tests/test_oracle_known_answers.py and against the reference plugin's output in tests/golden/hit_golden.npz.
# --------------------------------------------------------------------------------------------
def find_peaks_1d(x, *, height=None, threshold=None, distance=None, prominence=None, width=None, rel_height=0.5):
    """Returns (peaks, left_ips, right_ips, prominences) for the conditions the reference uses
    (lower bounds only).  Plain loops: use on small inputs."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    peaks = []
    i, i_max = 1, n - 1
    while i < i_max:  # _local_maxima_1d: strict rise, plateau midpoint, strict fall
        if x[i - 1] < x[i]:
            ahead = i + 1
            while ahead < i_max and x[ahead] == x[i]:
                ahead += 1
            if x[ahead] < x[i]:
                peaks.append((i + ahead - 1) // 2)
                i = ahead
        i += 1
    peaks = np.array(peaks, dtype=np.int64)
    if height is not None:
        peaks = peaks[x[peaks] >= height]
    if threshold is not None and len(peaks):
        peaks = peaks[np.minimum(x[peaks] - x[peaks - 1], x[peaks] - x[peaks + 1]) >= threshold]
    if distance is not None and len(peaks):
        d = int(np.ceil(distance))
        keep = np.ones(len(peaks), dtype=bool)
        order = np.argsort(x[peaks], kind="stable")
        for j in order[::-1].tolist():  # highest first
            if not keep[j]:
                continue
            k = j - 1
            while k >= 0 and peaks[j] - peaks[k] < d:
                keep[k] = False
                k -= 1
            k = j + 1
            while k < len(peaks) and peaks[k] - peaks[j] < d:
                keep[k] = False
                k += 1
        peaks = peaks[keep]
    proms = np.zeros(len(peaks))
    lb = np.zeros(len(peaks), dtype=np.int64)
    rb = np.zeros(len(peaks), dtype=np.int64)
    for q, p in enumerate(peaks.tolist()):  # _peak_prominences, wlen=None
        i, left_min, lb[q] = p, x[p], p
        while i >= 0 and x[i] <= x[p]:
            if x[i] < left_min:
                left_min, lb[q] = x[i], i
            i -= 1
        i, right_min, rb[q] = p, x[p], p
        while i <= n - 1 and x[i] <= x[p]:
            if x[i] < right_min:
                right_min, rb[q] = x[i], i
            i += 1
        proms[q] = x[p] - max(left_min, right_min)
    if prominence is not None:
        sel = proms >= prominence
        peaks, proms, lb, rb = peaks[sel], proms[sel], lb[sel], rb[sel]
    lips = np.zeros(len(peaks))
    rips = np.zeros(len(peaks))
    for q, p in enumerate(peaks.tolist()):  # _peak_widths
        h = x[p] - proms[q] * rel_height
        i = p
        while lb[q] < i and h < x[i]:
            i -= 1
        lip = float(i)
        if x[i] < h:
            lip += (h - x[i]) / (x[i + 1] - x[i])
        i = p
        while i < rb[q] and h < x[i]:
            i += 1
        rip = float(i)
        if x[i] < h:
            rip -= (h - x[i]) / (x[i - 1] - x[i])
        lips[q], rips[q] = lip, rip
    if width is not None:
        sel = (rips - lips) >= width
        peaks, proms, lips, rips = peaks[sel], proms[sel], lips[sel], rips[sel]
    return peaks, lips, rips, proms


def _peak_height(w, edge_start, edge_end, method, ext):
    """peak_finding.py:567-614: minmax over the rounded (half-to-even) edge window +- ext, or the diff sum."""
    s = max(0, int(np.round(edge_start)))
    e = min(len(w) - 1, int(np.round(edge_end)))
    if method == "diff":
        return float(np.sum(np.diff(-w)[s:e])) if e > s else 0.0
    if method != "minmax":
        raise ValueError(f"unsupported height_method: {method}")
    ext = max(0, int(ext))
    win = w[max(0, s - ext):min(len(w), e + ext)]
    return float(np.max(win) - np.min(win))


def hit_find_peaks(waves, meta, *, source: str, use_derivative=True, height=30.0, distance=2, prominence=0.7, width=4,
                   threshold=None, height_method="minmax", height_window_extension=4):
    """HitFinderPlugin rows (peak_finding.py:213-565).  ``source``: "aos" - rows of st_waveforms /
    filtered_waveforms (``waves`` = list of per-record arrays in their stored dtype, negative pulses:
    detection = -diff(wave) or baseline - wave); "records" - ``waves`` = -RecordsView.signals() in float64
    (detection = +diff or the signal itself).  ``meta`` = structured array with timestamp, board, channel,
    record_id, dt, baseline."""
    from waveformanalysis_b200.dtypes import HIT_DTYPE

    rows = []
    for k, w in enumerate(waves):
        w = np.asarray(w)
        if len(w) == 0:
            continue
        if source == "records":
            det = np.diff(w) if use_derivative else (w - 0.0)
        elif use_derivative:
            det = -np.diff(w)
        else:
            det = np.float64(meta["baseline"][k]) - w
        peaks, lips, rips, _ = find_peaks_1d(det, height=height, threshold=threshold, distance=distance, prominence=prominence, width=width)
        dt_ns = int(meta["dt"][k])
        for p, a, b in zip(peaks.tolist(), lips.tolist(), rips.tolist()):
            rows.append((p, _peak_height(w, a, b, height_method, height_window_extension), 0.0, a, b, dt_ns,
                         int(meta["timestamp"][k] + p * (dt_ns * 1e3)), int(meta["board"][k]), int(meta["channel"][k]),
                         int(meta["record_id"][k])))
    return np.array(rows, dtype=HIT_DTYPE) if rows else np.zeros(0, dtype=HIT_DTYPE)


def stream_find_peaks(waves, meta, *, use_derivative=True, height=30.0, distance=2, prominence=0.7, width=4, threshold=None,
                      height_method="diff", minmax_window_expand=2):
    """signal_peaks_stream rows of one chunk (plugins/builtin/streaming/cpu/signal_peaks.py:234-401): the
    filtered row promoted to float64, detection = -diff(w) or baseline - w, heights by float64 cumsum
    differences ("diff") or max - min over the expanded window ("minmax"), stored float32."""
    from waveformanalysis_b200.dtypes import HIT_DTYPE

    rows = []
    for k, w in enumerate(waves):
        w = np.asarray(w, dtype=np.float64)
        det = -np.diff(w) if use_derivative else (np.float64(meta["baseline"][k]) - w)
        peaks, lips, rips, _ = find_peaks_1d(det, height=height, threshold=threshold, distance=distance, prominence=prominence, width=width)
        if len(peaks) == 0:
            continue
        if height_method == "diff":
            d = -np.diff(w)
            cs = np.concatenate(([0.0], np.cumsum(d, dtype=np.float64)))
            s = np.clip(np.rint(lips).astype(np.int64), 0, len(d))
            e = np.clip(np.rint(rips).astype(np.int64), 0, len(d))
            hts = np.where(e > s, cs[e] - cs[s], 0.0).astype(np.float32)
        elif height_method == "minmax":
            hts = np.zeros(len(peaks), dtype=np.float32)
            for i, (a, b) in enumerate(zip(lips, rips)):
                s = max(0, int(np.round(a)))
                e = min(len(w) - 1, int(np.round(b)))
                win = w[max(0, s - minmax_window_expand):min(len(w), e + minmax_window_expand)]
                hts[i] = np.max(win) - np.min(win)
        else:
            raise ValueError(f"unsupported height_method: {height_method}")
        dt_ns = int(meta["dt"][k])
        for p, a, b, h in zip(peaks.tolist(), lips.tolist(), rips.tolist(), hts.tolist()):
            rows.append((p, h, 0.0, a, b, dt_ns, int(meta["timestamp"][k] + p * (float(dt_ns) * 1e3)), int(meta["board"][k]),
                         int(meta["channel"][k]), int(meta["record_id"][k])))
    return np.array(rows, dtype=HIT_DTYPE) if rows else np.zeros(0, dtype=HIT_DTYPE)


def build_records_ragged(timestamps_ps, boards, channels, sample_blocks, *, dt_ns: int, baseline_window=(0, 40)):
    """build_records for parts of different waveform widths (records_builder.py:212-302 per part, :870-945
    merge): same global order, baseline = mean over the window clipped to the row, ragged wave_pool."""
    ts = np.asarray(timestamps_ps, dtype=np.int64)
    n = len(ts)
    rec = np.zeros(n, dtype=RECORDS_DTYPE)
    rows, base = [], []
    for blk in sample_blocks:
        blk = np.asarray(blk)
        for r in blk:
            rows.append(r.astype(np.int16).view(np.uint16))
            win = r[baseline_window[0]:baseline_window[1]]
            base.append(np.mean(win.astype(np.float64)) if len(win) else np.nan)
    order = np.lexsort((np.arange(n), np.asarray(channels), np.asarray(boards), np.zeros(n, np.int32), ts))
    rec["timestamp"] = ts[order]
    rec["board"] = np.asarray(boards)[order]
    rec["channel"] = np.asarray(channels)[order]
    rec["baseline"] = np.asarray(base)[order]
    rec["baseline_upstream"] = np.nan
    rec["polarity"] = "unknown"
    rec["dt"] = dt_ns
    lens = np.array([len(rows[i]) for i in order.tolist()], dtype=np.int64)
    rec["event_length"] = lens
    rec["wave_offset"] = np.cumsum(lens) - lens
    rec["time"] = rec["timestamp"] // 1000
    rec["record_id"] = np.arange(n)
    pool = np.concatenate([rows[i] for i in order.tolist()]) if n else np.zeros(0, dtype=np.uint16)
    return rec, pool


# --------------------------------------------------------------------------------------------
# the step after the path: df, s1_s2, df_paired
# --------------------------------------------------------------------------------------------


def df_columns(features: np.ndarray, record_id=None, gains: dict | None = None) -> dict:
    """Restates DataFramePlugin.compute (dataframe.py:192-311) as columns: rows in timestamp order (the reference
    sorts with pandas' default quicksort, so only inputs with unique timestamps have a defined order; this
    restatement is stable), `order` = the DataFrame index.  gains {(board, channel): gain} -> area_pe / height_pe =
    float64(value) / gain, NaN where the channel has no entry (:296-309)."""
    n = len(features)
    ts = np.asarray(features["timestamp"], dtype=np.int64)
    order = np.argsort(ts, kind="stable")
    rid = np.arange(n, dtype=np.int64) if record_id is None else np.asarray(record_id, dtype=np.int64)
    out = dict(order=order.astype(np.int64), timestamp=ts[order], record_id=rid[order], area=features["area"][order],
               height=features["height"][order], amp=features["amp"][order], max_abs_diff=features["max_abs_diff"][order],
               board=features["board"][order], channel=features["channel"][order])
    if gains is not None:
        g = np.full(n, np.nan, dtype=np.float64)
        for i in range(n):
            v = gains.get((int(out["board"][i]), int(out["channel"][i])))
            if v is not None:
                g[i] = v
        out["area_pe"] = out["area"].astype(np.float64) / g
        out["height_pe"] = out["height"].astype(np.float64) / g
    return out


def _in_range(values: np.ndarray, bounds) -> np.ndarray:  # s1_s2_classifier.py:54-68
    if bounds is None:
        return np.ones(len(values), dtype=bool)
    lo, hi = bounds
    ok = ~np.isnan(values)
    if lo is not None:
        ok &= ~(values < lo)
    if hi is not None:
        ok &= ~(values > hi)
    return ok


def s1s2_classify(widths: np.ndarray, features: np.ndarray, *, width_unit="ns", s1_width_range=None, s2_width_range=None,
                  s1_area_range=None, s2_area_range=None, s1_height_range=None, s2_height_range=None,
                  conflict_policy="unknown") -> np.ndarray:
    """Restates S1S2ClassifierPlugin.compute (s1_s2_classifier.py:133-228) for the packed dtypes (basic_features has
    no record_id column, so height / area are features[record_id] when 0 <= record_id < len(features), else NaN).
    Ranges are already normalised: None or (lo, hi) with at least one bound."""
    from waveformanalysis_b200.dtypes import S1_S2_CLASSIFIER_DTYPE

    n = len(widths)
    out = np.zeros(n, dtype=S1_S2_CLASSIFIER_DTYPE)
    if n == 0:
        return out
    rid = widths["record_id"].astype(np.int64)
    ok = (rid >= 0) & (rid < len(features))
    height = np.full(n, np.nan, dtype=np.float64)
    area = np.full(n, np.nan, dtype=np.float64)
    height[ok] = features["height"][rid[ok]]
    area[ok] = features["area"][rid[ok]]
    wv = (widths["total_width_samples"] if width_unit == "samples" else widths["total_width"]).astype(np.float64)
    s1_en = any(r is not None for r in (s1_width_range, s1_area_range, s1_height_range))
    s2_en = any(r is not None for r in (s2_width_range, s2_area_range, s2_height_range))
    s1 = _in_range(wv, s1_width_range) & _in_range(area, s1_area_range) & _in_range(height, s1_height_range) & s1_en
    s2 = _in_range(wv, s2_width_range) & _in_range(area, s2_area_range) & _in_range(height, s2_height_range) & s2_en
    label = np.zeros(n, dtype=np.int8)
    label[s1 & ~s2] = 1
    label[s2 & ~s1] = 2
    label[s1 & s2] = {"unknown": 0, "prefer_s1": 1, "prefer_s2": 2}[conflict_policy]
    out["label"] = label
    out["width_ns"] = widths["total_width"]
    out["width_samples"] = widths["total_width_samples"]
    out["height"] = height
    out["area"] = area
    out["timestamp"] = widths["timestamp"]
    out["board"] = widths["board"]
    out["channel"] = widths["channel"]
    out["record_id"] = rid
    out["peak_position"] = widths["peak_position"]
    return out


def pair_events(offsets, member_ts, member_area, member_height, dt_ns, time_window_ns, n_channels: int) -> dict:
    """Restates EventAnalyzer.pair_events (analyzer.py:66-110) on CSR events: keep = dt/ns <= window, delta_t =
    (last - first member timestamp) / 1000.0, i-th member's area / height or NaN."""
    offsets = np.asarray(offsets, dtype=np.int64)
    n = len(offsets) - 1
    keep = np.asarray(dt_ns, dtype=np.float64) <= time_window_ns
    delta = np.zeros(n, dtype=np.float64)
    a = np.full((n, n_channels), np.nan, dtype=np.float32)
    h = np.full((n, n_channels), np.nan, dtype=np.float32)
    for e in range(n):
        s, t = int(offsets[e]), int(offsets[e + 1])
        delta[e] = (int(member_ts[t - 1]) - int(member_ts[s])) / 1000.0
        for i in range(min(n_channels, t - s)):
            a[e, i] = member_area[s + i]
            h[e, i] = member_height[s + i]
    return dict(keep=keep, delta_t=delta, area_ch=a, height_ch=h)
