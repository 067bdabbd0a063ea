"""Multi-GPU execution of the hot path: one process per GPU (torchrun), time-chunk sharding of the
time-sorted records, and ONE all-gather of the hit grouping columns for cross-channel event
grouping.

* Records are independent units for baseline / filter / hits / features / widths, so a rank simply
  owns a contiguous slice [lo, hi) of the time-sorted record array - a time shard - and no
  collective is needed on that path (weak or strong scaling with zero data-path communication).
* Event grouping needs every hit's absolute window.  Hits are tiny next to samples (36 B of
  grouping columns per hit against 2*L B of samples per record), so every rank all-gathers the
  columns (counts first, then padded payloads) over NCCL / NVLink and computes the global
  boundary flags and event ids redundantly on its own GPU; the result is identical to the
  single-GPU one because ranks concatenate in record order.  This replaces the halo-and-clip
  scheme of the reference's StreamingPlugin (core/plugins/core/streaming.py:318-324, 592-664)
  with an exact global pass.
"""

from __future__ import annotations

import os

import numpy as np

GROUP_COLUMNS = (("timestamp", np.int64), ("position", np.int64), ("start", np.int32), ("end", np.int32), ("dt", np.int32),
                 ("record_id", np.int64), ("board", np.int16), ("channel", np.int16))


def world():
    """(rank, world_size, local_rank) from the torchrun environment."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_bounds(n: int, world_size: int) -> np.ndarray:
    """Boundaries of equal-count contiguous shards: rank r owns [b[r], b[r+1])."""
    return (np.arange(world_size + 1, dtype=np.int64) * int(n)) // int(world_size)


def shard_slice(records: np.ndarray, pool: np.ndarray, rank: int, world_size: int):
    """The rank's time shard of a time-sorted run: (records[lo:hi], pool view covering them, lo)."""
    b = shard_bounds(len(records), world_size)
    lo, hi = int(b[rank]), int(b[rank + 1])
    return records[lo:hi], pool, lo


def init_process_group(backend: str | None = None):
    import torch
    import torch.distributed as dist

    rank, ws, local = world()
    if ws == 1 or dist.is_initialized():
        return
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if backend == "nccl":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend)


def _device_for_collectives():
    import torch
    import torch.distributed as dist

    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def allgather_columns(columns: dict) -> dict:
    """All-gather variable-length 1-D numpy columns (same length on a rank) in rank order.

    The columns are packed into one row-structured byte buffer, so the exchange is one tiny
    all_gather of the per-rank counts plus ONE all_gather of the payload padded to the maximum
    count.  Returns the concatenated columns plus ``counts`` (per-rank lengths)."""
    import torch
    import torch.distributed as dist

    names = list(columns)
    n_local = len(columns[names[0]]) if names else 0
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        out = {k: np.asarray(v) for k, v in columns.items()}
        out["counts"] = np.array([n_local], dtype=np.int64)
        return out
    ws = dist.get_world_size()
    dev = _device_for_collectives()
    row_dtype = np.dtype([(k, np.asarray(columns[k]).dtype) for k in names])
    rows = np.zeros(n_local, dtype=row_dtype)
    for k in names:
        rows[k] = columns[k]
    cnt = torch.tensor([n_local], dtype=torch.int64, device=dev)
    counts_t = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(ws)]
    dist.all_gather(counts_t, cnt)
    counts = np.array([int(c.item()) for c in counts_t], dtype=np.int64)
    nmax = max(int(counts.max()), 1)
    t = torch.zeros(nmax * row_dtype.itemsize, dtype=torch.uint8, device=dev)
    if n_local:
        t[: n_local * row_dtype.itemsize] = torch.from_numpy(rows.view(np.uint8).reshape(-1)).to(dev)
    parts = [torch.empty_like(t) for _ in range(ws)]
    dist.all_gather(parts, t)
    gathered = np.concatenate([parts[r][: int(counts[r]) * row_dtype.itemsize].cpu().numpy() for r in range(ws)]).view(row_dtype)
    out = {k: np.ascontiguousarray(gathered[k]) for k in names}
    out["counts"] = counts
    return out


def hit_group_columns(hits: np.ndarray) -> dict:
    names = hits.dtype.names
    sn, en = ("sample_start", "sample_end") if "sample_start" in names else ("edge_start", "edge_end")
    return {
        "timestamp": hits["timestamp"].astype(np.int64), "position": hits["position"].astype(np.int64),
        "start": hits[sn].astype(np.int32), "end": hits[en].astype(np.int32), "dt": hits["dt"].astype(np.int32),
        "record_id": hits["record_id"].astype(np.int64), "board": hits["board"].astype(np.int16),
        "channel": hits["channel"].astype(np.int16),
    }


def columns_as_hits(cols: dict) -> np.ndarray:
    """Gathered columns -> a minimal structured array accepted by ops.group_hit_windows."""
    n = len(cols["timestamp"])
    h = np.zeros(n, dtype=[("timestamp", "i8"), ("position", "i8"), ("edge_start", "i4"), ("edge_end", "i4"), ("dt", "i4"),
                           ("record_id", "i8"), ("board", "i2"), ("channel", "i2")])
    h["timestamp"], h["position"] = cols["timestamp"], cols["position"]
    h["edge_start"], h["edge_end"], h["dt"] = cols["start"], cols["end"], cols["dt"]
    h["record_id"], h["board"], h["channel"] = cols["record_id"], cols["board"], cols["channel"]
    return h


def group_hits_distributed(local_hits: np.ndarray, time_window_ns: float, group_fn=None) -> dict:
    """Globally consistent event grouping of per-rank hit rows.

    Every rank gathers all ranks' grouping columns, groups the global set (``group_fn`` defaults to
    the device implementation ``ops.group_hit_windows``) and returns the global event table plus
    ``local_event_of_hit`` (event id of each of its own hits) and ``hit_offset`` (index of its first
    hit in the global order)."""
    if group_fn is None:
        from . import ops

        group_fn = ops.group_hit_windows
    cols = allgather_columns(hit_group_columns(local_hits))
    counts = cols.pop("counts")
    rank = world()[0] if len(counts) > 1 else 0
    ev = group_fn(columns_as_hits(cols), time_window_ns)
    start = int(counts[:rank].sum())
    ev["hit_offset"] = start
    ev["local_event_of_hit"] = ev["event_of_hit"][start : start + int(counts[rank])]
    ev["counts"] = counts
    return ev


def process_shard(records: np.ndarray, pool: np.ndarray, **kw) -> dict:
    """Run the fused pass on this rank's time shard of a time-sorted run.  ``event_index`` and the
    padded matrix width are those of the whole run, so concatenating the ranks' rows in rank order
    reproduces the single-GPU output byte for byte."""
    from . import engine

    rank, ws, _ = world()
    sub, pool, lo = shard_slice(records, pool, rank, ws)
    lmax = int(records["event_length"].max()) if len(records) else 0
    out = engine.process_host(sub, pool, row_base=lo, lmax=lmax, **kw)
    out["row_base"] = lo
    return out


def gather_rows(rows: np.ndarray) -> np.ndarray:
    """Concatenate every rank's structured rows in rank order (all ranks get the result)."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return rows
    ws = dist.get_world_size()
    dev = _device_for_collectives()
    item = rows.dtype.itemsize
    cnt = torch.tensor([len(rows)], dtype=torch.int64, device=dev)
    counts_t = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(ws)]
    dist.all_gather(counts_t, cnt)
    counts = [int(c.item()) for c in counts_t]
    nmax = max(max(counts), 1)
    t = torch.zeros(nmax * item, dtype=torch.uint8, device=dev)
    if len(rows):
        t[: len(rows) * item] = torch.from_numpy(np.ascontiguousarray(rows).view(np.uint8).reshape(-1)).to(dev)
    parts = [torch.empty_like(t) for _ in range(ws)]
    dist.all_gather(parts, t)
    return np.concatenate([parts[r][: counts[r] * item].cpu().numpy() for r in range(ws)]).view(rows.dtype)


# ------------------------------------------------------------------------------------------------
# Time-sharded hit_merge -> hit_grouped that SCALES: boundary zones instead of a global gather
# ------------------------------------------------------------------------------------------------
# Every rank holds the hit rows of its time shard (record order).  Event grouping and same-channel hit merging only
# couple hits that are closer than G = max(time_window_ns, merge_gap_ns).  So the ranks agree on CUT TIMES that lie in
# gaps of the global hit stream wider than G, one per shard boundary; a rank then owns the hits whose window starts
# between its two cuts, and no merge chain or event crosses a cut: hit_merge and the grouping run locally on the
# device, and only two counts per rank are exchanged to number clusters and events globally.
#
# What travels: the first and the last K hit rows of every rank (the BOUNDARY ZONES, K = 4096 rows by default, 60 B
# each) in ONE all_gather, plus one all_gather of 16 bytes per rank - O(ranks * K), independent of the run length.
# The zones serve two purposes: every rank derives the same cut for every boundary from them, and a rank imports the
# neighbour's zone rows that fall on its side of the cut (this is the halo of core/plugins/core/streaming.py:318-324,
# 611-632, with the clip to the main range of :380-445 replaced by exact ownership).
#
# Soundness of a cut C at the boundary between ranks r and r + 1 (rows are in record order, records in time order):
#   * rows of r that are NOT in its tail zone belong to records that start no later than the record of the first
#     tail-zone row, so their windows end before lo = that record's start + the longest record span; C > lo + G.
#   * rows of r + 1 that are NOT in its head zone belong to records that start no earlier than the record of the last
#     head-zone row: their windows start at or after hi = that record's start; C <= hi.
#   * C is the window start of a zone row with every earlier zone window ending more than G before it.
# If no such cut exists the zones are doubled; zones that cover whole shards always work (then the exchange is the
# old global gather).


def _rows_windows(rows: np.ndarray):
    """abs_start / abs_end (event_grouping.py:365-367) and the start of the record of packed hit rows, on the host."""
    sn, en = ("sample_start", "sample_end") if "sample_start" in rows.dtype.names else ("edge_start", "edge_end")
    dt_ps = rows["dt"].astype(np.float64) * 1e3
    ts, pos = rows["timestamp"].astype(np.float64), rows["position"].astype(np.float64)
    a0 = ts + (rows[sn].astype(np.float64) - pos) * dt_ps
    a1 = ts + (rows[en].astype(np.float64) - pos) * dt_ps
    return a0, a1, ts - pos * dt_ps


def plan_cuts(heads: list, tails: list, whole: list, gap_ps: float, span_ps: float):
    """Cut times for the boundaries between consecutive NON-EMPTY ranks.  heads[r] / tails[r]: the first / last zone
    rows of rank r (numpy, packed hit rows, disjoint), whole[r]: the two zones are the whole shard.  Returns (cuts, ok):
    cuts[r] is the cut between rank r and the next non-empty rank (None for empty ranks and the last one); ok is False
    when some boundary has no admissible gap inside the zones (the caller grows the zones)."""
    R = len(heads)
    cuts: list = [None] * R
    live = [r for r in range(R) if len(heads[r]) or len(tails[r])]
    for a, b in zip(live[:-1], live[1:]):
        end_a = np.concatenate([heads[a], tails[a]]) if whole[a] else tails[a]
        start_b = np.concatenate([heads[b], tails[b]]) if whole[b] else heads[b]
        ta0, ta1, trec = _rows_windows(end_a)
        ha0, ha1, hrec = _rows_windows(start_b)
        lo = -np.inf if whole[a] else float(trec[0]) + span_ps + gap_ps
        hi = np.inf if whole[b] else float(hrec[-1])
        s0 = np.concatenate([ta0, ha0])
        s1 = np.concatenate([ta1, ha1])
        o = np.argsort(s0, kind="stable")
        s0, s1 = s0[o], s1[o]
        reach = np.maximum.accumulate(s1)
        ok_i = np.flatnonzero((s0[1:] > reach[:-1] + gap_ps) & (s0[1:] > lo) & (s0[1:] <= hi)) + 1
        if len(ok_i) == 0:
            return cuts, False
        # the admissible gap nearest to the nominal boundary (the first window of the next shard)
        nominal = float(ha0.min())
        cuts[a] = float(s0[ok_i[np.argmin(np.abs(s0[ok_i] - nominal))]])
    return cuts, True


def _allgather_zones(head: np.ndarray, tail: np.ndarray, k: int, whole: bool, n_local: int):
    """ONE all_gather of every rank's zones (2k rows, padded) with a 4 x int64 header (rows in head, rows in tail, zone
    covers the shard, local row count).  Returns per-rank lists."""
    import torch
    import torch.distributed as dist

    item = head.dtype.itemsize
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return [head], [tail], [whole], [n_local], 0
    ws = dist.get_world_size()
    dev = _device_for_collectives()
    buf = np.zeros(32 + 2 * k * item, dtype=np.uint8)
    buf[:32].view(np.int64)[:] = (len(head), len(tail), int(whole), n_local)
    buf[32:32 + len(head) * item] = head.view(np.uint8).reshape(-1)
    buf[32 + k * item:32 + k * item + len(tail) * item] = tail.view(np.uint8).reshape(-1)
    t = torch.from_numpy(buf).to(dev)
    parts = [torch.empty_like(t) for _ in range(ws)]
    dist.all_gather(parts, t)
    heads, tails, wholes, counts = [], [], [], []
    for p in parts:
        b = p.cpu().numpy()
        nh, nt, w, nl = (int(x) for x in b[:32].view(np.int64))
        heads.append(b[32:32 + nh * item].view(head.dtype).copy())
        tails.append(b[32 + k * item:32 + k * item + nt * item].view(head.dtype).copy())
        wholes.append(bool(w))
        counts.append(nl)
    return heads, tails, wholes, counts, int(t.numel()) * ws


def _allgather_i64(vals) -> np.ndarray:
    import torch
    import torch.distributed as dist

    v = np.asarray(vals, dtype=np.int64)
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return v[None, :]
    dev = _device_for_collectives()
    t = torch.from_numpy(v).to(dev)
    parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, t)
    return np.stack([p.cpu().numpy() for p in parts])


class HostRows:
    """Backend for hit rows held as numpy arrays (CPU tests: the compute is the oracle's)."""

    def __init__(self, merge_fn, group_fn):
        self.merge_fn, self.group_fn = merge_fn, group_fn

    def count(self, rows):
        return len(rows)

    def head_tail(self, rows, k):
        n = len(rows)
        kh = min(k, n)
        kt = min(k, n - kh)
        return rows[:kh].copy(), rows[n - kt:].copy()

    def assemble(self, rows, k_head, k_tail, first: np.ndarray, last: np.ndarray):
        return np.concatenate([first, rows[k_head:len(rows) - k_tail], last])

    def merge_and_group(self, owned, merge_gap_ns, max_total_width_ns, time_window_ns):
        clusters, merged, comps = self.merge_fn(owned, merge_gap_ns=merge_gap_ns, max_total_width_ns=max_total_width_ns)
        ev = self.group_fn(merged, time_window_ns, component_rows=comps, component_hits=owned)
        return dict(merged=merged, event_of_merged=np.asarray(ev["event_of_hit"], dtype=np.int64), n_events=len(ev["t_min"]),
                    n_clusters=len(merged))


class DeviceRows:
    """Backend for hit rows that stay on the device (uint8 tensor of packed 60-byte rows + row count)."""

    def __init__(self, dtype):
        self.dtype = dtype

    def count(self, rows):
        return int(rows[1])

    def head_tail(self, rows, k):
        t, n = rows
        item = self.dtype.itemsize
        kh = min(k, n)
        kt = min(k, n - kh)
        head = t[: kh * item].cpu().numpy().view(self.dtype)
        tail = t[(n - kt) * item: n * item].cpu().numpy().view(self.dtype)
        return head, tail

    def assemble(self, rows, k_head, k_tail, first: np.ndarray, last: np.ndarray):
        import torch

        from .engine import upload

        t, n = rows
        item = self.dtype.itemsize
        parts = []
        if len(first):
            parts.append(upload(first))
        parts.append(t[k_head * item:(n - k_tail) * item])
        if len(last):
            parts.append(upload(last))
        out = torch.cat(parts) if len(parts) > 1 else parts[0]
        return out, len(first) + (n - k_head - k_tail) + len(last)

    def merge_and_group(self, owned, merge_gap_ns, max_total_width_ns, time_window_ns):
        from . import ops

        t, n = owned
        m = ops.hit_merge_device(t, n, merge_gap_ns=merge_gap_ns, max_total_width_ns=max_total_width_ns)
        g = ops.group_rows_device(m["merged"], m["n_clusters"], 72, time_window_ns, abs_start=m["abs_start"], abs_end=m["abs_end"])
        return dict(merged=m["merged"], event_of_merged=g["event_of_row"], n_events=g["n_events"], n_clusters=m["n_clusters"], detail=m)


def merge_group_sharded(rows, backend, *, time_window_ns: float, merge_gap_ns: float = 0.0, max_total_width_ns: float = 10000.0,
                        span_ns: float, zone_rows: int = 4096) -> dict:
    """hit_merge -> hit_grouped of a time-sharded run without gathering the hits.

    ``rows``: this rank's packed hit_threshold rows in record order (a numpy array with ``HostRows``, a
    (uint8 device tensor, count) pair with ``DeviceRows``); ``span_ns``: the longest record of the run in ns.
    Returns the rank's ``merged`` rows (clusters whose window starts between the rank's cuts, in the reference's
    channel-major order), ``event_of_merged`` with GLOBAL event ids, ``cluster_base`` / ``event_base`` (the rank's
    offsets in the global numbering), the counts of every rank, the cuts, and ``gathered_bytes`` (what this rank
    received in the collectives)."""
    rank, ws, _ = world()
    gap_ps = max(float(time_window_ns), float(merge_gap_ns), 0.0) * 1e3
    n = backend.count(rows)
    k = max(int(zone_rows), 1)
    gathered = 0
    while True:
        head, tail = backend.head_tail(rows, k)
        whole = len(head) + len(tail) == n
        heads, tails, wholes, counts, nbytes = _allgather_zones(head, tail, k, whole, n)
        gathered += nbytes
        cuts, ok = plan_cuts(heads, tails, wholes, gap_ps, float(span_ns) * 1e3)
        if ok or all(wholes):
            break
        k *= 2
    if not ok:
        raise RuntimeError("no gap wider than the grouping window between two time shards: the run cannot be cut here")
    R = len(heads)
    me = rank if R > 1 else 0
    live = [r for r in range(R) if counts[r] > 0]
    lo_cut, hi_cut = -np.inf, np.inf
    if me in live:
        i = live.index(me)
        lo_cut = cuts[live[i - 1]] if i > 0 else -np.inf
        hi_cut = cuts[me] if i + 1 < len(live) else np.inf
    elif live:  # an empty shard owns nothing
        lo_cut = hi_cut = np.inf

    def inside(r):
        a0 = _rows_windows(r)[0]
        return r[(a0 >= lo_cut) & (a0 < hi_cut)]

    # rows of other ranks' zones that start between my cuts (in global record order: lower ranks first), my own zone rows
    # that do; rows outside the zones never change owner
    lower = [inside(z) for q in range(me) for z in (heads[q], tails[q])]
    upper = [inside(z) for q in range(me + 1, R) for z in (heads[q], tails[q])]
    first = np.concatenate(lower + [inside(head)]) if n or lower else head[:0]
    last = np.concatenate([inside(tail)] + upper) if n or upper else head[:0]
    k_head, k_tail = len(head), len(tail)
    owned = backend.assemble(rows, k_head, k_tail, first, last)
    res = backend.merge_and_group(owned, merge_gap_ns, max_total_width_ns, time_window_ns)
    table = _allgather_i64([res["n_clusters"], res["n_events"]])
    gathered += table.nbytes
    res["cluster_base"] = int(table[:rank, 0].sum())
    res["event_base"] = int(table[:rank, 1].sum())
    res["event_of_merged"] = res["event_of_merged"][: res["n_clusters"]] + res["event_base"]
    res["clusters_per_rank"], res["events_per_rank"] = table[:, 0].copy(), table[:, 1].copy()
    res["cuts"], res["gathered_bytes"], res["zone_rows"], res["owned"] = cuts, gathered, k, owned
    return res


def assemble_merged(merged_all: np.ndarray, event_of_merged_all: np.ndarray):
    """Rank-order concatenation of the ranks' hit_merged rows -> the single-process order: the reference orders the
    merged rows by (board, channel) first and by window start inside a channel (hit_merge.py:125-140), and the ranks'
    pieces of one channel follow each other in time, so a STABLE sort by (board, channel) is all it takes."""
    key = merged_all["board"].astype(np.int64) * 65536 + (merged_all["channel"].astype(np.int64) + 32768)
    o = np.argsort(key, kind="stable")
    return merged_all[o], np.asarray(event_of_merged_all)[o]
