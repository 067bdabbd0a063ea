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
