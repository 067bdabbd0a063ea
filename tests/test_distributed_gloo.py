"""CPU, world_size 2 over gloo: the multi-rank plumbing of the event grouping (counts + padded
column all-gather, rank-order concatenation, local slices of the global event ids).  The grouping
itself is done by the numpy oracle here (no GPU); on GPUs the same code path runs over NCCL with
the device kernel (tools/dist_check.py)."""

import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world_size, port, golden_path, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world_size), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    from oracle import np_oracle as O
    from waveformanalysis_b200 import distributed as D

    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    g = np.load(golden_path)
    hits = g["m0_merged"]
    # shard the hits the way time shards of records would: contiguous blocks in record order
    order = np.argsort(hits["record_id"], kind="stable")
    hits = hits[order]
    b = D.shard_bounds(len(hits), world_size)
    local = hits[b[rank]: b[rank + 1]]
    if rank == 1:
        local = local[:0] if os.environ.get("WFB_EMPTY_RANK") else local
    ev = D.group_hits_distributed(local, 100.0, group_fn=O.group_hit_windows)
    want = O.group_hit_windows(hits if not os.environ.get("WFB_EMPTY_RANK") else hits[: b[1]], 100.0)
    ok = (np.array_equal(ev["t_min"], want["t_min"]) and np.array_equal(ev["n_hits"], want["n_hits"])
          and np.array_equal(ev["local_event_of_hit"], want["event_of_hit"][ev["hit_offset"]: ev["hit_offset"] + len(local)])
          and int(ev["counts"].sum()) == len(want["event_of_hit"]))
    q.put((rank, bool(ok), int(ev["hit_offset"]), len(local)))
    dist.destroy_process_group()


@pytest.mark.parametrize("empty_rank", [False, True])
def test_group_hits_distributed_two_ranks(empty_rank, monkeypatch):
    import torch.multiprocessing as mp

    if empty_rank:
        monkeypatch.setenv("WFB_EMPTY_RANK", "1")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    golden = os.path.join(ROOT, "tests", "golden", "hotpath_golden.npz")
    procs = [ctx.Process(target=_worker, args=(r, 2, port, golden, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
    assert res[0][2] == 0 and res[1][2] == res[0][3]


def test_shard_bounds_cover_everything():
    from waveformanalysis_b200.distributed import shard_bounds

    for n in (0, 1, 7, 1000, 16_000_001):
        for w in (1, 2, 4, 8):
            b = shard_bounds(n, w)
            assert b[0] == 0 and b[-1] == n and np.all(np.diff(b) >= 0) and np.diff(b).max() - np.diff(b).min() <= 1


# ---- time-sharded hit_merge -> hit_grouped with boundary zones (distributed.merge_group_sharded) ----------------------


def _sharded_worker(rank, world_size, port, zone_rows, merge_gap, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world_size), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    from oracle import np_oracle as O
    from waveformanalysis_b200 import distributed as D
    from waveformanalysis_b200.synth import make_raw_run, records_from_raw

    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    raw = make_raw_run(6, 120, 400, seed=77, coincidence_fraction=0.6)
    rec, pool = records_from_raw(raw)
    hits = O.threshold_hits(rec, pool, threshold=12.0)
    b = D.shard_bounds(len(rec), world_size)
    mine = hits[(hits["record_id"] >= b[rank]) & (hits["record_id"] < b[rank + 1])]
    backend = D.HostRows(O.hit_merge, O.group_hit_windows)
    res = D.merge_group_sharded(mine, backend, time_window_ns=100.0, merge_gap_ns=merge_gap, span_ns=400 * 2.0, zone_rows=zone_rows)
    merged = D.gather_rows(res["merged"])
    ev = D.gather_rows(np.asarray(res["event_of_merged"], dtype=np.int64).view([("e", "i8")]))["e"]
    ok, why = True, ""
    if rank == 0:
        _, want_m, want_c = O.hit_merge(hits, merge_gap_ns=merge_gap, max_total_width_ns=10000.0)
        want_ev = O.group_hit_windows(want_m, 100.0, component_rows=want_c, component_hits=hits)
        got_m, got_ev = D.assemble_merged(merged, ev)
        fields = [f for f in want_m.dtype.names if f != "component_offset"]
        for f in fields:
            if not np.array_equal(got_m[f], want_m[f], equal_nan=want_m[f].dtype.kind == "f"):
                ok, why = False, f"merged.{f}"
        if ok and not np.array_equal(got_ev, want_ev["event_of_hit"]):
            ok, why = False, "event ids"
        if ok and int(res["events_per_rank"].sum()) != len(want_ev["t_min"]):
            ok, why = False, "event count"
        if merge_gap > 0 and len(want_m) >= len(hits):
            ok, why = False, "the case merges nothing"
    q.put((rank, ok, why, int(res["gathered_bytes"]), int(res["zone_rows"])))
    dist.destroy_process_group()


@pytest.mark.parametrize("zone_rows,merge_gap", [(8, 50.0), (64, 0.0), (64, 50.0), (1_000_000, 50.0)])
def test_merge_group_sharded_two_ranks(zone_rows, merge_gap):
    """hit_merge(merge_gap_ns = 50) -> hit_grouped over two time shards equals the single-process result row for row:
    clusters and events that would straddle the shard boundary are handed over through the boundary zones."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, zone_rows, merge_gap, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res


def test_plan_cuts_needs_a_gap_wider_than_the_window():
    from waveformanalysis_b200 import distributed as D
    from waveformanalysis_b200.dtypes import THRESHOLD_HIT_DTYPE

    def rows(starts_ns, width_ns=20):
        r = np.zeros(len(starts_ns), dtype=THRESHOLD_HIT_DTYPE)
        r["dt"] = 1
        r["timestamp"] = np.asarray(starts_ns, dtype=np.int64) * 1000
        r["position"] = 0
        r["edge_start"] = 0
        r["edge_end"] = width_ns
        return r

    a = rows([0, 100, 200, 300])
    b = rows([330, 360, 1000])
    cuts, ok = D.plan_cuts([a, b], [a[:0], b[:0]], [True, True], gap_ps=100e3, span_ps=0.0)
    assert ok and cuts[0] == 1000e3  # 320 -> 330 and 350 -> 360 are no gaps; 380 -> 1000 is
    cuts, ok = D.plan_cuts([a, b[:2]], [a[:0], b[:0]], [True, True], gap_ps=100e3, span_ps=0.0)
    assert not ok


@pytest.mark.parametrize("world_size,zone_rows", [(3, 16), (4, 8)])
def test_merge_group_sharded_more_ranks(world_size, zone_rows):
    """The same with three and four time shards (cuts between every pair of neighbours, middle ranks hand rows over on both
    sides, zones of a few rows so that they double several times before a gap is found)."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, world_size, port, zone_rows, 50.0, q)) for r in range(world_size)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
