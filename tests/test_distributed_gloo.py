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
