#!/usr/bin/env python
"""Pinned host <-> device copy bandwidth of this box: the ceiling of the e2e number.  Alone, or on every rank of a
torchrun launch at the same time (all ranks start together behind a barrier): the aggregate shows how much of the
host side of PCIe the ranks of one box share.

    python tools/pcie_probe.py
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_probe.py
"""
import os

import torch

rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    gbs = n / (e0.elapsed_time(e1) / reps) / 1e6
    if world > 1:  # per-rank rates, summed: the box aggregate
        t = torch.tensor([gbs], dtype=torch.float64, device="cuda")
        lo = t.clone()
        dist.all_reduce(t)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        return float(t.item()), float(lo.item())
    return gbs, gbs


def both():
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


for label, fn in (("H2D 1 GiB", lambda: d.copy_(h, non_blocking=True)), ("D2H 1 GiB", lambda: h2.copy_(d2, non_blocking=True)),
                  ("H2D + D2H concurrently (each direction)", both)):
    total, slowest = timed(fn)
    if rank == 0:
        print(f"{label}: {total:.1f} GB/s aggregate over {world} rank(s), slowest rank {slowest:.1f} GB/s")
if world > 1:
    dist.destroy_process_group()
