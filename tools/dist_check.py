#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py

Every rank processes its time shard of one synthetic run (fused features + hits), the ranks
all-gather the hit grouping columns over NCCL and each computes the global event grouping on its
own GPU.  Rank 0 checks that the rank-order concatenation equals the single-GPU result byte for
byte and that the event ids are the global ones."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from waveformanalysis_b200 import distributed as D
from waveformanalysis_b200 import engine, ops
from waveformanalysis_b200.synth import make_raw_run, records_from_raw

D.init_process_group("nccl")
rank, ws, local = D.world()
torch.cuda.set_device(local)
raw = make_raw_run(16, 4096, 800, seed=2024, coincidence_fraction=0.5)
rec, pool = records_from_raw(raw)
t0 = time.perf_counter()
out = D.process_shard(rec, pool, threshold=15.0, want_counts=True)
feats = D.gather_rows(out["features"])
hits = D.gather_rows(out["hits"])
_, merged_local, _ = ops.hit_merge_default(out["hits"]) if False else (None, None, None)
ev = D.group_hits_distributed(out["hits"], 100.0)
dist.barrier()
t1 = time.perf_counter()
if rank == 0:
    full = engine.process_host(rec, pool, threshold=15.0)
    assert np.array_equal(feats.view(np.uint8), full["features"].view(np.uint8)), "features differ"
    assert np.array_equal(hits.view(np.uint8), full["hits"].view(np.uint8)), "hits differ"
    want = ops.group_hit_windows(full["hits"], 100.0)
    for k in ("t_min", "t_max", "n_hits", "event_of_hit", "members"):
        assert np.array_equal(ev[k], want[k]), k
    print(f"dist_check OK: world={ws} records={len(rec)} hits={len(hits)} events={len(ev['t_min'])} "
          f"counts={ev['counts'].tolist()} wall={t1 - t0:.3f}s backend={dist.get_backend()}")
lo = ev["hit_offset"]
assert np.array_equal(ev["local_event_of_hit"], ev["event_of_hit"][lo: lo + len(out["hits"])])
dist.barrier()
dist.destroy_process_group()
