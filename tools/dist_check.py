#!/usr/bin/env python
"""Multi-GPU parity check of the time-sharded pipeline (run under torchrun, one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/dist_check.py \
        [--records-per-rank 2000000] [--verify 1]

Every rank generates ITS time shard of one synthetic run on its GPU, runs the fused features + hits pass and then
hit_merge (merge_gap_ns = 50) -> hit_grouped (time_window_ns = 100) with ``distributed.merge_group_sharded``: the hit
rows never leave the device, the ranks exchange their boundary zones (one all_gather of 2 x 4096 rows per rank) and
two counts.  With --verify 1 rank 0 additionally processes the WHOLE run on its own GPU and checks that the ranks'
feature rows, hit rows, merged rows (after the channel-major assembly) and event ids equal the single-GPU ones byte
for byte; at sizes that no single GPU holds (--verify 0) the check is structural: counts add up, every rank's windows
lie between its cuts, event ids are contiguous across ranks."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from waveformanalysis_b200 import distributed as D
from waveformanalysis_b200 import engine, ops
from waveformanalysis_b200.dtypes import HIT_MERGED_DTYPE, THRESHOLD_HIT_DTYPE

ap = argparse.ArgumentParser()
ap.add_argument("--records-per-rank", type=int, default=1_000_000)
ap.add_argument("--verify", type=int, default=1)
ap.add_argument("--zone-rows", type=int, default=4096)
args = ap.parse_args()

D.init_process_group("nccl")
rank, ws, local = D.world()
torch.cuda.set_device(local)
n, L, NCH, SEED = args.records_per_rank, 800, 16, 4242
THR, WINDOW, GAP = 15.0, 100.0, 50.0


def fused(run):
    res = run.features_hits(threshold=THR, hit_cap=1024)
    torch.cuda.synchronize()
    total = int(res["total"].item())
    out = {"hits": torch.empty((total + 16) * 60, dtype=torch.uint8, device="cuda")}
    res = run.features_hits(threshold=THR, hit_cap=total + 16, out=out)
    torch.cuda.synchronize()
    run.check()
    return res, total


run = engine.DeviceRun.synth(n, L, NCH, seed=SEED, record_base=rank * n)
res, total = fused(run)
if ws > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
sh = D.merge_group_sharded((res["hits"][: total * 60], total), D.DeviceRows(THRESHOLD_HIT_DTYPE), time_window_ns=WINDOW, merge_gap_ns=GAP,
                           span_ns=L * 2.0, zone_rows=args.zone_rows)
torch.cuda.synchronize()
if ws > 1:
    dist.barrier()
t1 = time.perf_counter()
ncl, nev = sh["n_clusters"], sh["n_events"]
ev = sh["event_of_merged"]
# structural checks on every rank
a0 = sh["detail"]["abs_start"][:ncl]
i = [r for r in range(ws) if sh["clusters_per_rank"][r] > 0]
lo = sh["cuts"][i[i.index(rank) - 1]] if rank in i and i.index(rank) > 0 else -np.inf
hi = sh["cuts"][rank] if rank in i and i.index(rank) + 1 < len(i) else np.inf
if ncl:
    assert float(a0.min().item()) >= lo and float(a0.max().item()) < hi, "a cluster starts outside the rank's cuts"
    assert int(ev.min().item()) == sh["event_base"] and int(ev.max().item()) == sh["event_base"] + nev - 1, "event ids are not contiguous"
owned_n = sh["owned"][1]
tot = torch.tensor([total, owned_n, ncl, nev], dtype=torch.int64, device="cuda")
if ws > 1:
    dist.all_reduce(tot)
assert int(tot[0]) == int(tot[1]), "hits were lost or duplicated by the ownership exchange"
msg = (f"world={ws} records/rank={n} hits={int(tot[0])} clusters={int(tot[2])} events={int(tot[3])} zone_rows={sh['zone_rows']} "
       f"gathered_bytes/rank={sh['gathered_bytes']} merge+group wall={1e3 * (t1 - t0):.1f} ms")

if args.verify:
    feats_l = res["features"][: n * 36].cpu().numpy()
    hits_l = res["hits"][: total * 60].cpu().numpy().view(THRESHOLD_HIT_DTYPE)
    merged_l = sh["merged"][: ncl * 72].cpu().numpy().view(HIT_MERGED_DTYPE)
    ev_l = ev.cpu().numpy().view([("e", "i8")])
    feats = D.gather_rows(feats_l.view([("b", "u1", (36,))]))
    hits = D.gather_rows(hits_l)
    merged = D.gather_rows(merged_l)
    evs = D.gather_rows(ev_l)["e"]
    if rank == 0:
        del run, res, sh
        torch.cuda.empty_cache()
        full = engine.DeviceRun.synth(n * ws, L, NCH, seed=SEED)
        fres, ftotal = fused(full)
        assert np.array_equal(feats.view(np.uint8).reshape(-1), fres["features"][: n * ws * 36].cpu().numpy()), "features differ"
        assert np.array_equal(hits.view(np.uint8).reshape(-1), fres["hits"][: ftotal * 60].cpu().numpy()), "hit rows differ"
        m = ops.hit_merge_device(fres["hits"][: ftotal * 60], ftotal, merge_gap_ns=GAP)
        g = ops.group_rows_device(m["merged"], m["n_clusters"], 72, WINDOW, abs_start=m["abs_start"], abs_end=m["abs_end"])
        want_m = m["merged"][: m["n_clusters"] * 72].cpu().numpy().view(HIT_MERGED_DTYPE)
        want_ev = g["event_of_row"][: m["n_clusters"]].cpu().numpy()
        got_m, got_ev = D.assemble_merged(merged, evs)
        assert len(got_m) == len(want_m), (len(got_m), len(want_m))
        for f in want_m.dtype.names:
            if f == "component_offset":
                continue
            assert np.array_equal(got_m[f], want_m[f], equal_nan=want_m[f].dtype.kind == "f"), f"hit_merged.{f} differs"
        assert np.array_equal(got_ev, want_ev), "event ids differ"
        assert g["n_events"] == int(tot[3])
        merged_over = int((want_m["component_count"] > 1).sum())
        msg += f" | byte-identical to the single-GPU run ({n * ws} records, {merged_over} clusters of several hits)"
if rank == 0:
    print("dist_check OK:", msg, f"backend={dist.get_backend() if ws > 1 else 'single'}")
if ws > 1:
    dist.barrier()
    dist.destroy_process_group()
