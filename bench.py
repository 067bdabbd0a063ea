#!/usr/bin/env python
"""bench.py - records/s of the fused records -> baseline -> threshold hits -> basic_features path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): 16 channels x 1M records x 800 int16 samples per GPU
(16 M records, 25.6 GB of samples), synthetic, generated on the device.
A step = one fused pass over all records of the rank.  Weak scaling: every rank owns its own
time shard of 16 M records; there is no data-path collective in this pass.

Printed JSON line (rank 0): `value` = records/s with inputs resident in HBM (CUDA events, max
over ranks); `e2e` = the same pass through the reference-facing plugin call with pinned HOST
buffers (H2D + kernels + D2H inside the timed region); `roofline` for the fused kernel against
the measured HBM copy bandwidth; `cpu_baseline` = the CPU oracle port on a bounded sample.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "records_per_s"
UNIT = "records/s"
N_CHANNELS = 16
N_SAMPLES = 800
RECORDS_PER_GPU = 16 * 1_000_000
THRESHOLD = 15.0
FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--records", type=int, default=RECORDS_PER_GPU, help="records per GPU (default: configs[1])")
    ap.add_argument("--e2e-records", type=int, default=4_000_000, help="records per GPU for the host-buffer e2e leg")
    ap.add_argument("--cpu-sample", type=int, default=16384, help="records per CPU worker for the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="skip the c3 / c4 / c5 legs (BASELINE configs[2..4])")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def workload_config(args, extra=None):
    cfg = {
        "workload": f"fused records->baseline->hits->basic_features, {N_CHANNELS} ch x {args.records // N_CHANNELS} records x {N_SAMPLES} int16 samples per GPU (BASELINE configs[1])",
        "records_per_gpu": args.records,
        "samples_per_record": N_SAMPLES,
        "threshold": THRESHOLD,
        "height_range": [40, 90],
        "area_range": [0, None],
        "l2_policy": "inputs (25.6 GB per GPU) exceed the 126 MB L2; no flush needed",
        "sharding": "time shards, one per rank, no data-path collective",
    }
    if extra:
        cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------------------------
# CPU legs (oracle port)
# ------------------------------------------------------------------------------------------------


def _cpu_worker(task):
    seed, n = task
    import numpy as np  # noqa: F401

    from oracle import np_oracle as O
    from waveformanalysis_b200.synth import make_raw_run, records_from_raw

    raw = make_raw_run(N_CHANNELS, max(1, n // N_CHANNELS), N_SAMPLES, seed=seed)
    rec, pool = records_from_raw(raw)
    t0 = time.perf_counter()
    O.basic_features(rec, pool)
    O.threshold_hits(rec, pool, threshold=THRESHOLD)
    return len(rec), time.perf_counter() - t0


def cpu_port_throughput(n_per_worker: int, workers: int, seed0: int = 100):
    """records/s of the numpy oracle (basic_features + threshold_hits) over `workers` processes.
    Wall time is that of the slowest worker (generation of the synthetic sample is not timed)."""
    import multiprocessing as mp

    tasks = [(seed0 + i, n_per_worker) for i in range(workers)]
    if workers == 1:
        res = [_cpu_worker(tasks[0])]
    else:
        ctx = mp.get_context("spawn")
        with ctx.Pool(workers) as pool:
            res = pool.map(_cpu_worker, tasks)
    n_total = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    return n_total / wall, n_total, wall


# ---- the UNMODIFIED reference (pip --target install under baseline/_ref, tools/install_reference.sh) ----------


def reference_root():
    """Where the reference package can be imported from, or None.  baseline/_ref is git-ignored but travels
    to the GPU box; /root/reference only exists in the build container."""
    for cand in (os.environ.get("WFB_REFERENCE_ROOT"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "waveform_analysis")):
            return cand
    return None


def _ref_import(root):
    from unittest.mock import MagicMock

    # matplotlib is not in the image and the reference imports it unconditionally (core/context.py:34)
    for mod in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.colors", "matplotlib.figure", "matplotlib.axes",
                "matplotlib.gridspec", "matplotlib.lines", "matplotlib.collections", "matplotlib.cm", "matplotlib.ticker", "matplotlib.dates"):
        sys.modules.setdefault(mod, MagicMock())
    if root not in sys.path:
        sys.path.insert(0, root)
    import waveform_analysis  # noqa: F401


def _ref_worker(task):
    """One real reference Context (core/context.py) with profiles.cpu_default(): records + wave_pool of a
    synthetic shard are seeded, basic_features and hit_threshold (records source, threshold 15) are pulled
    with ctx.get_data - the reference's own plugins, scheduler, memmap cache write and profiler."""
    root, seed, n = task
    import contextlib
    import io
    import logging
    import shutil
    import tempfile

    _ref_import(root)
    logging.disable(logging.CRITICAL)
    from waveform_analysis.core.context import Context
    from waveform_analysis.core.plugins import profiles as ref_profiles

    from waveformanalysis_b200.synth import make_raw_run, records_from_raw

    raw = make_raw_run(N_CHANNELS, max(1, n // N_CHANNELS), N_SAMPLES, seed=seed)
    rec, pool = records_from_raw(raw)
    tmp = tempfile.mkdtemp(prefix="wfb_ref_")
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            ctx = Context(storage_dir=tmp)
            ctx.register(*ref_profiles.cpu_default())
            ctx.set_config({"wave_source": "records"}, plugin_name="basic_features")
            ctx.set_config({"wave_source": "records", "threshold": THRESHOLD}, plugin_name="hit_threshold")
            ctx._set_data("run", "records", rec)
            ctx._set_data("run", "wave_pool", pool)
            t0 = time.perf_counter()
            feats = ctx.get_data("run", "basic_features")
            hits = ctx.get_data("run", "hit_threshold")
            wall = time.perf_counter() - t0
        compute = 0.0
        try:  # the reference's own Profiler keys (core/context_execution.py:147)
            for key in ("plugin.basic_features.compute", "plugin.hit_threshold.compute"):
                compute += float(ctx.profiler.durations[key])
        except Exception:
            compute = float("nan")
        return len(rec), wall, compute, len(feats), len(hits)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


class ReferencePool:
    """Worker processes that each run real reference Contexts (one shard per task): the reference's hot path is
    single-threaded Python, so all host cores are used the way its BatchProcessor does - one Context per process."""

    def __init__(self, workers: int):
        import multiprocessing as mp

        self.root = reference_root()
        self.workers = workers
        self.pool = mp.get_context("spawn").Pool(workers) if workers > 1 else None

    def step(self, n_per_worker: int, seed0: int):
        tasks = [(self.root, seed0 + i, n_per_worker) for i in range(self.workers)]
        t0 = time.perf_counter()
        res = self.pool.map(_ref_worker, tasks) if self.pool else [_ref_worker(tasks[0])]
        outer = time.perf_counter() - t0
        n_total = sum(r[0] for r in res)
        wall = max(r[1] for r in res)  # slowest Context: input generation and interpreter start-up are not timed
        compute = max(r[2] for r in res)
        return dict(value=n_total / wall, n_total=n_total, wall=wall, compute=compute, outer=outer,
                    hits_per_record=sum(r[4] for r in res) / max(n_total, 1))

    def close(self):
        if self.pool:
            self.pool.close()
            self.pool.join()


def run_reference(args):
    """--impl reference: the reference's OWN implementation of the path on the host cores: real Context +
    profiles.cpu_default() from baseline/_ref (kind 'reference'); the numpy oracle port only when the
    reference package is absent (kind 'port')."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    vals = []
    if reference_root() is not None:
        kind = "reference"
        n_w = max(N_CHANNELS, args.cpu_sample // 2)
        rp = ReferencePool(cores)
        try:
            for step in range(args.warmup + args.steps):
                r = rp.step(n_w, seed0=1000 + 17 * step)
                if step >= args.warmup:
                    vals.append((r["value"], r["n_total"], r["wall"], r["compute"], r["hits_per_record"]))
        finally:
            rp.close()
        sample = (f"{vals[0][1]} records per step ({cores} processes x {n_w} records, one real Context each): reference plugins "
                  f"basic_features + hit_threshold (records source) through Context.get_data incl. memmap cache write; "
                  f"plugin compute only (Profiler keys): {vals[0][1] / vals[0][3]:.0f} records/s")
        hpr = sum(v[4] for v in vals) / len(vals)
    else:
        kind = "port"
        for step in range(args.warmup + args.steps):
            v, n_total, wall = cpu_port_throughput(args.cpu_sample // 2, cores, seed0=1000 + 17 * step)
            if step >= args.warmup:
                vals.append((v, n_total, wall))
        sample = f"{vals[0][1]} records per step ({cores} processes x {args.cpu_sample // 2} records), numpy oracle basic_features+threshold_hits"
        hpr = None
    value = sum(v[0] for v in vals) / len(vals)
    ms = 1e3 * sum(v[2] for v in vals) / len(vals)
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u16",
        "data": "synthetic",
        "config": workload_config(args, {"hits_per_record": round(hpr, 1)}),  # nominal (one decimal): the same in both arms
        "hits_per_record_measured": hpr,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# GPU legs
# ------------------------------------------------------------------------------------------------


class ClockSampler:
    QUERY = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_launch():
    """dram bytes per launch of the fused kernel from the committed ncu capture, if present."""
    path = os.path.join(ROOT, "profiles", "fused_kernel_traffic.json")
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return None


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from waveformanalysis_b200 import engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    n = args.records
    # every rank holds ITS time shard of one synthetic run (timestamps, record ids and offsets are those of the whole run)
    run = engine.DeviceRun.synth(n, N_SAMPLES, N_CHANNELS, seed=1234 + 1, record_base=rank * n)
    torch.cuda.synchronize()
    # size the hit buffer from one counting pass (hits beyond the capacity are only counted)
    res = run.features_hits(threshold=THRESHOLD, hit_cap=1024)
    torch.cuda.synchronize()
    n_hits = int(res["total"].item())
    cap = n_hits + 1024
    out = {
        "features": torch.empty(n * 36, dtype=torch.uint8, device="cuda"),
        "hits": torch.empty(cap * 60, dtype=torch.uint8, device="cuda"),
        "total": torch.zeros(1, dtype=torch.int64, device="cuda"),
    }
    kw = dict(threshold=THRESHOLD, hit_cap=cap, out=out)

    for _ in range(max(args.warmup, 3)):
        run.features_hits(**kw)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for i in range(args.steps):
        run.features_hits(**kw)
        ev[i + 1].record()
    barrier()
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[args.steps])
    clocks = sampler.stop() if sampler else None
    run.check()
    assert int(out["total"].item()) == n_hits, "hit count changed between passes"
    total_ms = max_over_ranks(total_ms)
    ms_per_step = total_ms / args.steps
    records_all = sum_over_ranks(float(n))
    hits_all = sum_over_ranks(float(n_hits))
    value = records_all / (ms_per_step * 1e-3)

    # ---- roofline of the fused kernel (rank 0's launches; SURVEY.md 8(d): 2L + 72 + 60h bytes per record)
    h = n_hits / n
    bytes_per_record = 2 * N_SAMPLES + 72 + 60 * h
    kernel_ms = sum(step_ms) / len(step_ms)
    achieved = n * bytes_per_record / (kernel_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    traffic = ncu_traffic_per_launch()
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": (traffic or {}).get("dram_bytes_per_launch"),
        "peak_source": peak_src, "kernel": "lpr_kernel<features,hits,u16> (fused_lpr.cu)",
        "bytes_per_record": bytes_per_record, "hits_per_record": h, "kernel_ms": kernel_ms,
        "raw_sample_GBps": n * 2 * N_SAMPLES / (kernel_ms * 1e-3) / 1e9,
    }
    if traffic:
        roofline["traffic_source"] = traffic.get("source")

    # ---- the same kernel without the hit path (records -> baseline -> basic_features only): how far the
    # sample stream itself is from the HBM roofline; algorithmic bytes 2L + 72 per record
    feat_only = None
    if rank == 0:
        kwf = dict(hits=False, out={"features": out["features"]})
        for _ in range(3):
            run.features_hits(**kwf)
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            run.features_hits(**kwf)
        f1.record()
        torch.cuda.synchronize()
        fms = f0.elapsed_time(f1) / args.steps
        fgbs = n * (2 * N_SAMPLES + 72) / (fms * 1e-3) / 1e9
        feat_only = {"kernel": "lpr_kernel<features,-,u16>", "kernel_ms": fms, "records_per_s": n / (fms * 1e-3),
                     "achieved": fgbs, "peak": peak, "unit": "GB/s", "frac": fgbs / peak, "bytes_per_record": 2 * N_SAMPLES + 72}
    if world > 1:
        dist.barrier()

    # ---- secondary legs: BASELINE configs[2..4] (the headline keys above are configs[1])
    c4 = None if args.no_legs else run_c4(args, torch, dist, run, out, n_hits, kernel_ms, barrier, max_over_ranks, sum_over_ranks, world)
    c3 = None if (args.no_legs or rank != 0) else run_c3(args, engine, torch, peak)
    c5 = None if (args.no_legs or rank != 0) else run_c5(args, engine, torch, peak)
    if world > 1:
        dist.barrier()

    # ---- e2e: host buffers through the reference-facing call
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, engine, torch, barrier, max_over_ranks, sum_over_ranks, rank)

    # free the device-resident run before the CPU leg
    del run, out, res
    torch.cuda.empty_cache()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        v, n_total, wall = cpu_port_throughput(args.cpu_sample, cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n_total} records ({cores} processes x {args.cpu_sample}), numpy oracle basic_features+threshold_hits, wall {wall:.2f} s"}
        if reference_root() is not None:  # the reference's own plugins through a real Context; the port stays as a second figure
            rp = ReferencePool(cores)
            try:
                rp.step(N_CHANNELS * 8, seed0=7)  # imports / numba warm-up
                r = rp.step(args.cpu_sample // 2, seed0=100)
            finally:
                rp.close()
            cpu = {"value": r["value"], "unit": UNIT, "cores": cores, "kind": "reference",
                   "sample": (f"{r['n_total']} records ({cores} processes x {args.cpu_sample // 2}, one real Context + profiles.cpu_default() each): "
                              f"basic_features + hit_threshold through Context.get_data, wall {r['wall']:.2f} s"),
                   "compute_only_records_per_s": r["n_total"] / r["compute"] if r["compute"] == r["compute"] else None,
                   "port": {"value": v, "kind": "port", "sample": f"{n_total} records, numpy oracle, wall {wall:.2f} s"}}

    if rank == 0:
        line = {
            "metric": METRIC,
            "value": value,
            "unit": UNIT,
            "n_gpus": world,
            "steps": args.steps,
            "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "u16",
            "data": "synthetic",
            "config": workload_config(args, {"hits_per_record": round(hits_all / records_all, 1)}),  # nominal; roofline carries the exact value
            "raw_sample_GBps": records_all * 2 * N_SAMPLES / (ms_per_step * 1e-3) / 1e9,
            "roofline": roofline,
            "roofline_features_only": feat_only,
            "c3": c3,
            "c4": c4,
            "c5": c5,
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": args.steps * 1,
            "clocks": clocks,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()



def _timed(torch, fn, reps=3):
    """Average device time of fn() in ms (CUDA events on the current stream).  Two untimed calls first, the result held
    the way the timed loop holds it: fn() allocates its output, and while one result is alive the next call needs a second
    block - a cudaMalloc of several GB inside the timed region otherwise."""
    r = fn()
    r = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r


def run_c4(args, torch, dist, run, out, n_hits, fused_ms, barrier, max_over_ranks, sum_over_ranks, world):
    """BASELINE configs[3]: the full pipeline on time shards - the fused pass above, then hit_merge (merge_gap_ns 50) ->
    hit_grouped (time_window_ns 100) across the shards with distributed.merge_group_sharded: the hit rows stay on the
    device, the ranks exchange their boundary zones (ONE all_gather) and two counts (one all_gather of 16 B)."""
    from waveformanalysis_b200 import distributed as D
    from waveformanalysis_b200.dtypes import THRESHOLD_HIT_DTYPE

    rows = (out["hits"][: n_hits * 60], n_hits)
    be = D.DeviceRows(THRESHOLD_HIT_DTYPE)

    def go():
        return D.merge_group_sharded(rows, be, time_window_ns=100.0, merge_gap_ns=50.0, span_ns=N_SAMPLES * 2.0)

    go()  # allocator warm-up
    barrier()
    t0 = time.perf_counter()
    sh = go()
    torch.cuda.synchronize()
    barrier()
    mg_ms = max_over_ranks(1e3 * (time.perf_counter() - t0))
    recs = sum_over_ranks(float(run.n))
    res = {
        "workload": "time shards: fused records->hits/features, hit_merge(merge_gap_ns=50) -> hit_grouped(time_window_ns=100) across shards",
        "records_per_s": recs / ((fused_ms + mg_ms) * 1e-3),
        "fused_ms": fused_ms, "merge_group_ms": mg_ms,
        "hits": sum_over_ranks(float(n_hits)), "clusters": float(sh["clusters_per_rank"].sum()), "events": float(sh["events_per_rank"].sum()),
        "collectives": ("all_gather of the boundary zones (first / last %d hit rows of every rank, 60 B each) + all_gather of 2 x int64 "
                        "(cluster and event counts); no gather of the hits" % sh["zone_rows"]),
        "gathered_bytes_per_rank": int(sh["gathered_bytes"]), "timing": "wall clock between barriers, max over ranks (host planning of the cuts included)",
    }
    del sh
    return res


def run_c3(args, engine, torch, peak):
    """BASELINE configs[2]: wave_pool_filtered (SG and Butterworth) and hit -> waveform_width on a 64-channel V1725-like
    run (dt = 4 ns, positive pulses), device resident, rank 0."""
    import ctypes as C

    import numpy as np

    from waveformanalysis_b200 import _lib, ops
    from waveformanalysis_b200.dtypes import HIT_DTYPE

    n, L = 2_000_000, N_SAMPLES
    run = engine.DeviceRun.synth(n, L, 64, dt_ns=4, seed=303)
    # positive pulses: mirror the samples and the baselines around the 14-bit mid-scale, mark the records 'positive'
    run.pool.copy_(16383 - run.pool)
    meta = run.meta.view(torch.uint8).view(-1, 48)
    base = meta[:, 8:16].contiguous().view(torch.float64).view(-1)
    meta[:, 8:16] = (16383.0 - base).view(torch.uint8).view(-1, 8)
    meta[:, 36] = 1  # WFB_POL_POSITIVE
    torch.cuda.synchronize()
    sos = ops.butter_bandpass_sos(4, 0.01, 0.1, 0.25)
    sg_ms, d_sg = _timed(torch, lambda: ops.filter_run_device(run, {"filter_type": "SG", "sg_window_size": 11, "sg_poly_order": 2}))
    bw_ms, _ = _timed(torch, lambda: ops.filter_run_device(run, {"filter_type": "BW", "sos": sos}), reps=2)
    frun = engine.DeviceRun(run.meta, d_sg, n, 1, L)
    lib = _lib.load()
    p = _lib.PeakParams(wave_kind=_lib.WAVE_REC_F32, use_derivative=1, height=8.0, prominence=0.7, width=2.0, threshold=0.0, has_threshold=0,
                        distance=2, height_method=0, height_window_extension=4, lmax=L, level_f32=0)
    ws = torch.empty(lib.wfb_find_peaks_workspace_bytes(n), dtype=torch.uint8, device="cuda")
    total = torch.zeros(1, dtype=torch.int64, device="cuda")
    cap = 8 * n
    rows = torch.empty(cap * HIT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")

    def peaks():
        _lib.check(lib.wfb_find_peaks(engine._ptr(frun.pool), frun.pool_len, engine._ptr(frun.meta), n, C.byref(p), engine._ptr(rows), cap,
                                      C.c_void_p(0), engine._ptr(total), engine._ptr(ws), ws.numel(), engine._stream()), "wfb_find_peaks")
        return int(total.item())

    hit_ms, nh = _timed(torch, peaks, reps=2)
    hits = rows[: nh * HIT_DTYPE.itemsize].view(-1, HIT_DTYPE.itemsize)
    pos = hits[:, 0:8].contiguous().view(torch.int64).view(-1)
    ts = hits[:, 28:36].contiguous().view(torch.int64).view(-1)
    bc = hits[:, 36:40].contiguous().view(torch.int16).view(-1, 2)
    rid = hits[:, 40:48].contiguous().view(torch.int64).view(-1)
    wp = _lib.WidthParams(0.1, 0.9, 0.9, 0.1, 0.25, 1, 1)
    wout = torch.empty(max(nh, 1) * 56, dtype=torch.uint8, device="cuda")
    valid = torch.empty(max(nh, 1), dtype=torch.uint8, device="cuda")
    board, chan = bc[:, 0].contiguous(), bc[:, 1].contiguous()

    def widths():
        _lib.check(lib.wfb_waveform_width(engine._ptr(d_sg), n, L, L, engine._ptr(rid), engine._ptr(pos), engine._ptr(ts), engine._ptr(board),
                                          engine._ptr(chan), engine._ptr(rid), nh, C.byref(wp), engine._ptr(wout), engine._ptr(valid),
                                          engine._stream()), "wfb_waveform_width")
        return int(valid[:nh].sum().item())

    w_ms, nw = _timed(torch, widths)
    # hit_threshold + basic_features on the filtered (float32) pool: the lane-per-record kernel's float32 variant
    fres = frun.features_hits(threshold=THRESHOLD, hit_cap=1024)
    torch.cuda.synchronize()
    fnh = int(fres["total"].item())
    fout = {"features": torch.empty(n * 36, dtype=torch.uint8, device="cuda"), "hits": torch.empty((fnh + 16) * 60, dtype=torch.uint8, device="cuda"),
            "total": torch.zeros(1, dtype=torch.int64, device="cuda")}
    f32_ms, _ = _timed(torch, lambda: frun.features_hits(threshold=THRESHOLD, hit_cap=fnh + 16, out=fout))
    f32_bpr = 4 * L + 72 + 60 * fnh / n
    gb = n * L * 6 / 1e9
    return {
        "fused_f32": {"ms": f32_ms, "records_per_s": n / (f32_ms * 1e-3), "hits_per_record": fnh / n, "bytes_per_record": f32_bpr,
                      "frac_hbm": n * f32_bpr / (f32_ms * 1e-3) / 1e9 / peak,
                      "call": "wfb_features_hits on the SG-filtered float32 pool (threshold hits + basic_features, lane-per-record float32 kernel)"},
        "workload": f"64 ch V1725-like (dt 4 ns, positive pulses), {n} records x {L} samples, device resident: wave_pool_filtered, hit on the filtered pool, waveform_width",
        "sg": {"ms": sg_ms, "GBps": gb / (sg_ms * 1e-3), "frac_hbm": gb / (sg_ms * 1e-3) / peak, "bytes_per_record": 6 * L},
        "bw": {"ms": bw_ms, "GBps": gb / (bw_ms * 1e-3), "frac_hbm": gb / (bw_ms * 1e-3) / peak, "bytes_per_record": 6 * L,
               "note": "sosfiltfilt order 4 band-pass = 4 sections forward + backward in float64 per sample: FP64-issue bound, not HBM bound"},
        "hit": {"ms": hit_ms, "records_per_s": n / (hit_ms * 1e-3), "peaks": nh},
        "waveform_width": {"ms": w_ms, "hits_per_s": nh / (w_ms * 1e-3) if nh else None, "rows": nw},
        "records_per_s_sg_hit_width": n / ((sg_ms + hit_ms + w_ms) * 1e-3),
    }


def run_c5(args, engine, torch, peak):
    """BASELINE configs[4]: record-length sweep of the fused pass (device resident, rank 0; the channel count only
    changes the metadata, so it is fixed at 16 and stated).  ~0.27 G samples per point."""
    pts = []
    for L in (256, 512, 1024, 2048, 4096, 8192):
        n = max(4096, (268_435_456 // L) // 128 * 128)
        r = engine.DeviceRun.synth(n, L, 16, seed=500 + L)
        res = r.features_hits(threshold=THRESHOLD, hit_cap=1024)
        torch.cuda.synchronize()
        nh = int(res["total"].item())
        o = {"features": torch.empty(n * 36, dtype=torch.uint8, device="cuda"), "hits": torch.empty((nh + 16) * 60, dtype=torch.uint8, device="cuda"),
             "total": torch.zeros(1, dtype=torch.int64, device="cuda")}
        ms, _ = _timed(torch, lambda: r.features_hits(threshold=THRESHOLD, hit_cap=nh + 16, out=o), reps=5)
        bpr = 2 * L + 72 + 60 * nh / n
        pts.append({"L": L, "records": n, "hits_per_record": nh / n, "ms": ms, "records_per_s": n / (ms * 1e-3),
                    "raw_sample_GBps": n * 2 * L / (ms * 1e-3) / 1e9, "frac_hbm": n * bpr / (ms * 1e-3) / 1e9 / peak})
        del r, o, res
    torch.cuda.empty_cache()
    return {"workload": "fused records->hits->basic_features, record length sweep, 16 channels, device resident (inputs 0.5 GB per point > L2)",
            "points": pts}


class PluginContext:
    """The part of the reference Context a plugin's compute() uses (core/context.py get_config / get_data and the
    plugin registry), holding host arrays: what ContextExecutionDomain.execute_plugin_compute hands to plugin.compute
    (core/context_execution.py:140-183) without the cache write to disk."""

    def __init__(self, config, plugins):
        self.config = config
        self._results = {}
        self._plugins = plugins

    def get_config(self, plugin, name):
        p = plugin.provides
        if p in self.config and isinstance(self.config[p], dict) and name in self.config[p]:
            return self.config[p][name]
        if name in self.config:
            return self.config[name]
        return plugin.options[name].default if name in plugin.options else None

    def get_data(self, run_id, name):
        return self._results.get((run_id, name))


def run_e2e(args, engine, torch, barrier, max_over_ranks, sum_over_ranks, rank):
    """The same pass end to end through the reference-facing boundary: B200BasicFeaturesPlugin.compute and
    B200ThresholdHitPlugin.compute on HOST records + wave_pool (pinned), host feature / hit rows out, every step.
    The first plugin uploads the run once and computes both outputs in one fused pass; the second returns the rows
    left for it (plugins/_fused.py).  Every step uses a new run id, so every step pays the upload."""
    import numpy as np

    from waveformanalysis_b200 import residency
    from waveformanalysis_b200.dtypes import RECORDS_DTYPE
    from waveformanalysis_b200.plugins import B200BasicFeaturesPlugin, B200RecordsPlugin, B200ThresholdHitPlugin, B200WavePoolPlugin

    n = min(args.e2e_records, args.records)
    dev = engine.DeviceRun.synth(n, N_SAMPLES, N_CHANNELS, seed=77 + rank, with_rows=True)
    pool_pin = torch.empty(n * N_SAMPLES, dtype=torch.int16).pin_memory()
    rows_pin = torch.empty(n * 102, dtype=torch.uint8).pin_memory()
    pool_pin.copy_(dev.pool)
    rows_pin.copy_(dev.records_rows)
    torch.cuda.synchronize()
    del dev
    torch.cuda.empty_cache()
    records = rows_pin.numpy().view(RECORDS_DTYPE)
    pool = pool_pin.numpy().view(np.uint16)
    plugins = {"records": B200RecordsPlugin(), "wave_pool": B200WavePoolPlugin(), "basic_features": B200BasicFeaturesPlugin(),
               "hit_threshold": B200ThresholdHitPlugin()}
    ctx = PluginContext({"wave_source": "records", "hit_threshold": {"threshold": THRESHOLD}, "hit_threshold_stream": {"threshold": THRESHOLD}}, plugins)

    def step(k):
        run_id = f"e2e_{k}"
        ctx._results[(run_id, "records")] = records
        ctx._results[(run_id, "wave_pool")] = pool
        feats = plugins["basic_features"].compute(ctx, run_id)
        hits = plugins["hit_threshold"].compute(ctx, run_id)
        residency.release(run_id)
        ctx._results.clear()
        return feats, hits

    # two untimed steps: the result arrays come from torch's caching pinned allocator, whose blocks are allocated once
    # and recycled when the caller drops the previous step's arrays (a Context replaces them by memmap views at once)
    for k in (-2, -1):
        feats, hits = step(k)
        n_hits0, n_feats0 = len(hits), len(feats)
        del feats, hits
    steps = max(2, min(args.steps, 5))
    uploads0, rowhits0 = residency.STATS["uploads"], residency.STATS["row_hits"]
    barrier()
    t0 = time.perf_counter()
    for k in range(steps):
        feats, hits = step(k)
        ok = len(hits) == n_hits0 and len(feats) == n_feats0
        del feats, hits
    barrier()
    wall = max_over_ranks(time.perf_counter() - t0)
    assert ok and n_feats0 == n
    assert residency.STATS["uploads"] - uploads0 == steps and residency.STATS["row_hits"] - rowhits0 == steps, residency.STATS
    n_all = sum_over_ranks(float(n))
    # the same host arrays through the streaming backend (plugins/streaming.py): time chunks of 262144 records, the upload
    # of chunk k + 1 overlapping the fused pass of chunk k, hit + feature rows of every chunk back on the host
    from waveformanalysis_b200.plugins import B200HitThresholdStreamPlugin

    stream = B200HitThresholdStreamPlugin()
    ctx._results[("stream", "records")] = records
    ctx._results[("stream", "wave_pool")] = pool

    def stream_pass():
        rows = 0
        for chunk in stream.compute(ctx, "stream"):
            rows += len(chunk.data)
        return rows

    stream_pass()
    walls = []
    for _ in range(3):  # median of three passes: one pass in five runs into a host-side hiccup (allocator / GC) of ~0.2 s
        barrier()
        t0 = time.perf_counter()
        rows = stream_pass()
        barrier()
        walls.append(max_over_ranks(time.perf_counter() - t0))
        assert rows == n_hits0, (rows, n_hits0)
    stream_wall = sorted(walls)[1]
    ctx._results.clear()
    # the same plugin calls on PAGEABLE host arrays (what a Context hands over when the data was just computed; cached data
    # comes as np.memmap views): the driver stages the upload, so the link runs below its pinned rate
    n_pg = min(n, 1_000_000)
    rec_pg = np.array(records[:n_pg])
    pool_pg = np.array(pool[: n_pg * N_SAMPLES])

    def pageable_step(k):
        run_id = f"pg_{k}"
        ctx._results[(run_id, "records")] = rec_pg
        ctx._results[(run_id, "wave_pool")] = pool_pg
        f = plugins["basic_features"].compute(ctx, run_id)
        h = plugins["hit_threshold"].compute(ctx, run_id)
        residency.release(run_id)
        ctx._results.clear()
        return len(f) + len(h)

    pageable_step(-1)
    barrier()
    t0 = time.perf_counter()
    for k in range(2):
        pageable_step(k)
    barrier()
    pageable_wall = max_over_ranks(time.perf_counter() - t0)
    pageable_leg = {"value": sum_over_ranks(float(n_pg)) * 2 / pageable_wall, "unit": UNIT, "records_per_gpu": n_pg,
                    "note": "same plugin calls, inputs in ordinary (pageable) numpy arrays"}
    stream_leg = {"value": n_all / stream_wall, "unit": UNIT, "chunk_records": int(stream.chunk_size), "chunks": stream.stream_stats["chunks"],
                  "overlapped_chunks": stream.stream_stats["overlapped_chunks"], "pass_walls_s": [round(w, 4) for w in walls],
                  "call": "B200HitThresholdStreamPlugin.compute (hit_threshold_stream): three-slot device pipeline, rows of every chunk to the host"}
    return {
        "stream": stream_leg,
        "pageable": pageable_leg,
        "value": n_all * steps / wall,
        "unit": UNIT,
        "h2d_bytes_per_step": int(n * (2 * N_SAMPLES + 102)),
        "d2h_bytes_per_step": int(n * 36 + n_hits0 * 60),
        "records_per_gpu": n,
        "steps": steps,
        "pool_uploads_per_step": 1,
        "call": ("B200BasicFeaturesPlugin.compute + B200ThresholdHitPlugin.compute (records source) on pinned host records + wave_pool: "
                 "one upload, one fused pass, host feature / hit rows out"),
    }


_RESULT_FD = None


def emit(line: dict) -> None:
    """The ONE JSON line of the contract, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    global _RESULT_FD
    args = parse_args()
    # library chatter (the NCCL version banner, warnings printed by child processes) must not land on stdout:
    # keep the original stdout for the result line only and point fd 1 at stderr for everything else
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
