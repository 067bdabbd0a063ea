"""GPU: device residency across plugin calls (waveformanalysis_b200/residency.py, plugins/_fused.py).

The first of basic_features / hit_threshold uploads the run once and computes both in one fused pass; the sibling
then returns the rows left for it; wave_pool_filtered keeps its result in HBM for the plugins that read it.  The rows
must be the reference's whatever the order of the calls."""

import numpy as np
import pytest

from conftest import assert_rows_match
from fakes import Ctx

pytestmark = pytest.mark.gpu

FX_BF = ("height", "amp", "max_abs_diff")
FX_HIT = ("height", "width", "rise_time", "fall_time")


@pytest.fixture()
def env():
    from waveformanalysis_b200 import plugins, residency

    residency.release()
    for k in residency.STATS:
        residency.STATS[k] = 0
    yield plugins, residency
    residency.release()


def make_ctx(P, data, config):
    plugins = {"records": P.B200RecordsPlugin(), "wave_pool": P.B200WavePoolPlugin(),
               "basic_features": P.B200BasicFeaturesPlugin(), "hit_threshold": P.B200ThresholdHitPlugin(),
               "wave_pool_filtered": P.B200WavePoolFilteredPlugin(), "waveform_width_integral": P.B200WaveformWidthIntegralPlugin()}
    return Ctx(config, data, plugins=plugins), plugins


@pytest.mark.parametrize("first", ["basic_features", "hit_threshold"])
def test_one_upload_and_one_fused_pass_for_both_plugins(env, golden, first):
    P, residency = env
    cfg = {"wave_source": "records",
           "basic_features": {"channel_config": {"channels": {"0:1": {"fixed_baseline": 8000.5}, "0:3": {"fixed_baseline": 7990.0}}}},
           "hit_threshold": {"threshold": 12.0, "left_extension": 5, "right_extension": 0, "channel_config": {"channels": {"0:2": {"threshold": 40.0}}}}}
    ctx, plugins = make_ctx(P, {"records": golden["records"], "wave_pool": golden["wave_pool"]}, cfg)
    order = [first, "hit_threshold" if first == "basic_features" else "basic_features"]
    out = {name: plugins[name].compute(ctx, "run") for name in order}
    assert residency.STATS["uploads"] == 1 and residency.STATS["row_hits"] == 1, residency.STATS
    assert_rows_match(out["basic_features"], golden["bf_fixed"], what="bf_fixed", float_exact=FX_BF)
    assert_rows_match(out["hit_threshold"], golden["hits_chan"], what="hits_chan", float_exact=FX_HIT)
    # a second request finds the run resident (no upload), the sibling rows were consumed: a fresh fused pass
    again = plugins["hit_threshold"].compute(ctx, "run")
    assert residency.STATS["uploads"] == 1
    assert_rows_match(again, golden["hits_chan"], what="hits_chan again", float_exact=FX_HIT)


def test_changed_config_or_inputs_never_reuse_stale_rows(env, golden):
    P, residency = env
    data = {"records": golden["records"], "wave_pool": golden["wave_pool"]}
    ctx, plugins = make_ctx(P, data, {"wave_source": "records", "hit_threshold": {"threshold": 15.0}})
    plugins["basic_features"].compute(ctx, "run")  # leaves hit rows for threshold 15
    ctx.config["hit_threshold"] = {"threshold": 12.0, "left_extension": 5, "right_extension": 0,
                                   "channel_config": {"channels": {"0:2": {"threshold": 40.0}}}}
    assert_rows_match(plugins["hit_threshold"].compute(ctx, "run"), golden["hits_chan"], what="other config", float_exact=FX_HIT)
    # another pool under the same run id and name: the fingerprint differs, the run is uploaded again
    pool2 = golden["wave_pool"].copy()
    pool2[::3] = pool2[::3] // 2
    from oracle import np_oracle as O

    ctx2, plugins2 = make_ctx(P, {"records": golden["records"], "wave_pool": pool2}, {"wave_source": "records", "hit_threshold": {"threshold": 15.0}})
    got = plugins2["hit_threshold"].compute(ctx2, "run")
    assert_rows_match(got, O.threshold_hits(golden["records"], pool2, threshold=15.0), what="changed pool", float_exact=FX_HIT)
    assert residency.STATS["uploads"] == 2


def test_filtered_pool_stays_resident(env, golden):
    P, residency = env
    rec, pool = golden["filt_records"], golden["filt_pool"]
    ctx, plugins = make_ctx(P, {"records": rec, "wave_pool": pool}, {"wave_source": "records", "use_filtered": True,
                                                                   "hit_threshold": {"threshold": 15.0}})
    sg = plugins["wave_pool_filtered"].compute(ctx, "run")
    assert np.allclose(sg, golden["filt_sg"], rtol=1e-5, atol=1e-3)
    ctx._set_data("run", "wave_pool_filtered", sg)
    uploads = residency.STATS["uploads"]
    # exact against the oracle on the SAME filtered samples (the host copy the plugin handed back); against the
    # reference chain within the float32 noise of scipy's edge fit (2 x 5 edge samples per record, see test_gpu_r2)
    from oracle import np_oracle as O

    bf = plugins["basic_features"].compute(ctx, "run")
    hits = plugins["hit_threshold"].compute(ctx, "run")
    assert_rows_match(bf, O.basic_features(rec, sg), what="filt_bf vs oracle on the same samples")
    assert_rows_match(hits, O.threshold_hits(rec, sg, threshold=15.0), what="filt_hits vs oracle on the same samples")
    assert_rows_match(bf, golden["filt_bf"], what="filt_bf", rtol=1e-4, atol=0.1)
    assert_rows_match(hits, golden["filt_hits"], what="filt_hits", rtol=1e-4, atol=0.1)
    assert residency.STATS["uploads"] == uploads, "the filtered pool was uploaded although it was produced on the device"
    assert residency.STATS["row_hits"] == 1


def test_without_a_sibling_only_the_own_rows_are_computed(env, golden, monkeypatch):
    P, residency = env
    ctx = Ctx({"wave_source": "records"}, {"records": golden["records"], "wave_pool": golden["wave_pool"]})  # no plugin registry
    assert_rows_match(P.B200BasicFeaturesPlugin().compute(ctx, "run"), golden["bf_default"], what="bf", float_exact=FX_BF)
    assert residency.STATS["row_hits"] == 0 and not residency._ROWS
    monkeypatch.setenv("WFB_FUSE_SIBLINGS", "0")
    ctx2, plugins = make_ctx(P, {"records": golden["records"], "wave_pool": golden["wave_pool"]}, {"wave_source": "records", "threshold": 15.0})
    assert_rows_match(plugins["hit_threshold"].compute(ctx2, "run"), golden["hits_thr15"], what="hits", float_exact=FX_HIT)
    assert not residency._ROWS


@pytest.mark.parametrize("chunk_records", [0, 37])
def test_host_pipeline_leaves_the_run_resident(env, chunk_records):
    """wfb_process_host_resident: same rows as the plain host pipeline, and the records + pool it uploaded chunk by
    chunk are a complete DeviceRun afterwards (second pass from HBM gives the same rows again)."""
    from waveformanalysis_b200 import engine
    from waveformanalysis_b200.synth import make_raw_run, records_from_raw

    rec, pool = records_from_raw(make_raw_run(5, 90, 333, seed=5))
    want = engine.process_host(rec, pool, threshold=12.0, chunk_records=chunk_records)
    got = engine.process_host(rec, pool, threshold=12.0, chunk_records=chunk_records, keep_resident=True, pinned_results=True)
    assert np.array_equal(got["features"], want["features"]) and np.array_equal(got["hits"], want["hits"])
    run = got["run"]
    assert run.n == len(rec) and run.lmax == 333 and run.dt_range == (int(rec["dt"].min()), int(rec["dt"].max()))
    assert np.array_equal(run.pool_to_host(), pool)
    again = run.run_to_host(threshold=12.0)
    assert np.array_equal(again["features"], want["features"]) and np.array_equal(again["hits"], want["hits"])
    # records that are not in wave_offset order: the call repeats with per-chunk ranges over every record
    perm = np.random.default_rng(0).permutation(len(rec))
    got = engine.process_host(rec[perm], pool, threshold=12.0, chunk_records=41, keep_resident=True)
    want = engine.process_host(rec[perm], pool, threshold=12.0)
    assert np.array_equal(got["features"], want["features"]) and np.array_equal(got["hits"], want["hits"])
    again = got["run"].run_to_host(threshold=12.0)
    assert np.array_equal(again["hits"], want["hits"]) and np.array_equal(got["run"].pool_to_host(), pool)
    # samples that no record refers to are uploaded as well
    sub = rec[10:-10:3]
    got = engine.process_host(sub, pool, threshold=12.0, chunk_records=29, keep_resident=True)
    assert np.array_equal(got["run"].pool_to_host(), pool)
    assert np.array_equal(got["hits"], engine.process_host(sub, pool, threshold=12.0)["hits"])


def test_pageable_and_memmap_sources_go_through_the_stager(env, tmp_path):
    """Host sources of several MB that are not pinned (ordinary numpy arrays, np.memmap views of a cache file - what a
    Context hands to a plugin, core/context_execution.py:241-251) are copied through the library's ring of pinned pieces
    by a few threads; pinned sources go by DMA directly.  Same rows either way, and engine.upload round-trips."""
    import torch

    from waveformanalysis_b200 import engine
    from waveformanalysis_b200.dtypes import RECORDS_DTYPE

    n, L = 60_000, 800  # 96 MB of samples: three pieces of the staging ring per chunk and more
    dev = engine.DeviceRun.synth(n, L, 16, seed=9, with_rows=True)
    pool_pin = torch.empty(n * L, dtype=torch.int16).pin_memory()
    rows_pin = torch.empty(n * 102, dtype=torch.uint8).pin_memory()
    pool_pin.copy_(dev.pool)
    rows_pin.copy_(dev.records_rows)
    torch.cuda.synchronize()
    rec_p, pool_p = rows_pin.numpy().view(RECORDS_DTYPE), pool_pin.numpy().view(np.uint16)
    want = engine.process_host(rec_p, pool_p, threshold=15.0, chunk_records=25_000)
    rec_a, pool_a = np.array(rec_p), np.array(pool_p)  # pageable copies
    got = engine.process_host(rec_a, pool_a, threshold=15.0, chunk_records=25_000)
    assert got["features"].tobytes() == want["features"].tobytes() and got["hits"].tobytes() == want["hits"].tobytes()
    path = tmp_path / "pool.bin"
    pool_a.tofile(path)
    pool_m = np.memmap(path, dtype=np.uint16, mode="r")
    rec_a.setflags(write=False)
    got = engine.process_host(rec_a, pool_m, threshold=15.0, keep_resident=True, pinned_results=True)
    assert got["features"].tobytes() == want["features"].tobytes() and got["hits"].tobytes() == want["hits"].tobytes()
    assert np.array_equal(got["run"].pool_to_host(), pool_a)
    up = engine.upload(pool_m)
    assert np.array_equal(up.cpu().numpy().view(np.uint16), pool_a)
