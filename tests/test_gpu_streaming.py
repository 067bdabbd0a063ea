"""GPU: the streaming backend (plugins/streaming.py, engine.StreamSlots) - records in time chunks with halo through the
device pipeline must give the rows of the non-streaming plugins byte for byte."""

import numpy as np
import pytest

from fakes import Ctx

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def run_data():
    from waveformanalysis_b200.synth import make_raw_run, records_from_raw

    rec, pool = records_from_raw(make_raw_run(8, 400, 500, seed=41, coincidence_fraction=0.4))
    rec["timestamp"][1500:] += 40_000_000_000_000  # a pause: two time segments
    rec["polarity"][::9] = "positive"
    return rec, pool


@pytest.mark.parametrize("chunk_size,halo_ns", [(1000, 0), (257, 0), (300, 4000), (5000, 0)])
def test_hit_stream_equals_the_plain_plugins(run_data, chunk_size, halo_ns):
    from waveformanalysis_b200 import engine
    from waveformanalysis_b200.plugins import B200HitThresholdStreamPlugin

    rec, pool = run_data
    cfg = {"wave_source": "records", "threshold": 14.0, "left_extension": 1, "right_extension": 2, "height_range": (10, 120)}
    want = engine.process_host(rec, pool, threshold=14.0, left_extension=1, right_extension=2, height_range=(10, 120))
    plugin = B200HitThresholdStreamPlugin()
    plugin.chunk_size = chunk_size
    plugin.required_halo_ns = halo_ns
    chunks = list(plugin.compute(Ctx(cfg, {"records": rec, "wave_pool": pool}), "run"))
    got_h = np.concatenate([c.data for c in chunks])
    got_f = np.concatenate([c.metadata["basic_features"] for c in chunks])
    assert got_h.tobytes() == want["hits"].tobytes()
    assert got_f.tobytes() == want["features"].tobytes()
    n_chunks = sum(-(-m // chunk_size) for m in (1500, len(rec) - 1500))
    assert len(chunks) == n_chunks and plugin.stream_stats["chunks"] == n_chunks
    assert plugin.stream_stats["overlapped_chunks"] == n_chunks - 1
    assert plugin.stream_stats["bytes_uploaded"] >= pool.nbytes + rec.nbytes


def test_hit_stream_channel_rules_and_single_chunk_entry(run_data):
    """Per-channel thresholds / fixed baselines reach every chunk; compute_chunk alone (the protocol's entry point) gives
    the same rows as the pipeline."""
    from waveformanalysis_b200 import engine
    from waveformanalysis_b200.plugins import B200HitThresholdStreamPlugin

    rec, pool = run_data
    cc = {"channels": {"0:3": {"threshold": 25.0}, "0:5": {"fixed_baseline": 8001.5}}}
    cfg = {"wave_source": "records", "threshold": 14.0, "channel_config": cc}
    ctx = Ctx(cfg, {"records": rec, "wave_pool": pool})
    want = engine.process_host(rec, pool, threshold=14.0, thresholds={(0, 3): 25.0}, fixed_baselines={(0, 5): 8001.5})
    plugin = B200HitThresholdStreamPlugin()
    plugin.chunk_size = 700
    chunks = list(plugin.compute(ctx, "run"))
    assert np.concatenate([c.data for c in chunks]).tobytes() == want["hits"].tobytes()
    assert np.concatenate([c.metadata["basic_features"] for c in chunks]).tobytes() == want["features"].tobytes()
    one = B200HitThresholdStreamPlugin()
    one.chunk_size = 700
    parts = [one.compute_chunk(c, ctx, "run") for c in one._get_input_chunks(ctx, "run")]
    assert np.concatenate([p.data for p in parts]).tobytes() == want["hits"].tobytes()


def test_stream_slots_overlap_and_reuse():
    """Slots are reused every second chunk; a chunk staged while another one computes does not disturb it."""
    import torch

    from waveformanalysis_b200 import engine
    from waveformanalysis_b200.synth import make_raw_run, records_from_raw

    rec, pool = records_from_raw(make_raw_run(4, 600, 800, seed=5))
    want = engine.process_host(rec, pool, threshold=12.0)["hits"]
    slots = engine.StreamSlots()
    bounds = [0, 500, 1300, 1301, 2400]
    jobs = []
    for a, b in zip(bounds[:-1], bounds[1:]):
        lo, hi = int(rec["wave_offset"][a]), int(rec["wave_offset"][b - 1] + rec["event_length"][b - 1])
        st = slots.stage(rec[a:b], pool[lo:hi])
        with torch.cuda.stream(slots.compute_stream):
            run = slots.device_run(st, lmax=800, pool_base=lo, row_base=a)
            st.out = run.features_hits(threshold=12.0)
            slots.mark_done(st)
        jobs.append(st)
        if len(jobs) >= 2:  # collect the chunk before: its slot is the next one to be refilled
            prev = jobs[-2]
            slots.finish(prev)
            with torch.cuda.stream(slots.compute_stream):
                total = int(prev.out["total"].item())
                prev.rows_host = prev.out["hits"][: total * 60].cpu().numpy()
    slots.finish(jobs[-1])
    with torch.cuda.stream(slots.compute_stream):
        total = int(jobs[-1].out["total"].item())
        jobs[-1].rows_host = jobs[-1].out["hits"][: total * 60].cpu().numpy()
    assert np.concatenate([j.rows_host for j in jobs]).tobytes() == want.tobytes()
    assert [j.slot for j in jobs] == [0, 1, 0, 1] and slots.chunks == 4


def test_hit_stream_on_the_filtered_pool_with_channel_rules(run_data):
    """use_filtered = True: the stream reads records + wave_pool_filtered (float32 lane-per-record kernel per chunk);
    per-channel thresholds and fixed baselines apply; rows equal the oracle's on the same float32 pool."""
    from oracle import np_oracle as O
    from waveformanalysis_b200.plugins import B200HitThresholdStreamPlugin

    rec, pool = run_data
    rng = np.random.default_rng(3)
    poolf = pool.astype(np.float32) + rng.normal(0, 0.25, len(pool)).astype(np.float32)
    cc = {"channels": {"0:2": {"threshold": 30.0}, "0:6": {"fixed_baseline": 7995.25}}}
    cfg = {"wave_source": "records", "use_filtered": True, "threshold": 13.0, "channel_config": cc, "height_range": (0, None)}
    ctx = Ctx(cfg, {"records": rec, "wave_pool": pool, "wave_pool_filtered": poolf})
    plugin = B200HitThresholdStreamPlugin()
    plugin.chunk_size = 900
    plugin.required_halo_ns = 1000
    chunks = list(plugin.compute(ctx, "run"))
    got_h = np.concatenate([c.data for c in chunks])
    got_f = np.concatenate([c.metadata["basic_features"] for c in chunks])
    want_h = O.threshold_hits(rec, poolf, threshold=13.0, thresholds={(0, 2): 30.0})
    want_f = O.basic_features(rec, poolf, height_range=(0, None), fixed_baseline={(0, 6): 7995.25})
    from conftest import assert_rows_match

    assert_rows_match(got_h, want_h, what="stream hits on the filtered pool", float_exact=("height",))
    assert_rows_match(got_f, want_f, what="stream features on the filtered pool", float_exact=("height", "amp", "max_abs_diff"))
    assert len(want_h) > 2000


def test_abandoned_stream_leaves_nothing_in_flight(run_data):
    """A consumer that stops after the first chunk: the generator's finally waits for the chunks still on the device;
    the next pass on the same plugin object gives the full result."""
    from waveformanalysis_b200 import engine
    from waveformanalysis_b200.plugins import B200HitThresholdStreamPlugin

    rec, pool = run_data
    ctx = Ctx({"wave_source": "records", "threshold": 14.0}, {"records": rec, "wave_pool": pool})
    plugin = B200HitThresholdStreamPlugin()
    plugin.chunk_size = 400
    gen = plugin.compute(ctx, "run")
    first = next(gen)
    gen.close()
    assert len(first.data) > 0 and plugin.stream_stats["chunks"] >= 1
    want = engine.process_host(rec, pool, threshold=14.0)["hits"]
    assert np.concatenate([c.data for c in plugin.compute(ctx, "run")]).tobytes() == want.tobytes()
