"""CPU: host logic of the GPU streaming backend (plugins/streaming.py) - chunk iteration over records with time
breaks and halo, the two-slot pipeline order, clipping to the main records.  The two device hooks are replaced by the
numpy oracle here; tests/test_gpu_streaming.py and tests/test_real_context.py run the real thing on the GPU."""

import numpy as np
import pytest

from fakes import Ctx


def _run_stream(records, pool, cfg, **attrs):
    from oracle import np_oracle as O
    from waveformanalysis_b200.plugins.streaming import B200HitThresholdStreamPlugin

    calls = []

    class Fake(B200HitThresholdStreamPlugin):
        def _make_slots(self):
            return None

        def begin_chunk(self, chunk, slots, context, run_id, **kw):
            calls.append(("begin", int(chunk.metadata["row_base"])))
            c = self._run_cfg
            hits = O.threshold_hits(chunk.data, pool, threshold=c["threshold"], left_extension=c["left_extension"], right_extension=c["right_extension"])
            feats = O.basic_features(chunk.data, pool, height_range=c["height_range"], area_range=c["area_range"])
            feats["event_index"] += int(chunk.metadata["row_base"])
            return hits, feats

        def end_chunk(self, job, chunk, context, run_id):
            calls.append(("end", int(chunk.metadata["row_base"])))
            return self.main_rows_chunk(job[0], job[1], chunk, run_id)

    plugin = Fake()
    for k, v in attrs.items():
        setattr(plugin, k, v)
    ctx = Ctx(dict(cfg, wave_source="records"), {"records": records, "wave_pool": pool})
    chunks = list(plugin.compute(ctx, "run"))
    return plugin, chunks, calls


@pytest.fixture(scope="module")
def run_data():
    from waveformanalysis_b200.synth import make_raw_run, records_from_raw

    rec, pool = records_from_raw(make_raw_run(6, 150, 200, seed=31, coincidence_fraction=0.5))
    # two long pauses: three time segments
    rec["timestamp"][300:] += 30_000_000_000_000
    rec["timestamp"][620:] += 50_000_000_000_000
    return rec, pool


@pytest.mark.parametrize("halo_ns", [0, 3000])
def test_stream_chunks_cover_every_record_once(run_data, halo_ns):
    from oracle import np_oracle as O

    rec, pool = run_data
    cfg = {"threshold": 12.0, "height_range": (10, 60)}
    plugin, chunks, calls = _run_stream(rec, pool, cfg, chunk_size=128, required_halo_ns=halo_ns)
    want_h = O.threshold_hits(rec, pool, threshold=12.0)
    want_f = O.basic_features(rec, pool, height_range=(10, 60))
    got_h = np.concatenate([c.data for c in chunks])
    got_f = np.concatenate([c.metadata["basic_features"] for c in chunks])
    assert got_h.tobytes() == want_h.tobytes() and got_f.tobytes() == want_f.tobytes()
    assert sum(c.metadata["n_records"] for c in chunks) == len(rec)
    # segments end at the pauses; a chunk never spans one
    seg = [c.metadata["segment_id"] for c in chunks]
    assert sorted(set(seg)) == [0, 1, 2] and seg == sorted(seg)
    for c in chunks:
        assert c.start == c.metadata["main_start"] and c.end == c.metadata["main_end"] and c.start < c.end
    # pipeline order: chunks k + 1 and k + 2 begin before chunk k ends (three slots)
    order = [kind for kind, _ in calls]
    assert order[:4] == ["begin", "begin", "begin", "end"] and order[-3:] == ["end", "end", "end"]
    assert order.count("begin") == order.count("end") == len(chunks)
    assert [b for kind, b in calls if kind == "end"] == sorted(b for kind, b in calls if kind == "end")  # results in input order


def test_halo_extends_the_input_chunks_inside_their_segment(run_data):
    from waveformanalysis_b200.plugins.streaming import B200HitThresholdStreamPlugin

    rec, pool = run_data
    plugin = B200HitThresholdStreamPlugin()
    plugin.chunk_size = 100
    plugin.required_halo_ns = 5000
    chunks = list(plugin._record_chunks(rec, "run"))
    ts = rec["timestamp"].astype(np.int64)
    end = ts + rec["event_length"].astype(np.int64) * rec["dt"].astype(np.int64) * 1000
    grew = 0
    for c in chunks:
        lo = c.metadata["row_base"]
        m0, m1 = c.metadata["main_rows"]
        hi = lo + len(c.data)
        grew += (m0 > 0) + (lo + m1 < hi)
        # every record of the segment that touches the extended range is in the chunk, nothing from another segment
        seg_rows = np.flatnonzero((end > c.start) & (ts < c.end))
        assert lo <= seg_rows.min() and seg_rows.max() < hi
        assert c.start <= ts[lo:hi].min() and end[lo:hi].max() <= c.end  # what core/processing/chunk.py:130-150 validates
        assert ts[lo:hi].max() - ts[lo:hi].min() < 10_000_000_000_000
    assert grew > len(chunks) // 2


def test_stream_rejects_other_sources_and_bad_dt(run_data):
    from waveformanalysis_b200.plugins.streaming import B200HitThresholdStreamPlugin

    rec, pool = run_data
    bad = rec.copy()
    bad["dt"][5] = 0
    plugin = B200HitThresholdStreamPlugin()
    with pytest.raises(ValueError, match="dt must be positive"):
        list(plugin._record_chunks(bad, "run"))
    assert list(plugin._record_chunks(rec[:0], "run")) == []


def test_records_host_scan_matches_numpy(run_data):
    """wfb_records_host_scan is plain host code in the library (no device): columns and ranges equal the numpy ones."""
    from waveformanalysis_b200 import engine

    rec, _ = run_data
    rec = np.concatenate([rec] * 80)  # enough rows for several threads
    rec["event_length"][7] = 0
    rec["event_length"][11] = -3
    got = engine.records_host_scan(rec, times=True)
    ts = rec["timestamp"].astype(np.int64)
    lens = rec["event_length"].astype(np.int64)
    assert np.array_equal(got["ts"], ts)
    assert np.array_equal(got["end"], ts + np.maximum(lens, 0) * rec["dt"].astype(np.int64) * 1000)
    live = lens > 0
    assert got["lo"] == rec["wave_offset"][live].min() and got["hi"] == (rec["wave_offset"][live] + lens[live]).max()
    assert got["lmax"] == lens.max() and got["dt_min"] == rec["dt"].min()
    empty = engine.records_host_scan(rec[:0])
    assert (empty["lo"], empty["hi"], empty["lmax"]) == (0, 0, 0)
