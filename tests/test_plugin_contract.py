"""CPU: the B200 plugins expose the reference's plugin contract (names, versions, dtypes, options,
dynamic dependencies) and the host-side helpers resolve configs exactly like the reference.
Comparisons against the live reference run when /root/reference is present (build container)."""

import os
import sys
from unittest.mock import MagicMock

import numpy as np
import pytest

from fakes import Ctx

REF = os.environ.get("WFB_REFERENCE_ROOT", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_REF = os.path.isdir(os.path.join(REF, "waveform_analysis"))


def _import_ref():
    for mod in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.colors", "matplotlib.figure",
                "matplotlib.axes", "matplotlib.gridspec", "matplotlib.lines", "matplotlib.collections", "matplotlib.cm",
                "matplotlib.ticker", "matplotlib.dates"):
        sys.modules.setdefault(mod, MagicMock())
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import waveform_analysis  # noqa: F401


def test_profile_provides_the_hot_path_names():
    from waveformanalysis_b200 import profiles

    names = [p.provides for p in profiles.b200_default()]
    assert names == ["st_waveforms", "records", "wave_pool", "wave_pool_filtered", "basic_features", "hit_threshold", "hit", "waveform_width",
                     "waveform_width_integral", "hit_merge_clusters", "hit_merged", "hit_merged_components", "hit_grouped",
                     "df", "df_events", "df_paired", "s1_s2"]
    for p in profiles.b200_default():
        assert callable(p.compute) and isinstance(p.options, dict) and p.version


def test_dynamic_dependencies_follow_wave_source():
    from waveformanalysis_b200.plugins import B200BasicFeaturesPlugin, B200ThresholdHitPlugin, B200WaveformWidthPlugin

    p = B200ThresholdHitPlugin()
    assert p.resolve_depends_on(Ctx({"wave_source": "records", "use_filtered": True})) == ["records", "wave_pool_filtered"]
    assert p.resolve_depends_on(Ctx({"wave_source": "records"})) == ["records", "wave_pool"]
    assert p.resolve_depends_on(Ctx({})) == ["st_waveforms"]
    assert B200BasicFeaturesPlugin().resolve_depends_on(Ctx({"use_filtered": True})) == ["filtered_waveforms"]
    assert B200WaveformWidthPlugin().resolve_depends_on(Ctx({"use_filtered": True})) == ["hit", "filtered_waveforms"]
    with pytest.raises(ValueError, match="Invalid wave_source"):
        p.resolve_depends_on(Ctx({"wave_source": "nope"}))


def test_channel_config_layers_and_errors():
    from waveformanalysis_b200.channel_config import per_channel_option, resolve_channel_values

    cfg = {"defaults": {"threshold": 7.0}, "groups": {"g": {"channels": ["0:1", (0, 2)], "config": {"threshold": 8.0}}},
           "channels": {"0:2": {"threshold": 9.0}}}
    assert resolve_channel_values(cfg, "run", 0, 0, {"threshold": 10.0})["threshold"] == 7.0
    assert resolve_channel_values(cfg, "run", 0, 1, {"threshold": 10.0})["threshold"] == 8.0
    assert resolve_channel_values(cfg, "run", 0, 2, {"threshold": 10.0})["threshold"] == 9.0
    per_run = {"run": {"0:0": {"threshold": 25.0}}, "other": {"0:0": {"threshold": 1.0}}}
    assert resolve_channel_values(per_run, "run", 0, 0, {"threshold": 10.0})["threshold"] == 25.0
    with pytest.raises(ValueError, match="Invalid channel key"):
        resolve_channel_values({"run": {"1": {"threshold": 5.0}}}, "run", 0, 1, {})
    got = per_channel_option(cfg, "run", np.array([0, 0, 1]), np.array([1, 2, 1]), "threshold", 10.0)
    assert got == {(0, 1): 8.0, (0, 2): 9.0, (1, 1): 7.0}


def test_structured_rows_as_pool_without_repacking():
    from waveformanalysis_b200.aos import structured_as_records
    from waveformanalysis_b200.dtypes import create_record_dtype

    st = np.zeros(5, dtype=create_record_dtype(37))
    st["wave"] = np.arange(5 * 37).reshape(5, 37) - 40
    st["polarity"] = ["positive", "negative", "unknown", "", "positive"]
    rec, pool, signed = structured_as_records(st, raw_polarity=True)
    assert signed and pool.dtype == np.uint16 and np.shares_memory(pool, st)
    for i in range(5):
        o = int(rec["wave_offset"][i])
        assert np.array_equal(pool[o : o + 37].view(np.int16), st["wave"][i])
    assert rec["polarity"].tolist() == ["rawpos", "unknown", "unknown", "unknown", "rawpos"]


@pytest.mark.skipif(not HAVE_REF, reason="reference checkout not available")
def test_contract_matches_reference_plugins():
    _import_ref()
    from waveform_analysis.core.plugins.builtin.cpu import (basic_features, dataframe, event_analysis, hit_finder, hit_merge, peak_finding, records,
                                                            s1_s2_classifier, waveform_width, waveform_width_integral, waveforms)

    from waveformanalysis_b200 import plugins as P

    pairs = [
        (P.B200BasicFeaturesPlugin, basic_features.BasicFeaturesPlugin),
        (P.B200ThresholdHitPlugin, hit_finder.ThresholdHitPlugin),
        (P.B200HitFinderPlugin, peak_finding.HitFinderPlugin),
        (P.B200WavePoolFilteredPlugin, records.WavePoolFilteredPlugin),
        (P.B200WaveformWidthPlugin, waveform_width.WaveformWidthPlugin),
        (P.B200WaveformWidthIntegralPlugin, waveform_width_integral.WaveformWidthIntegralPlugin),
        (P.B200HitMergeClustersPlugin, hit_merge.HitMergeClustersPlugin),
        (P.B200HitMergePlugin, hit_merge.HitMergePlugin),
        (P.B200HitMergedComponentsPlugin, hit_merge.HitMergedComponentsPlugin),
        (P.B200HitGroupedPlugin, event_analysis.HitGroupedPlugin),
        (P.B200GroupedEventsPlugin, event_analysis.GroupedEventsPlugin),
        (P.B200RecordsPlugin, records.RecordsPlugin),
        (P.B200WavePoolPlugin, records.WavePoolPlugin),
        (P.B200DataFramePlugin, dataframe.DataFramePlugin),
        (P.B200PairedEventsPlugin, event_analysis.PairedEventsPlugin),
        (P.B200S1S2ClassifierPlugin, s1_s2_classifier.S1S2ClassifierPlugin),
        (P.B200WaveformsPlugin, waveforms.WaveformsPlugin),
    ]
    for ours, ref in pairs:
        assert ours.provides == ref.provides
        assert ours.version == ref.version, ours.provides
        assert ours.save_when == ref.save_when, ours.provides
        assert list(ours.depends_on) == list(ref.depends_on), ours.provides
        if ref.output_dtype is not None:
            assert np.dtype(ours.output_dtype) == np.dtype(ref.output_dtype), ours.provides
        assert set(ours.options) == set(ref.options), (ours.provides, set(ours.options) ^ set(ref.options))
        for k, opt in ref.options.items():
            assert ours.options[k].default == opt.default, (ours.provides, k)
            assert ours.options[k].type == opt.type, (ours.provides, k)
            assert getattr(ours.options[k], "track", True) == getattr(opt, "track", True), (ours.provides, k)


@pytest.mark.skipif(not HAVE_REF, reason="reference checkout not available")
def test_channel_config_matches_reference_resolution():
    _import_ref()
    from waveform_analysis.core.hardware.channel import resolve_effective_channel_config

    from waveformanalysis_b200.channel_config import resolve_channel_values

    cfgs = [
        None,
        {"channels": {"0:1": {"threshold": 5.0}}},
        {"0:1": {"threshold": 5.0}, (1, 2): {"threshold": 6.0}},
        {"run": {"defaults": {"threshold": 3.0}, "groups": [{"channels": ["1:2"], "config": {"threshold": 4.0}}]}},
        {"defaults": {"fixed_baseline": 1.0}, "groups": {"a": {"channels": [(0, 0)], "config": {"fixed_baseline": 2.0}}},
         "channels": {"0:0": {"fixed_baseline": 3.0}}},
    ]
    for cfg in cfgs:
        for b, c in ((0, 0), (0, 1), (1, 2)):
            want = resolve_effective_channel_config(context=None, plugin=None, run_id="run", board=b, channel=c,
                                                    base_values={"threshold": 10.0, "fixed_baseline": None}, channel_config=cfg).values
            got = resolve_channel_values(cfg, "run", b, c, {"threshold": 10.0, "fixed_baseline": None})
            assert got == want, (cfg, b, c)


REAL_CONTEXT_SCRIPT = r"""
import sys
from unittest.mock import MagicMock
for mod in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.colors", "matplotlib.figure", "matplotlib.axes",
            "matplotlib.gridspec", "matplotlib.lines", "matplotlib.collections", "matplotlib.cm", "matplotlib.ticker", "matplotlib.dates"):
    sys.modules.setdefault(mod, MagicMock())
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
from waveform_analysis.core.context import Context
from waveform_analysis.core.plugins import profiles as ref_profiles
from waveform_analysis.core.plugins.core.base import Plugin
from waveformanalysis_b200 import plugin_api, profiles
assert plugin_api.HAVE_REFERENCE and plugin_api.Plugin is Plugin
ctx = Context(storage_dir=sys.argv[3])
ctx.register(*ref_profiles.cpu_default())
for p in profiles.b200_default():
    assert isinstance(p, Plugin)
    ctx.register(p, allow_override=True)
assert type(ctx._plugins["basic_features"]).__name__ == "B200BasicFeaturesPlugin"
ctx.set_config({"wave_source": "records"}, plugin_name="basic_features")
lineage = ctx.get_lineage("basic_features")
assert lineage["plugin_class"] == "B200BasicFeaturesPlugin", lineage
assert set(lineage["depends_on"]) == {"records", "wave_pool"}, lineage
ctx.set_config({"wave_source": "records", "use_filtered": True}, plugin_name="hit_threshold")
assert set(ctx.get_lineage("hit_threshold")["depends_on"]) == {"records", "wave_pool_filtered"}
print("REAL_CONTEXT_OK")
"""


@pytest.mark.skipif(not HAVE_REF, reason="reference checkout not available")
def test_registers_in_a_real_context(tmp_path):
    """ctx.register(..., allow_override=True) swaps the provider and changes the lineage class
    (core/context.py:532-621, core/foundation/mixins.py:80-107).  Runs in a fresh interpreter so that
    the B200 plugins pick up the reference's own Plugin / Option base classes."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-c", REAL_CONTEXT_SCRIPT, root, REF, str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert "REAL_CONTEXT_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-3000:]


@pytest.mark.skipif(not HAVE_REF, reason="reference checkout not available")
def test_streaming_plugin_runs_inside_the_reference_framework():
    """B200SignalPeaksStreamPlugin subclasses the reference's streaming plugin and replaces its executor by a two-slot
    device pipeline.  Here (no GPU) the two device hooks are replaced by the oracle, so the test covers the host glue: the
    reference's own chunk iterator, clipping and result chunks around our pipeline must reproduce the reference rows."""
    code = r"""
import os, sys
import numpy as np
sys.path.insert(0, %(root)r)
sys.path.insert(0, os.path.join(%(root)r, "tests"))
sys.path.insert(0, os.path.join(%(root)r, "tests", "golden"))
from make_golden import Ctx, import_reference
import_reference()
from waveform_analysis.core.plugins.builtin.streaming.cpu.signal_peaks import SignalPeaksStreamPlugin
from waveform_analysis.core.processing.dtypes import create_record_dtype
from waveform_analysis.core.plugins.builtin.cpu.filtering import create_filtered_waveform_dtype
from oracle import np_oracle as O
from waveformanalysis_b200 import ops, plugins as P
assert issubclass(P.B200SignalPeaksStreamPlugin, SignalPeaksStreamPlugin)
assert set(P.B200SignalPeaksStreamPlugin.options) == set(SignalPeaksStreamPlugin.options)
assert P.B200SignalPeaksStreamPlugin.version == SignalPeaksStreamPlugin.version

from waveform_analysis.core.processing.chunk import Chunk
# no GPU here: the two device hooks of the pipeline are replaced by the oracle, everything around them is the real thing
def fake_begin(self, chunk, slots, context, run_id, **kw):
    return O.stream_find_peaks(list(chunk.metadata["filtered_waveforms"]["wave"]), chunk.data, **self.peak_options())
def fake_end(self, peaks, chunk, context, run_id):
    if len(peaks) == 0:
        return None
    return Chunk(data=peaks, start=int(peaks["timestamp"].min()), end=int(peaks["timestamp"].max()), run_id=run_id,
                 data_type=self.provides, data_kind=self.output_data_kind, time_field="timestamp")
P.B200SignalPeaksStreamPlugin.begin_chunk = fake_begin
P.B200SignalPeaksStreamPlugin.end_chunk = fake_end
P.B200SignalPeaksStreamPlugin._make_slots = lambda self: None

g = np.load(os.path.join(%(root)r, "tests", "golden", "hit_golden.npz"))
rec, pool, fp = g["records"], g["pool"], g["filtered_pool"]
n, L = len(rec), 800
st = np.zeros(n, dtype=create_record_dtype(L))
for f in ("baseline", "baseline_upstream", "polarity", "timestamp", "record_id", "dt", "event_length", "board", "channel"):
    st[f] = rec[f]
st["wave"] = pool.reshape(n, L).view(np.int16)
stf = np.zeros(n, dtype=create_filtered_waveform_dtype(st.dtype))
for f in st.dtype.names:
    if f != "wave":
        stf[f] = st[f]
stf["wave"] = fp.reshape(n, L)
for tag, cfg in (("stream_default", {"height": 10.0}), ("stream_minmax", {"height": 10.0, "height_method": "minmax", "minmax_window_expand": 3, "width": 2})):
    chunks = list(P.B200SignalPeaksStreamPlugin().compute(Ctx(cfg, {"filtered_waveforms": stf, "st_waveforms": st}), "run"))
    rows = np.concatenate([c.data for c in chunks])
    assert rows.tobytes() == g[tag].tobytes(), tag
    assert np.array_equal(np.array([[c.start, c.end] for c in chunks]), g[tag + "_bounds"])
print("ok")
""" % {"root": ROOT}
    import subprocess

    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "ok" in res.stdout, res.stderr[-2000:]


def test_gain_map_resolution_matches_the_reference_rules():
    """dataframe.py:115-190 + channel.py:571-619: mappings unwrapped, non-positive gains dropped with a warning,
    run blocks and a `channels` sub-tree selected, bad keys raise."""
    from after_cases import GAINS_CONFIG, GAINS_RESOLVED
    from waveformanalysis_b200.plugins.dataframe import B200DataFramePlugin, resolve_gain_values

    with pytest.warns(UserWarning, match="board0:ch3"):
        assert resolve_gain_values(GAINS_CONFIG, "run", True, "df") == GAINS_RESOLVED
    assert resolve_gain_values({"run": {"channels": {"1:2": 4.0}}, "0:0": 9.0}, "run", True, "df") == {(1, 2): 4.0}
    assert resolve_gain_values(GAINS_CONFIG, "run", False, "df") == {}
    with pytest.raises(ValueError, match="Invalid channel key"):
        resolve_gain_values({"nonsense": 1.0}, "run", True, "df")
    plug = B200DataFramePlugin()
    assert plug._resolve_gain_map(Ctx({}), "run", True) == ({}, False)
    assert plug._resolve_gain_map(Ctx({"gain_adc_per_pe": {"0:0": 2.0}}), "run", True) == ({(0, 0): 2.0}, True)

    class RunCfg(Ctx):
        def get_run_config(self, run_id):
            return {"calibration": {"gain_adc_per_pe": {"0:1": 5.0}}}

        def has_explicit_config(self, plugin, name):
            return name in self.config

    assert plug._resolve_gain_map(RunCfg({}), "run", True) == ({(0, 1): 5.0}, True)
    assert plug._resolve_gain_map(RunCfg({"gain_adc_per_pe": {}}), "run", True) == ({}, False)  # explicit empty map wins
    assert plug.resolve_depends_on(Ctx({"wave_source": "records"})) == ["records", "basic_features"]
    assert plug.resolve_depends_on(Ctx({"use_filtered": True})) == ["filtered_waveforms", "basic_features"]


def test_paired_frame_host_logic(monkeypatch):
    """pair_events_frame around a stand-in for the device op (the oracle): row filter, delta_t, float64 widening
    of columns with missing members - against the live reference's DataFrame in after_golden.npz."""
    import pandas as pd

    from after_cases import load_after
    from oracle import np_oracle as O
    from waveformanalysis_b200 import ops
    from waveformanalysis_b200.plugins.dataframe import pair_events_frame

    monkeypatch.setattr(ops, "pair_events", O.pair_events)
    A = load_after()
    off = A["pair_ev_offsets"]
    rag = lambda flat: [flat[off[i]:off[i + 1]] for i in range(len(off) - 1)]  # noqa: E731
    ev = pd.DataFrame({"event_id": np.arange(len(off) - 1), "dt/ns": A["pair_ev_dt_ns"], "n_hits": A["pair_ev_n_hits"],
                       "areas": rag(A["pair_ev_areas"]), "heights": rag(A["pair_ev_heights"]), "timestamps": rag(A["pair_ev_timestamps"])})
    for name in ("a", "b", "c"):
        nch, start = (int(v) for v in A[f"pair_{name}_nch_start"])
        paired = pair_events_frame(ev, nch, start, float(A[f"pair_{name}_tw"]))
        assert np.array_equal(paired.index.to_numpy(), A[f"pair_{name}_index"])
        assert np.array_equal(paired["delta_t"].to_numpy(), A[f"pair_{name}_delta_t"])
        for i in range(nch):
            for kind in ("area", "height"):
                col = f"{kind}_ch{start + i}"
                assert str(paired[col].dtype) == str(A[f"pair_{name}_{col}_dtype"])
                assert np.array_equal(paired[col].to_numpy(), A[f"pair_{name}_{col}"], equal_nan=True)
    assert len(pair_events_frame(ev[:0], 2, 6, 100.0)) == 0


def test_cache_entry_roundtrip(tmp_path):
    from waveformanalysis_b200.cache_format import read_cache_entry, write_cache_entry
    from waveformanalysis_b200.dtypes import BASIC_FEATURES_DTYPE

    rows = np.zeros(17, dtype=BASIC_FEATURES_DTYPE)
    rows["height"] = np.arange(17)
    rows["timestamp"] = np.arange(17) * 1000
    write_cache_entry(str(tmp_path), "run", "run-basic_features-abc", rows, {"lineage": {"plugin": "x"}})
    back = read_cache_entry(str(tmp_path), "run", "run-basic_features-abc")
    assert back.dtype == rows.dtype and np.array_equal(np.asarray(back), rows)
    assert read_cache_entry(str(tmp_path), "run", "missing") is None
    write_cache_entry(str(tmp_path), "run", "run-empty-abc", rows[:0])
    assert read_cache_entry(str(tmp_path), "run", "run-empty-abc") is None


@pytest.mark.skipif(not HAVE_REF, reason="reference checkout not available")
def test_cache_entry_is_loaded_by_the_reference_storage(tmp_path):
    """What write_cache_entry stores, the reference's MemmapStorage finds, loads and reports (memmap.py:528-613,
    689-760); and the other way round."""
    _import_ref()
    from waveform_analysis.core.storage.memmap import MemmapStorage

    from waveformanalysis_b200.cache_format import read_cache_entry, write_cache_entry
    from waveformanalysis_b200.dtypes import S1_S2_CLASSIFIER_DTYPE, THRESHOLD_HIT_DTYPE

    store = MemmapStorage(str(tmp_path))
    rng = np.random.default_rng(3)
    for dt in (THRESHOLD_HIT_DTYPE, S1_S2_CLASSIFIER_DTYPE, np.dtype("f4")):
        rows = np.frombuffer(rng.integers(0, 255, 40 * dt.itemsize, dtype=np.uint8).tobytes(), dtype=dt).copy()
        key = f"run-thing{dt.itemsize}-0123abcd"
        write_cache_entry(str(tmp_path), "run", key, rows, {"lineage": {"a": 1}})
        assert store.exists(key, "run")
        got = store.load_memmap(key, "run")
        assert got.dtype == dt and np.asarray(got).tobytes() == rows.tobytes()
        assert store.get_metadata(key, "run")["lineage"] == {"a": 1}
        key2 = key + "-ref"
        store.save_memmap(key2, rows, run_id="run")
        mine = read_cache_entry(str(tmp_path), "run", key2)
        assert mine.dtype == dt and np.asarray(mine).tobytes() == rows.tobytes()
