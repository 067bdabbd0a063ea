#!/usr/bin/env python
"""Run the records route of the hot path through a REAL reference ``Context`` and dump every result.

    python tests/real_context_run.py cpu|b200 OUT.npz WORKDIR

``cpu``: the reference's own plugins (``profiles.cpu_default()``).  ``b200``: the same Context with
``ctx.register(*b200_default(), allow_override=True)`` on top (core/context.py:532-621), i.e. exactly the
binding INTEGRATION.md shows.  Inputs: two synthetic V1725 ``.bin`` files written to WORKDIR and handed to
the Context as ``raw_files``; everything downstream (st_waveforms, records, wave_pool, wave_pool_filtered, basic_features,
hit_threshold, the three hit-merge outputs with merge_gap_ns = 50, hit_grouped) is pulled with
``ctx.get_data`` (core/context_execution.py:140-183), so dependency resolution, the memmap cache write and
the memmap views handed to downstream plugins are the reference's.  Used by tests/test_real_context.py
(run in fresh interpreters so that the B200 plugins subclass the reference's own Plugin base).
"""

from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)


def main():
    mode, out_path, work = sys.argv[1], sys.argv[2], sys.argv[3]
    scenario = sys.argv[4] if len(sys.argv) > 4 else "records"
    import numpy as np
    from refctx import import_reference

    import_reference()
    from waveform_analysis.core.context import Context
    from waveform_analysis.core.plugins import profiles as ref_profiles

    from waveformanalysis_b200.synth import make_v1725_blob

    os.makedirs(work, exist_ok=True)
    paths = []
    for name, kw in (("run_b0_seg0.bin", dict(n_events=260, n_channels=16, seed=11, lengths=(300, 300), tie_every=9)),
                     ("run_b1_seg0.bin", dict(n_events=200, n_channels=8, seed=12, lengths=(300, 300), tie_every=0, t0=23))):
        p = os.path.join(work, name)
        with open(p, "wb") as f:
            f.write(make_v1725_blob(**kw))
        paths.append(p)

    if scenario in ("csv", "mixed"):
        # VX2730 CSV files (utils/formats/vx2730.py:70-107: ';' separated, two header rows in the first file of a channel,
        # board;channel;timestamp_ps;4 unused columns;samples...), one file per channel, ties in time across channels
        rng = np.random.default_rng(5)
        paths = []
        n_ev, L = 160, 120
        t = np.cumsum(rng.integers(200, 900, n_ev)) * 2000
        for c in range(4):
            pth = os.path.join(work, f"wave_CH{c}_0.CSV")
            tc = t + (0 if c % 2 == 0 else rng.integers(-3, 4, n_ev) * 2000)
            with open(pth, "w") as f:
                f.write("HEADER LINE 1;;;\nBOARD;CHANNEL;TIMETAG;ENERGY;ENERGYSHORT;FLAGS;PROBE_CODE;SAMPLES\n")
                for i in range(n_ev):
                    w = 8000 + 10 * c + rng.normal(0, 3, L)
                    s0 = int(rng.integers(45, L - 30))
                    w[s0:s0 + 12] -= rng.uniform(30, 400) * np.exp(-np.arange(12) / 5.0)
                    f.write(";".join(["0", str(c), str(int(tc[i])), "0", "0", "0", "0"] + [str(int(v)) for v in np.round(w)]) + "\n")
            os.utime(pth, (1_700_000_000, 1_700_000_000))  # records.time = file mtime + timestamp // 1000: the same in both runs
            paths.append(pth)
    ctx = Context(storage_dir=os.path.join(work, f"cache_{mode}"))
    ctx.register(*ref_profiles.cpu_default())
    if mode == "b200":
        from waveformanalysis_b200 import plugin_api, profiles

        assert plugin_api.HAVE_REFERENCE
        ctx.register(*profiles.b200_default(), allow_override=True)
        assert type(ctx._plugins["hit_threshold"]).__name__ == "B200ThresholdHitPlugin"
    if scenario in ("csv", "mixed"):
        ctx.set_config({"daq_adapter": "vx2730"})
    else:
        ctx.set_config({"daq_adapter": "v1725"})
        for name in ("records", "wave_pool"):
            ctx.set_config({"daq_adapter": "v1725", "dt": 4}, plugin_name=name)
    if scenario in ("records", "csv"):
        ctx.set_config({"wave_source": "records", "height_range": (10, 60)}, plugin_name="basic_features")
        ctx.set_config({"wave_source": "records", "threshold": 12.0}, plugin_name="hit_threshold")
    elif scenario == "mixed":
        # polarity from the channel metadata (the float32 signal branch of the features), a Butterworth filter for one
        # channel next to the default Savitzky-Golay, hits on the filtered pool, charge widths on records
        ctx.set_config({"channel_metadata": {"channels": {"0:1": {"polarity": "negative"}, "0:2": {"polarity": "positive"}}}})
        fcc = {"channels": {"0:3": {"filter_type": "BW", "lowcut": 0.02, "highcut": 0.2, "fs": 1.0, "filter_order": 2}}}
        ctx.set_config({"channel_config": fcc}, plugin_name="wave_pool_filtered")
        ctx.set_config({"wave_source": "records", "height_range": (0, None), "area_range": (10, 100)}, plugin_name="basic_features")
        ctx.set_config({"wave_source": "records", "use_filtered": True, "threshold": 10.0, "right_extension": 1}, plugin_name="hit_threshold")
        ctx.set_config({"wave_source": "records"}, plugin_name="waveform_width_integral")
    else:
        # "waves": structured rows as the wave source (st_waveforms for the features, filtered_waveforms for the hits and
        # the charge widths), records on the filtered pool for `hit`, a per-channel threshold and a fixed baseline
        cc = {"channels": {"0:3": {"threshold": 30.0, "fixed_baseline": 7990.0}}}
        ctx.set_config({"wave_source": "st_waveforms", "height_range": (5, 200), "channel_config": cc}, plugin_name="basic_features")
        ctx.set_config({"wave_source": "filtered_waveforms", "threshold": 9.0, "left_extension": 1, "channel_config": cc}, plugin_name="hit_threshold")
        ctx.set_config({"wave_source": "records", "use_filtered": True, "height": 6.0, "width": 2}, plugin_name="hit")
        ctx.set_config({"wave_source": "filtered_waveforms"}, plugin_name="waveform_width_integral")
    for name in ("hit_merge_clusters", "hit_merged", "hit_merged_components"):
        ctx.set_config({"merge_gap_ns": 50.0}, plugin_name=name)
    ctx.set_config({"time_window_ns": 100.0}, plugin_name="hit_grouped")
    run = "run_real"
    ctx._set_data(run, "raw_files", [[p] for p in paths])

    out = {}
    names = ["st_waveforms", "records", "wave_pool", "wave_pool_filtered", "basic_features", "hit_threshold", "hit_merge_clusters",
             "hit_merged", "hit_merged_components", "hit_grouped"]
    if os.environ.get("WFB_REAL_CONTEXT_ALL", "1") != "0":  # the default-profile chain as well: hit -> waveform_width -> s1_s2, df ...
        names += ["hit", "waveform_width", "waveform_width_integral", "s1_s2", "df", "df_events", "df_paired"]
    if scenario == "waves":
        names = ["basic_features", "hit_threshold", "hit", "waveform_width", "waveform_width_integral"]
    if scenario == "csv":
        names = ["st_waveforms", "records", "wave_pool", "basic_features", "hit_threshold", "hit_merged", "hit_grouped"]
    if scenario == "mixed":
        names = ["st_waveforms", "records", "wave_pool_filtered", "basic_features", "hit_threshold", "waveform_width_integral"]
    for name in names:
        res = ctx.get_data(run, name)
        if hasattr(res, "columns"):  # DataFrame: one array per column (object columns flattened)
            for col in res.columns:
                v = res[col].to_numpy()
                if v.dtype == object:
                    lens = np.array([len(x) for x in v], dtype=np.int64)
                    flat = np.concatenate([np.asarray(x).reshape(-1) for x in v]) if len(v) else np.zeros(0)
                    out[f"{name}.{col}.len"] = lens
                    out[f"{name}.{col}.flat"] = flat
                else:
                    out[f"{name}.{col}"] = v
        else:
            out[name] = np.asarray(res)
    if scenario != "records":
        np.savez(out_path, **out)
        print("REAL_CONTEXT_DONE", mode, scenario, {k: (v.shape, str(v.dtype)[:40]) for k, v in out.items() if "." not in k})
        return
    # ---- the streaming side: signal_peaks_stream over the reference's own chunk iterator, hit_threshold_stream ----------
    from waveform_analysis.core.plugins.builtin.streaming.cpu.signal_peaks import SignalPeaksStreamPlugin

    if mode == "b200":
        from waveformanalysis_b200.plugins import B200HitThresholdStreamPlugin, B200SignalPeaksStreamPlugin

        peaks_plugin = B200SignalPeaksStreamPlugin()
        assert isinstance(peaks_plugin, SignalPeaksStreamPlugin)
    else:
        peaks_plugin = SignalPeaksStreamPlugin()
    ctx.register(peaks_plugin, allow_override=True)
    ctx.set_config({"height": 8.0}, plugin_name="signal_peaks_stream")
    ctx.get_data(run, "st_waveforms")
    ctx.get_data(run, "filtered_waveforms")
    # many chunks per channel: the device pipeline overlaps them
    chunks = list(peaks_plugin.compute(ctx, run, streaming_config={"chunk_size": 64, "parallel": False}))
    out["signal_peaks_stream"] = np.concatenate([c.data for c in chunks])
    out["signal_peaks_stream_bounds"] = np.array([[c.start, c.end] for c in chunks], dtype=np.int64)
    if mode == "b200":
        assert peaks_plugin.stream_stats["overlapped_chunks"] > 10, peaks_plugin.stream_stats
        hs = B200HitThresholdStreamPlugin()
        ctx.register(hs, allow_override=True)
        ctx.set_config({"wave_source": "records", "threshold": 12.0, "height_range": (10, 60)}, plugin_name="hit_threshold_stream")
        hchunks = list(hs.compute(ctx, run, streaming_config={"chunk_size": 500, "required_halo_ns": 2000}))
        assert len(hchunks) > 5 and hs.stream_stats["overlapped_chunks"] > 3, hs.stream_stats
        out["hit_threshold_stream"] = np.concatenate([c.data for c in hchunks])
        out["basic_features_stream"] = np.concatenate([c.metadata["basic_features"] for c in hchunks])
    else:  # the stream must reproduce the non-streaming rows
        out["hit_threshold_stream"] = out["hit_threshold"]
        out["basic_features_stream"] = out["basic_features"]
    np.savez(out_path, **out)
    print("REAL_CONTEXT_DONE", mode, {k: (v.shape, str(v.dtype)[:40]) for k, v in out.items() if "." not in k})


if __name__ == "__main__":
    main()
