#!/usr/bin/env python
"""Golden vectors for the `hit` plugin (HitFinderPlugin, scipy.signal.find_peaks per record): the LIVE
reference (core/plugins/builtin/cpu/peak_finding.py) on seeded synthetic inputs.

    python tests/golden/make_golden_hit.py        # rewrites tests/golden/hit_golden.npz

Build container only; the fixtures travel with the repo."""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

from make_golden import Ctx, import_reference, st_from_records  # noqa: E402
from waveformanalysis_b200.synth import make_raw_run, records_from_raw  # noqa: E402


def main():
    import_reference()
    from waveform_analysis.core.plugins.builtin.cpu.filtering import create_filtered_waveform_dtype
    from waveform_analysis.core.plugins.builtin.cpu.peak_finding import HitFinderPlugin
    from waveform_analysis.core.plugins.builtin.cpu.records import WavePoolFilteredPlugin
    from waveform_analysis.core.processing.dtypes import create_record_dtype

    def run(plugin, data, cfg):
        return plugin.compute(Ctx(cfg, data), "run")

    G = {}
    raw = make_raw_run(4, 80, 800, seed=321)
    rec, pool = records_from_raw(raw)
    st = st_from_records(rec, pool, create_record_dtype)
    fw = run(WavePoolFilteredPlugin(), {"records": rec, "wave_pool": pool}, {"max_workers": 1})
    stf = np.zeros(len(st), dtype=create_filtered_waveform_dtype(st.dtype))
    for f in st.dtype.names:
        if f != "wave":
            stf[f] = st[f]
    stf["wave"] = fw.reshape(len(st), 800)
    G["records"], G["pool"], G["filtered_pool"] = rec, pool, fw
    # default profile: derivative of the SG-filtered float32 waveforms, minmax height
    G["filt_default"] = run(HitFinderPlugin(), {"filtered_waveforms": stf, "st_waveforms": st}, {})
    G["filt_lowcut"] = run(HitFinderPlugin(), {"filtered_waveforms": stf, "st_waveforms": st},
                           {"height": 3.0, "prominence": 0.5, "width": 2, "distance": 6, "height_window_extension": 1})
    G["filt_thr"] = run(HitFinderPlugin(), {"filtered_waveforms": stf, "st_waveforms": st}, {"height": 5.0, "threshold": 0.5, "width": 1})
    G["filt_diffheight"] = run(HitFinderPlugin(), {"filtered_waveforms": stf, "st_waveforms": st}, {"height": 10.0, "height_method": "diff"})
    # raw int16 rows
    G["st_default"] = run(HitFinderPlugin(), {"st_waveforms": st}, {"use_filtered": False, "height": 12.0, "width": 2})
    G["st_level"] = run(HitFinderPlugin(), {"st_waveforms": st}, {"use_filtered": False, "use_derivative": False, "height": 20.0,
                                                                 "prominence": 4.0, "width": 3})
    # records source (RecordsView.signals, float32 baseline subtraction), unknown and positive polarity
    G["rec_default"] = run(HitFinderPlugin(), {"records": rec, "wave_pool": pool}, {"use_filtered": False, "wave_source": "records",
                                                                                     "height": 12.0, "width": 2})
    rawp = make_raw_run(3, 60, 800, seed=99, positive_pulses=True)
    rp_rec, rp_pool = records_from_raw(rawp, polarity="positive")
    G["pos_records"], G["pos_pool"] = rp_rec, rp_pool
    G["pos_level"] = run(HitFinderPlugin(), {"records": rp_rec, "wave_pool": rp_pool},
                         {"use_filtered": False, "wave_source": "records", "use_derivative": False, "height": 30.0, "prominence": 5.0,
                          "width": 2})
    G["pos_deriv"] = run(HitFinderPlugin(), {"records": rp_rec, "wave_pool": rp_pool},
                         {"use_filtered": False, "wave_source": "records", "height": 8.0, "width": 2})
    # ---------------------------------------------------------------- signal_peaks_stream (streaming plugin, float64 rows)
    from waveform_analysis.core.plugins.builtin.streaming.cpu.signal_peaks import SignalPeaksStreamPlugin

    def stream(cfg):
        plugin = SignalPeaksStreamPlugin()
        plugin.parallel = False
        chunks = list(plugin.compute(Ctx(cfg, {"filtered_waveforms": stf, "st_waveforms": st}), "run"))
        rows = np.concatenate([c.data for c in chunks])
        return rows, np.array([[c.start, c.end] for c in chunks], dtype=np.int64)

    G["stream_default"], G["stream_default_bounds"] = stream({"height": 10.0})
    G["stream_minmax"], G["stream_minmax_bounds"] = stream({"height": 10.0, "height_method": "minmax", "minmax_window_expand": 3, "width": 2})
    G["stream_level"], G["stream_level_bounds"] = stream({"use_derivative": False, "height": 25.0, "prominence": 4.0, "width": 3})
    out = os.path.join(HERE, "hit_golden.npz")
    np.savez_compressed(out, **G)
    print("wrote", out, {k: len(v) for k, v in G.items() if v.dtype.names and "position" in v.dtype.names}, os.path.getsize(out) / 1e3, "kB")


if __name__ == "__main__":
    main()
