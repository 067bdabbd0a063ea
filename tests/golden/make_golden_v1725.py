#!/usr/bin/env python
"""Golden vectors for the CAEN V1725 DAW_DEMO binary ingest: synthetic .bin blobs (built here from
the published layout the reference parses, utils/formats/v1725.py:62-114) run through the LIVE
reference (build_records_from_v1725_files, core/processing/records_builder.py:798-830).

    python tests/golden/make_golden_v1725.py      # rewrites tests/golden/v1725_golden.npz

Build container only; the fixtures travel with the repo."""

from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

from make_golden import import_reference  # noqa: E402
from waveformanalysis_b200.synth import make_v1725_blob  # noqa: E402


def main():
    import_reference()
    from waveform_analysis.core.processing.records_builder import build_records_from_v1725_files

    G = {}
    cases = {
        # one file, every event fires a different channel set, ragged lengths, trunc flags, timestamp ties
        "a": dict(files=[("run_b0_seg0.bin", dict(n_events=300, n_channels=16, seed=1, lengths=(8, 400), tie_every=7))], dt_ns=4),
        # two boards in two files plus a legacy name (board 0): cross-file merge with ties
        "b": dict(files=[("x_b3_seg0.bin", dict(n_events=120, n_channels=8, seed=2, lengths=(64, 64), tie_every=5)),
                         ("x_b1_seg1.bin", dict(n_events=150, n_channels=8, seed=3, lengths=(2, 130), tie_every=5, t0=40)),
                         ("CH2_0.bin", dict(n_events=90, n_channels=4, seed=4, lengths=(100, 100), tie_every=3, t0=17))], dt_ns=2),
    }
    for tag, case in cases.items():
        with tempfile.TemporaryDirectory() as tmp:
            paths = []
            for k, (name, kw) in enumerate(case["files"]):
                blob = make_v1725_blob(**kw)
                G[f"{tag}_blob{k}"] = np.frombuffer(blob, dtype=np.uint8)
                G[f"{tag}_name{k}"] = np.array(name)
                p = os.path.join(tmp, name)
                with open(p, "wb") as f:
                    f.write(blob)
                paths.append(p)
            bundle = build_records_from_v1725_files(paths, dt_ns=case["dt_ns"])
            G[f"{tag}_records"] = bundle.records
            G[f"{tag}_pool"] = bundle.wave_pool
            G[f"{tag}_dt_ns"] = np.array(case["dt_ns"])
            G[f"{tag}_nfiles"] = np.array(len(paths))
    # a truncated file: the reference stops at the first short read
    blob = make_v1725_blob(n_events=20, n_channels=4, seed=9, lengths=(32, 32))
    cut = blob[: len(blob) - 37]
    with tempfile.TemporaryDirectory() as tmp:
        p = os.path.join(tmp, "t_b5_seg0.bin")
        with open(p, "wb") as f:
            f.write(cut)
        bundle = build_records_from_v1725_files([p], dt_ns=4)
    G["cut_blob0"] = np.frombuffer(cut, dtype=np.uint8)
    G["cut_name0"] = np.array("t_b5_seg0.bin")
    G["cut_records"], G["cut_pool"], G["cut_dt_ns"], G["cut_nfiles"] = bundle.records, bundle.wave_pool, np.array(4), np.array(1)
    # ---------------------------------------------------------------- VX2730-style parts of different waveform widths
    from waveform_analysis.core.processing import records_builder as rb
    from waveform_analysis.utils.formats import get_adapter
    from waveformanalysis_b200.synth import make_raw_run

    adapter = get_adapter("vx2730")
    cols = adapter.format_spec.columns
    parts = []
    for c, (L, seed) in enumerate(((800, 5), (256, 6), (30, 7))):  # 30 < the 40-sample baseline window
        raw = make_raw_run(1, 90, L, seed=seed)
        arr = np.zeros((90, 7 + L), dtype=np.int64)
        arr[:, cols.board] = c % 2
        arr[:, cols.channel] = c
        arr[:, cols.timestamp] = raw["timestamps_ps"] + c
        arr[:, 7:] = raw["samples"]
        G[f"rag_ts{c}"], G[f"rag_samples{c}"] = arr[:, cols.timestamp].copy(), raw["samples"].astype(np.int16)
        parts.append(rb._build_records_part_from_raw_array(
            arr, channel_idx=c, default_dt_ns=2, cols=cols,
            normalize_timestamp_to_ps=adapter.format_spec.normalize_timestamp_to_ps, baseline_samples=None))
    bundle = rb.merge_records_parts(parts)
    bundle.records["record_id"] = np.arange(len(bundle.records))
    G["rag_records"], G["rag_pool"] = bundle.records, bundle.wave_pool
    out = os.path.join(HERE, "v1725_golden.npz")
    np.savez_compressed(out, **G)
    print("wrote", out, {k: v.shape for k, v in G.items() if k.endswith("records")}, os.path.getsize(out) / 1e3, "kB")


if __name__ == "__main__":
    main()
