"""st_waveforms on the B200 (reference: core/plugins/builtin/cpu/waveforms.py:644-799 WaveformStruct, :971-1260
WaveformsPlugin).

CSV / binary reading stays with the reference's DAQ readers (text parsing is host I/O); the per-channel raw
rows they return are uploaded once and the structured rows - baseline over the adapter's window (the first of the two
baselines), ``baseline_upstream`` from an upstream ``baseline`` plugin when ``use_upstream_baseline`` is set (the
second one, NaN otherwise), timestamps in ps, samples copied - are produced by ``wfb_structure_waveforms``.  Raw arrays
can also be seeded into the context as ``raw_arrays`` (one 2-D array per channel, header columns + samples).  V1725
``.bin`` input keeps the reference's own converter (its st_waveforms path goes through a per-waveform Python reader).
"""

from __future__ import annotations

from typing import Any

import numpy as np

from .. import ops
from ..dtypes import create_record_dtype
from ..plugin_api import HAVE_REFERENCE, Option, Plugin, resolve_dt_config
from .records import _apply_polarity, _read_raw_arrays


class B200WaveformsPlugin(Plugin):
    """Extract waveforms from raw files and structure them into ST_WAVEFORM rows (dual baseline)."""

    version = "0.10.0"
    provides = "st_waveforms"
    depends_on = []
    uses_run_config = True
    description = "Extract waveforms from raw CSV files and structure them into NumPy structured arrays."
    save_when = "always"
    # the reference's class default (DEFAULT_WAVE_LENGTH = 1500, core/processing/dtypes.py:16): the V1725 branch does not set a
    # per-run dtype, so the Context casts its rows to this width (core/plugins/core/validation.py:182-184)
    output_dtype = create_record_dtype(1500)
    options = {
        "daq_adapter": Option(default="vx2730", type=str, help="DAQ adapter name (e.g., 'vx2730')"),
        "wave_length": Option(default=None, type=int, help="Waveform length (samples); detected from the data when None"),
        "dt": Option(default=None, type=int, help="Sampling interval in ns for st_waveforms.dt (None=auto from adapter)."),
        "n_jobs": Option(default=None, type=int, help="unused on the GPU", track=False),
        "use_process_pool": Option(default=False, type=bool, help="unused on the GPU", track=False),
        "chunksize": Option(default=None, type=int, help="CSV read chunk size", track=False),
        "parse_engine": Option(default="auto", type=str, help="CSV engine: auto | polars | pyarrow | pandas", track=False),
        "use_upstream_baseline": Option(default=False, type=bool, help="use the baseline of an upstream 'baseline' plugin"),
        "baseline_samples": Option(default=None, type=None, help="Baseline range: int or (start, end) relative to samples_start."),
        "streaming_mode": Option(default=False, type=bool, help="unused on the GPU (rows are built in one pass)", track=False),
    }

    def resolve_depends_on(self, context: Any, run_id: str | None = None) -> list[str]:
        deps = ["raw_files"]
        if context.get_config(self, "use_upstream_baseline"):
            deps.append("baseline")
        return deps

    def compute(self, context: Any, run_id: str, **kwargs) -> np.ndarray:
        adapter_name = context.get_config(self, "daq_adapter")
        adapter_name = adapter_name.lower() if isinstance(adapter_name, str) else "vx2730"
        wave_length = context.get_config(self, "wave_length")
        dt_ns = resolve_dt_config(context, self, deprecated_keys=("dt_ns", "sampling_interval_ns"))
        if adapter_name == "v1725":
            if not HAVE_REFERENCE:
                raise RuntimeError("st_waveforms from V1725 .bin files needs the reference's converter (waveform_analysis)")
            from waveform_analysis.core.plugins.builtin.cpu.waveforms import WaveformsPlugin  # type: ignore

            return WaveformsPlugin.compute(self, context, run_id, **kwargs)
        arrays, adapter = _read_raw_arrays(context, run_id, adapter_name)
        if adapter is not None:
            cols = adapter.format_spec.columns
            c_board, c_chan, c_ts, s0 = cols.board, cols.channel, cols.timestamp, cols.samples_start
            bl0, bl1 = cols.baseline_start - s0, cols.baseline_end - s0
            if dt_ns is None and adapter.sampling_rate_hz:
                dt_ns = int(round(1e9 / float(adapter.sampling_rate_hz)))
            normalize = adapter.format_spec.normalize_timestamp_to_ps
        else:  # VX2730 CSV column layout (utils/formats/vx2730.py:80-107)
            c_board, c_chan, c_ts, s0, bl0, bl1 = 0, 1, 2, 7, 0, 40
            normalize = None
        if dt_ns is None:
            dt_ns = 2 if adapter is None else 1
        bs = context.get_config(self, "baseline_samples")
        if isinstance(bs, (list, tuple)):
            bl0, bl1 = int(bs[0]), int(bs[1])
        elif isinstance(bs, int):
            bl1 = bl0 + int(bs)
        upstream = None
        if context.get_config(self, "use_upstream_baseline"):
            try:
                upstream = context.get_data(run_id, "baseline")
            except Exception:
                upstream = None  # the reference logs a warning and fills NaN (waveforms.py:1161-1167)
        lengths = [a.shape[1] - s0 for _, a in arrays if a.ndim == 2 and len(a) and a.shape[1] > s0]
        wl = int(wave_length) if wave_length is not None else (max(lengths) if lengths else 800)
        dtype = create_record_dtype(wl)
        parts = []
        base = 0
        for ch_idx, arr in arrays:
            if arr.ndim != 2 or len(arr) == 0:
                continue
            t = arr[:, c_ts].astype(np.int64)
            ts = normalize(t, dt_ns=int(dt_ns)) if normalize is not None else t
            up = None
            if upstream is not None and ch_idx < len(upstream):
                u = upstream[ch_idx]
                if u is not None and len(u) == len(arr):  # waveforms.py:762-771
                    up = np.asarray(u, dtype=np.float64)
            # the reference clamps the window to the row's columns and uses NaN for an empty one
            parts.append(ops.structure_waveforms(ts, arr[:, c_board].astype(np.int16), arr[:, c_chan].astype(np.int16),
                                                 arr[:, s0:].astype(np.int16), dt_ns=int(dt_ns), wave_length=wl, baseline_window=(bl0, bl1),
                                                 baseline_upstream=up, record_base=base))
            base += len(arr)
        if not parts:
            return np.zeros(0, dtype=dtype)
        st = np.concatenate(parts) if len(parts) > 1 else parts[0]
        _apply_polarity(context, run_id, st)
        self.output_dtype = dtype
        return st
