"""signal_peaks_stream on the B200 (reference: core/plugins/builtin/streaming/cpu/signal_peaks.py:35-406).

The streaming framework itself - chunk iteration per channel / dt segment / time break, halo, clipping,
executors (core/plugins/core/streaming.py) - is host control logic and stays the reference's: when the
reference package is importable this class subclasses its ``SignalPeaksStreamPlugin`` and replaces only
``compute_chunk`` (the per-waveform scipy ``find_peaks`` loop) by one ``wfb_find_peaks`` launch per chunk.
Without the reference the class still offers ``compute_chunk`` for duck-typed chunks (``.data`` = st rows,
``.metadata["filtered_waveforms"]``), which is what the GPU-box tests use."""

from __future__ import annotations

from types import SimpleNamespace
from typing import Any

import numpy as np

from .. import ops
from ..plugin_api import HAVE_REFERENCE, Option, Plugin, resolve_dt_config

_OPTIONS = {
    "use_derivative": Option(default=True, type=bool, help="detect on the first difference (True) or on the level"),
    "height": Option(default=30.0, type=float, help="minimal peak height"),
    "distance": Option(default=2, type=int, help="minimal distance between peaks (samples)"),
    "prominence": Option(default=0.7, type=float, help="minimal prominence"),
    "width": Option(default=4, type=int, help="minimal width (samples)"),
    "threshold": Option(default=None, help="optional neighbour threshold"),
    "height_method": Option(default="diff", type=str, help="'diff' or 'minmax'"),
    "minmax_window_expand": Option(default=2, type=int, help="window extension of the minmax height"),
    "dt": Option(default=None, type=int, help="sample interval (ns), only used when the input has no dt field"),
}

if HAVE_REFERENCE:
    from waveform_analysis.core.plugins.builtin.streaming.cpu.signal_peaks import SignalPeaksStreamPlugin as _Base  # type: ignore
    from waveform_analysis.core.processing.chunk import TIMESTAMP_FIELD, Chunk  # type: ignore
else:
    _Base = Plugin
    TIMESTAMP_FIELD = "timestamp"

    def Chunk(**kw):  # noqa: N802 - stands in for core/processing/chunk.py:78-207
        return SimpleNamespace(**kw)


class B200SignalPeaksStreamPlugin(_Base):
    """Streaming peak detection: one device launch per chunk instead of a Python loop over waveforms."""

    provides = "signal_peaks_stream"
    depends_on = ["filtered_waveforms", "st_waveforms"]
    description = "Stream peak detection from filtered waveforms."
    version = "1.2.0"
    save_when = "never"
    output_dtype = None
    parallel = False          # chunks are processed in order on the caller's thread: the GPU is the parallel part
    executor_type = "thread"
    if not HAVE_REFERENCE:
        options = dict(_OPTIONS)
        output_data_kind = "peaks"

        def compute(self, context: Any, run_id: str, **kwargs):
            raise RuntimeError("signal_peaks_stream: the reference's streaming framework (waveform_analysis.core.plugins.core.streaming) "
                               "is needed to iterate chunks; call compute_chunk on your own chunks instead")

    def _load_config(self, context: Any) -> None:
        if HAVE_REFERENCE:
            super()._load_config(context)
            return
        for k in ("use_derivative", "height", "distance", "prominence", "width", "threshold", "height_method"):
            setattr(self, k, context.get_config(self, k))
        self.minmax_window_expand = max(0, int(context.get_config(self, "minmax_window_expand")))
        self.explicit_dt = resolve_dt_config(context, self, deprecated_keys=("sampling_interval_ns", "dt_ns"))

    def compute_chunk(self, chunk: Any, context: Any, run_id: str, **kwargs):
        st_chunk = chunk.data
        filtered_chunk = chunk.metadata.get("filtered_waveforms")
        if filtered_chunk is None or len(st_chunk) == 0:
            return None
        if not hasattr(self, "height_method"):
            self._load_config(context)
        peaks = ops.find_peaks_stream_chunk(
            st_chunk, filtered_chunk, explicit_dt=self.explicit_dt, event_offset=int(chunk.metadata.get("event_offset", 0)),
            use_derivative=bool(self.use_derivative), height=float(self.height), distance=int(self.distance),
            prominence=float(self.prominence), width=int(self.width),
            threshold=None if self.threshold is None else float(self.threshold), height_method=str(self.height_method),
            minmax_window_expand=int(self.minmax_window_expand))
        if len(peaks) == 0:
            return None
        return Chunk(data=peaks, start=int(np.min(peaks["timestamp"])), end=int(np.max(peaks["timestamp"])), run_id=run_id,
                     data_type=self.provides, data_kind=self.output_data_kind, time_field=TIMESTAMP_FIELD)
