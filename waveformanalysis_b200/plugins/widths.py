"""waveform_width / waveform_width_integral on the B200 (reference:
core/plugins/builtin/cpu/waveform_width.py:40-374, waveform_width_integral.py:42-235)."""

from __future__ import annotations

from typing import Any

import numpy as np

from .. import ops
from ..aos import structured_as_records
from ..dtypes import WAVEFORM_WIDTH_DTYPE, WAVEFORM_WIDTH_INTEGRAL_DTYPE
from ..plugin_api import Option, Plugin
from ..wave_source import WAVE_SOURCE_AUTO, load_wave_input, resolve_wave_input_spec


class B200WaveformWidthPlugin(Plugin):
    provides = "waveform_width"
    depends_on = []
    description = "Calculate rise/fall time based on peak detection results."
    version = "3.0.0"
    save_when = "always"
    output_dtype = WAVEFORM_WIDTH_DTYPE
    options = {
        "use_filtered": Option(default=False, type=bool, help="use filtered_waveforms"),
        "sampling_rate": Option(default=None, type=float, help="sampling rate (GHz); 0.5 when unset"),
        "rise_low": Option(default=0.1, type=float, help="low fraction of the rise time"),
        "rise_high": Option(default=0.9, type=float, help="high fraction of the rise time"),
        "fall_high": Option(default=0.9, type=float, help="high fraction of the fall time"),
        "fall_low": Option(default=0.1, type=float, help="low fraction of the fall time"),
        "interpolation": Option(default=True, type=bool, help="linear interpolation of the crossings"),
    }

    def resolve_depends_on(self, context: Any, run_id: str | None = None) -> list[str]:
        if context.get_config(self, "use_filtered"):
            return ["hit", "filtered_waveforms"]
        return ["hit", "st_waveforms"]

    def compute(self, context: Any, run_id: str, **_kwargs) -> np.ndarray:
        use_filtered = context.get_config(self, "use_filtered")
        sampling_rate = context.get_config(self, "sampling_rate")
        if sampling_rate is None:
            sampling_rate = 0.5
        hits = context.get_data(run_id, "hit")
        waveform_data = context.get_data(run_id, "filtered_waveforms" if use_filtered else "st_waveforms")
        if not isinstance(hits, np.ndarray):
            raise ValueError("waveform_width expects hit as a single structured array")
        if not isinstance(waveform_data, np.ndarray):
            raise ValueError("waveform_width expects st_waveforms as a single structured array")
        if len(hits) == 0 or len(waveform_data) == 0:
            return np.zeros(0, dtype=WAVEFORM_WIDTH_DTYPE)
        if "record_id" in hits.dtype.names:
            h = hits
        else:  # legacy hits carry event_index (waveform_width.py:154-158)
            h = np.zeros(len(hits), dtype=[(n, hits.dtype[n]) for n in hits.dtype.names] + [("record_id", "i8")])
            for n in hits.dtype.names:
                h[n] = hits[n]
            h["record_id"] = hits["event_index"]
        names = waveform_data.dtype.names or ()
        if "record_id" in names:
            rids = waveform_data["record_id"]
        else:  # row index addressing (:169-172)
            rids = np.arange(len(waveform_data), dtype=np.int64)
        return ops.waveform_width(h, rids, waveform_data["wave"], sampling_rate=sampling_rate,
                                  rise_low=context.get_config(self, "rise_low"), rise_high=context.get_config(self, "rise_high"),
                                  fall_high=context.get_config(self, "fall_high"), fall_low=context.get_config(self, "fall_low"),
                                  interpolation=context.get_config(self, "interpolation"))


class B200WaveformWidthIntegralPlugin(Plugin):
    provides = "waveform_width_integral"
    depends_on = []
    description = "Event-wise integral quantile width using st_waveforms or filtered_waveforms."
    version = "2.7.0"
    save_when = "always"
    output_dtype = WAVEFORM_WIDTH_INTEGRAL_DTYPE
    options = {
        "q_low": Option(default=0.10, type=float, help="low quantile"),
        "q_high": Option(default=0.90, type=float, help="high quantile"),
        "use_filtered": Option(default=False, type=bool, help="use the filtered waveform source"),
        "wave_source": Option(default=WAVE_SOURCE_AUTO, type=str, help="auto|records|st_waveforms|filtered_waveforms"),
        "sampling_rate": Option(default=0.5, type=float, help="sampling rate (GHz)"),
        "dt": Option(default=None, type=float, help="sample interval (ns), wins over sampling_rate"),
    }

    def resolve_depends_on(self, context: Any, run_id: str | None = None) -> list[str]:
        return list(resolve_wave_input_spec(context, self).depends_on)

    def compute(self, context: Any, run_id: str, **_kwargs) -> np.ndarray:
        q_low = float(context.get_config(self, "q_low"))
        q_high = float(context.get_config(self, "q_high"))
        dt = context.get_config(self, "dt")
        sampling_rate = context.get_config(self, "sampling_rate")
        wave_input = load_wave_input(context, self, run_id, needs_wave_samples=True)
        if dt is None:
            if sampling_rate <= 0:
                raise ValueError(f"sampling_rate ({sampling_rate}) must be > 0")
            dt = 1.0 / float(sampling_rate)
        if q_low <= 0 or q_high >= 1 or q_low >= q_high:
            raise ValueError(f"q_low/q_high invalid: q_low={q_low}, q_high={q_high}")
        signed = False
        if wave_input.spec.is_records:
            records, pool = wave_input.records, wave_input.wave_pool
            if records is None or pool is None:
                raise ValueError("waveform_width_integral failed to load records_view for records source")
        else:
            data = wave_input.waveform_data
            if data is None:
                raise ValueError(f"waveform_width_integral failed to load {wave_input.spec.data_name}")
            if len(data) == 0:
                return np.zeros(0, dtype=WAVEFORM_WIDTH_INTEGRAL_DTYPE)
            records, pool, signed = structured_as_records(data, raw_polarity=True)
            # st branch: raw float64 arithmetic, 'positive' keeps the sign (waveform_width_integral.py:187-191)
            records = records.copy()
            records["polarity"] = np.where(records["polarity"] == "rawpos", "rawpos", "unknown")
        if len(records) == 0:
            return np.zeros(0, dtype=WAVEFORM_WIDTH_INTEGRAL_DTYPE)
        run = None
        if wave_input.spec.is_records:
            from .. import residency
            from ..dtypes import RECORDS_DTYPE

            if records.dtype == RECORDS_DTYPE and residency.fits_device(int(pool.nbytes)):
                run = residency.device_run(run_id, records, pool, wave_input.spec.wave_pool_name or "wave_pool")
        return ops.width_integral(records, pool, q_low=q_low, q_high=q_high, dt=float(dt), signed_samples=signed, run=run)
