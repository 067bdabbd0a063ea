"""hit_threshold on the B200 (reference: core/plugins/builtin/cpu/hit_finder.py:82-413)."""

from __future__ import annotations

from typing import Any

import numpy as np

from .. import engine
from ..aos import structured_as_records
from ..channel_config import per_channel_option
from ..dtypes import THRESHOLD_HIT_DTYPE
from ..plugin_api import Option, Plugin, check_dt_array, resolve_dt_config
from ..wave_source import WAVE_SOURCE_AUTO, load_wave_input, resolve_wave_input_spec


class B200ThresholdHitPlugin(Plugin):
    """Threshold-only hit detector with THRESHOLD_HIT_DTYPE output."""

    provides = "hit_threshold"
    depends_on = []
    description = "Threshold-only hit detector with THRESHOLD_HIT_DTYPE output."
    version = "0.11.0"
    output_dtype = THRESHOLD_HIT_DTYPE
    save_when = "always"
    options = {
        "threshold": Option(default=10.0, type=float, help="hit threshold"),
        "use_filtered": Option(default=False, type=bool, help="use the filtered waveform source"),
        "wave_source": Option(default=WAVE_SOURCE_AUTO, type=str, help="auto|records|st_waveforms|filtered_waveforms"),
        "left_extension": Option(default=2, type=int, help="samples added left of the threshold region"),
        "right_extension": Option(default=2, type=int, help="samples added right of the threshold region"),
        "dt": Option(default=None, type=int, help="sample interval (ns), only used when the input has no dt field"),
        "channel_config": Option(default=None, type=dict, help="per (board, channel) overrides, may set threshold"),
    }

    def resolve_depends_on(self, context: Any, run_id: str | None = None) -> list[str]:
        return list(resolve_wave_input_spec(context, self).depends_on)

    def compute(self, context: Any, run_id: str, **_kwargs) -> np.ndarray:
        threshold = float(context.get_config(self, "threshold"))
        left_extension = max(0, int(context.get_config(self, "left_extension")))
        right_extension = max(0, int(context.get_config(self, "right_extension")))
        explicit_dt = resolve_dt_config(context, self, deprecated_keys=("sampling_interval_ns", "dt_ns"))
        channel_config_cfg = context.get_config(self, "channel_config")
        wave_input = load_wave_input(context, self, run_id, needs_wave_samples=True)
        signed = False
        if wave_input.spec.is_records:
            records, pool = wave_input.records, wave_input.wave_pool
            if records is None or pool is None:
                raise ValueError("hit_threshold failed to load records_view for records source")
            if len(records) == 0:
                return np.zeros(0, dtype=THRESHOLD_HIT_DTYPE)
            dt_scalar = check_dt_array(records, explicit_dt, self.provides, "records")
        else:
            data = wave_input.waveform_data
            if data is None:
                raise ValueError(f"hit_threshold failed to load {wave_input.spec.data_name}")
            if len(data) == 0:
                return np.zeros(0, dtype=THRESHOLD_HIT_DTYPE)
            dt_scalar = check_dt_array(data, explicit_dt, self.provides, wave_input.spec.data_name)
            records, pool, signed = structured_as_records(data, explicit_dt=dt_scalar, check_event_length=True)
        names = records.dtype.names
        boards = records["board"] if "board" in names else np.zeros(len(records), np.int16)
        channels = records["channel"] if "channel" in names else np.zeros(len(records), np.int16)
        thr = per_channel_option(channel_config_cfg, run_id, boards, channels, "threshold", threshold)
        thr = {k: float(v) for k, v in thr.items() if float(v) != threshold}
        out = engine.process_host(records, pool, features=False, hits=True, threshold=threshold, thresholds=thr,
                                  left_extension=left_extension, right_extension=right_extension, explicit_dt=dt_scalar,
                                  signed_samples=signed)
        return out["hits"]
