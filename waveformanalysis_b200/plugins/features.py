"""basic_features on the B200 (reference: core/plugins/builtin/cpu/basic_features.py:43-278)."""

from __future__ import annotations

from typing import Any

import numpy as np

from .. import engine
from . import _fused
from ..aos import structured_as_records
from ..channel_config import per_channel_option
from ..dtypes import BASIC_FEATURES_DTYPE
from ..plugin_api import Option, Plugin
from ..wave_source import WAVE_SOURCE_AUTO, load_wave_input, resolve_wave_input_spec

PEAK_RANGE = (40, 90)  # core/foundation/constants.py:22 FeatureDefaults.PEAK_RANGE


class B200BasicFeaturesPlugin(Plugin):
    """height / amp / area / max_abs_diff per record, fused single pass over the samples."""

    provides = "basic_features"
    depends_on = []
    description = "Compute basic height, amplitude, area, and max-abs-diff features from waveform data."
    version = "4.0.0"
    save_when = "always"
    output_dtype = BASIC_FEATURES_DTYPE
    options = {
        "height_range": Option(default=PEAK_RANGE, type=tuple, help="height range (start, end)"),
        "area_range": Option(default=(0, None), type=tuple, help="area range (start, end); end=None integrates to the end"),
        "use_filtered": Option(default=False, type=bool, help="use the filtered waveform source"),
        "wave_source": Option(default=WAVE_SOURCE_AUTO, type=str, help="auto|records|st_waveforms|filtered_waveforms"),
        "fixed_baseline": Option(default=None, type=dict, help="deprecated; use channel_config"),
        "channel_config": Option(default=None, type=dict, help="per (board, channel) overrides, may set fixed_baseline"),
    }

    def resolve_depends_on(self, context: Any, run_id: str | None = None) -> list[str]:
        return list(resolve_wave_input_spec(context, self).depends_on)

    def compute(self, context: Any, run_id: str, **kwargs) -> np.ndarray:
        wave_input = load_wave_input(context, self, run_id, needs_wave_samples=True)
        if wave_input.spec.is_records:
            records, pool = wave_input.records, wave_input.wave_pool
            if records is None or pool is None:
                raise ValueError("basic_features failed to load records_view for records source")
            # records source: one fused, device-resident pass shared with hit_threshold (plugins/_fused.py)
            return _fused.records_pass(self, context, run_id, wave_input.spec, records, pool, "features")
        channel_config_cfg = context.get_config(self, "channel_config")
        height_range = context.get_config(self, "height_range")
        area_range = context.get_config(self, "area_range")
        data = wave_input.waveform_data
        if data is None:
            raise ValueError(f"basic_features failed to load {wave_input.spec.data_name}")
        if len(data) == 0:
            return np.zeros(0, dtype=BASIC_FEATURES_DTYPE)
        records, pool, signed = structured_as_records(data, raw_polarity=True)
        names = records.dtype.names
        boards = records["board"] if "board" in names else np.zeros(len(records), np.int16)
        channels = records["channel"] if "channel" in names else np.zeros(len(records), np.int16)
        fixed = per_channel_option(channel_config_cfg, run_id, boards, channels, "fixed_baseline", None)
        fixed = {k: float(v) for k, v in fixed.items() if v is not None}
        out = engine.process_host(records, pool, features=True, hits=False, height_range=tuple(height_range),
                                  area_range=tuple(area_range), fixed_baselines=fixed, explicit_dt=1, signed_samples=signed)
        return out["features"]
