"""hit_threshold on the B200 (reference: core/plugins/builtin/cpu/hit_finder.py:82-413)."""

from __future__ import annotations

from typing import Any

import numpy as np

from .. import engine
from . import _fused
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
        clamp = None
        if wave_input.spec.is_records:
            records, pool = wave_input.records, wave_input.wave_pool
            if records is None or pool is None:
                raise ValueError("hit_threshold failed to load records_view for records source")
            if len(records) == 0:
                return np.zeros(0, dtype=THRESHOLD_HIT_DTYPE)
            # records source: one fused, device-resident pass shared with basic_features (plugins/_fused.py)
            return _fused.records_pass(self, context, run_id, wave_input.spec, records, pool, "hits")
        data = wave_input.waveform_data
        if data is None:
            raise ValueError(f"hit_threshold failed to load {wave_input.spec.data_name}")
        if len(data) == 0:
            return np.zeros(0, dtype=THRESHOLD_HIT_DTYPE)
        dt_scalar = check_dt_array(data, explicit_dt, self.provides, wave_input.spec.data_name)
        records, pool, signed, clamp = structured_as_records(data, explicit_dt=dt_scalar, return_clamp=True)
        names = records.dtype.names
        boards = records["board"] if "board" in names else np.zeros(len(records), np.int16)
        channels = records["channel"] if "channel" in names else np.zeros(len(records), np.int16)
        thr = per_channel_option(channel_config_cfg, run_id, boards, channels, "threshold", threshold)
        thr = {k: float(v) for k, v in thr.items() if float(v) != threshold}
        if clamp is not None:
            # rows whose event_length differs from the row width: the whole row is scanned, the edges are clamped to
            # the source's event_length (hit_finder.py:183-230, 388-391)
            run = engine.DeviceRun.from_host(records, pool, explicit_dt=dt_scalar, clamp_lengths=clamp)
            out = run.run_to_host(features=False, hits=True, threshold=threshold, rules=engine.make_rules(thr, None),
                                  left_extension=left_extension, right_extension=right_extension, signed_samples=signed)
            return out["hits"]
        out = engine.process_host(records, pool, features=False, hits=True, threshold=threshold, thresholds=thr,
                                  left_extension=left_extension, right_extension=right_extension, explicit_dt=dt_scalar,
                                  signed_samples=signed)
        return out["hits"]


class B200HitFinderPlugin(Plugin):
    """`hit`: scipy.signal.find_peaks per waveform on the device (reference:
    core/plugins/builtin/cpu/peak_finding.py:43-614 HitFinderPlugin; same options and HIT_DTYPE rows)."""

    provides = "hit"
    depends_on = []
    description = "Detect peaks in waveforms and extract peak features."
    version = "3.0.0"
    save_when = "always"
    from ..dtypes import HIT_DTYPE as output_dtype  # noqa: N811

    options = {
        "use_filtered": Option(default=True, type=bool, help="use filtered_waveforms (needs the filtered waveform plugin)"),
        "wave_source": Option(default=WAVE_SOURCE_AUTO, type=str, help="auto|records|st_waveforms|filtered_waveforms"),
        "use_derivative": Option(default=True, type=bool, help="detect on the first difference (True) or on the level"),
        "height": Option(default=30.0, type=float, help="minimal peak height"),
        "distance": Option(default=2, type=int, help="minimal distance between peaks (samples)"),
        "prominence": Option(default=0.7, type=float, help="minimal prominence"),
        "width": Option(default=4, type=int, help="minimal width (samples)"),
        "threshold": Option(default=None, help="optional neighbour threshold"),
        "height_method": Option(default="minmax", type=str, help="'diff' or 'minmax'"),
        "height_window_extension": Option(default=4, type=int, help="window extension of the minmax height"),
        "dt": Option(default=None, type=int, help="sample interval (ns), only used when the input has no dt field"),
        "parallel": Option(default=True, type=bool, help="unused on the GPU"),
        "n_workers": Option(default=0, type=int, help="unused on the GPU"),
        "chunk_size": Option(default=1024, type=int, help="unused on the GPU"),
        "parallel_min_events": Option(default=20480, type=int, help="unused on the GPU"),
    }

    def resolve_depends_on(self, context: Any, run_id: str | None = None) -> list[str]:
        return list(resolve_wave_input_spec(context, self).depends_on)

    def compute(self, context: Any, run_id: str, **_kwargs) -> np.ndarray:
        from .. import ops
        from ..dtypes import HIT_DTYPE

        threshold = context.get_config(self, "threshold")
        opts = dict(use_derivative=bool(context.get_config(self, "use_derivative")), height=float(context.get_config(self, "height")),
                    distance=int(context.get_config(self, "distance")), prominence=float(context.get_config(self, "prominence")),
                    width=int(context.get_config(self, "width")), threshold=None if threshold is None else float(threshold),
                    height_method=str(context.get_config(self, "height_method")),
                    height_window_extension=int(context.get_config(self, "height_window_extension")))
        explicit_dt = resolve_dt_config(context, self, deprecated_keys=("sampling_interval_ns", "dt_ns"))
        wave_input = load_wave_input(context, self, run_id, needs_wave_samples=True)
        if wave_input.spec.is_records:
            records, pool = wave_input.records, wave_input.wave_pool
            if records is None or pool is None:
                raise ValueError("hit failed to load records_view for records source")
            if len(records) == 0:
                return np.zeros(0, dtype=HIT_DTYPE)
            if "dt" not in (records.dtype.names or ()):
                if explicit_dt is None:
                    raise ValueError("[hit] records is missing required field 'dt'; provide explicit config 'dt'.")
                from numpy.lib import recfunctions as rfn

                records = rfn.append_fields(records, "dt", np.full(len(records), int(explicit_dt), np.int32), usemask=False)
            if np.any(records["dt"] <= 0):
                raise ValueError("[hit] dt must be > 0")
            from .. import residency
            from ..dtypes import RECORDS_DTYPE

            run = None
            if records.dtype == RECORDS_DTYPE and residency.fits_device(int(pool.nbytes)):
                run = residency.device_run(run_id, records, pool, wave_input.spec.wave_pool_name or "wave_pool")
            return ops.find_peaks_records(records, pool, run=run, **opts)
        data = wave_input.waveform_data
        if data is None:
            raise ValueError("hit failed to load waveform input")
        if len(data) == 0:
            return np.zeros(0, dtype=HIT_DTYPE)
        if "dt" in (data.dtype.names or ()) and np.any(data["dt"] <= 0):
            raise ValueError("[hit] dt must be > 0")
        return ops.find_peaks_waveforms(data, explicit_dt=explicit_dt, **opts)
