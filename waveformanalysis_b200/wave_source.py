"""wave_source / use_filtered -> dynamic depends_on and input loading, mirroring
core/plugins/builtin/cpu/_wave_source.py:66-229 so dependency resolution is identical to the
CPU plugins'."""

from __future__ import annotations

import logging
from dataclasses import dataclass
from typing import Any

import numpy as np

logger = logging.getLogger(__name__)

WAVE_SOURCE_AUTO = "auto"
WAVE_SOURCE_RECORDS = "records"
WAVE_SOURCE_ST = "st_waveforms"
WAVE_SOURCE_FILTERED = "filtered_waveforms"
WAVE_SOURCES = {WAVE_SOURCE_AUTO, WAVE_SOURCE_RECORDS, WAVE_SOURCE_ST, WAVE_SOURCE_FILTERED}


@dataclass(frozen=True)
class WaveInputSpec:
    source: str
    use_filtered: bool
    data_name: str
    depends_on: tuple
    is_records: bool
    wave_pool_name: str | None = None


@dataclass
class LoadedWaveInput:
    spec: WaveInputSpec
    records: np.ndarray | None = None
    wave_pool: np.ndarray | None = None
    waveform_data: np.ndarray | None = None


def normalize_wave_source(value: Any) -> str:
    if value is None:
        return WAVE_SOURCE_AUTO
    source = str(value).strip().lower()
    if source not in WAVE_SOURCES:
        raise ValueError(f"Invalid wave_source: {value!r}. Expected one of {sorted(WAVE_SOURCES)}.")
    return source


def resolve_wave_input_spec(context: Any, plugin: Any, *, use_filtered_option: str = "use_filtered", needs_wave_samples: bool = True) -> WaveInputSpec:
    source = normalize_wave_source(context.get_config(plugin, "wave_source"))
    has_uf = use_filtered_option in getattr(plugin, "options", {})
    use_filtered = bool(context.get_config(plugin, use_filtered_option)) if has_uf else False
    if source not in (WAVE_SOURCE_AUTO, WAVE_SOURCE_RECORDS) and use_filtered:
        logger.warning("Ignoring %s=%s because wave_source=%s explicitly selects data source.", plugin.provides, use_filtered_option, source)
    if source == WAVE_SOURCE_RECORDS:
        pool = "wave_pool_filtered" if use_filtered else "wave_pool"
        deps = (WAVE_SOURCE_RECORDS, pool) if needs_wave_samples else (WAVE_SOURCE_RECORDS,)
        return WaveInputSpec(source, use_filtered, WAVE_SOURCE_RECORDS, deps, True, pool)
    if source == WAVE_SOURCE_ST:
        return WaveInputSpec(source, use_filtered, WAVE_SOURCE_ST, (WAVE_SOURCE_ST,), False)
    if source == WAVE_SOURCE_FILTERED:
        return WaveInputSpec(source, use_filtered, WAVE_SOURCE_FILTERED, (WAVE_SOURCE_FILTERED,), False)
    name = WAVE_SOURCE_FILTERED if use_filtered else WAVE_SOURCE_ST
    return WaveInputSpec(source, use_filtered, name, (name,), False)


def _ensure_registered(context: Any, data_name: str, consumer: str, hint: str | None = None) -> None:
    plugins = getattr(context, "_plugins", None)
    if not isinstance(plugins, dict) or not plugins or data_name in plugins:
        return
    msg = f"{consumer} requires '{data_name}' but it is not registered."
    if hint:
        msg += f" Register {hint} to provide '{data_name}'."
    raise KeyError(msg)


def load_wave_input(context: Any, plugin: Any, run_id: str, *, use_filtered_option: str = "use_filtered", needs_wave_samples: bool = True) -> LoadedWaveInput:
    spec = resolve_wave_input_spec(context, plugin, use_filtered_option=use_filtered_option, needs_wave_samples=needs_wave_samples)
    if spec.is_records:
        _ensure_registered(context, WAVE_SOURCE_RECORDS, plugin.provides, "RecordsPlugin")
        records = context.get_data(run_id, WAVE_SOURCE_RECORDS)
        if not isinstance(records, np.ndarray):
            raise ValueError(f"records_view requires formal '{WAVE_SOURCE_RECORDS}' plugin output")
        pool = None
        if needs_wave_samples:
            pool_name = spec.wave_pool_name or "wave_pool"
            _ensure_registered(context, pool_name, plugin.provides,
                               "WavePoolFilteredPlugin" if pool_name == "wave_pool_filtered" else "WavePoolPlugin")
            pool = context.get_data(run_id, pool_name)
            if not isinstance(pool, np.ndarray):
                raise ValueError(f"records_view requires formal '{pool_name}' plugin output")
        return LoadedWaveInput(spec=spec, records=records, wave_pool=pool)
    _ensure_registered(context, spec.data_name, plugin.provides,
                       "FilteredWaveformsPlugin" if spec.data_name == WAVE_SOURCE_FILTERED else None)
    data = context.get_data(run_id, spec.data_name)
    if not isinstance(data, np.ndarray):
        raise ValueError(f"{plugin.provides} expects {spec.data_name} as a single structured array")
    return LoadedWaveInput(spec=spec, waveform_data=data)
