"""records / wave_pool on the B200 (reference: core/plugins/builtin/cpu/records.py:209-331,
core/processing/records_builder.py:429-642).

CSV reading stays with the reference's DAQ readers (text parsing is host I/O);
the per-channel raw int16 rows they return are uploaded once and the baseline, the global
(timestamp, pid, board, channel, input order) sort, the gather into the contiguous wave_pool and the
packed RECORDS_DTYPE rows are produced by the K1 kernels.  Without the reference package installed
the plugins still build records from raw arrays seeded into the context as ``raw_arrays``.
V1725 ``.bin`` streams (``daq_adapter="v1725"``) do not go through the reference's per-waveform
Python reader at all: the file bytes are indexed by one host pass over the header chain and decoded,
sorted and packed on the device (``ops.build_records_from_v1725``)."""

from __future__ import annotations

from typing import Any

import numpy as np

from .. import ops
from ..dtypes import RECORDS_DTYPE
from ..plugin_api import HAVE_REFERENCE, Option, Plugin

_CACHE_ATTR = "_b200_records_bundle"


def _seeded(context: Any, run_id: str, name: str):
    """Data seeded into the context under ``name`` (tests, embedding applications), None if there is none.  A real
    Context raises for names no plugin provides, so memory is looked at first."""
    results = getattr(context, "_results", None)
    if isinstance(results, dict) and (run_id, name) in results:
        return results[(run_id, name)]
    plugins = getattr(context, "_plugins", None)
    if isinstance(plugins, dict) and plugins and name not in plugins:
        return None
    try:
        return context.get_data(run_id, name) if hasattr(context, "get_data") else None
    except Exception:
        return None


def _read_raw_arrays(context: Any, run_id: str, adapter_name: str):
    """[(channel_index, raw 2-D array (n, header + L))] via the reference's readers."""
    raw = _seeded(context, run_id, "raw_arrays")
    if raw is not None:
        return list(enumerate(raw)), None
    if not HAVE_REFERENCE:
        raise RuntimeError("records: the reference's DAQ readers (waveform_analysis.utils.formats) are needed to parse raw files; "
                           "seed 'raw_arrays' into the context or install the reference package")
    from waveform_analysis.utils.formats import get_adapter  # type: ignore

    raw_files = context.get_data(run_id, "raw_files")
    if not isinstance(raw_files, list):
        raise ValueError("records expects raw_files as a list of per-channel file groups")
    adapter = get_adapter(adapter_name)
    reader = adapter.format_reader
    out = []
    for ch_idx, files in enumerate(raw_files):
        if not files:
            continue
        parts = [a for a in reader.read_files_generator(list(files), chunk_size=1) if a.size]
        if parts:
            out.append((ch_idx, np.vstack(parts)))
    return out, adapter


def _apply_polarity(context: Any, run_id: str, records: np.ndarray) -> None:
    """records["polarity"] from the channel metadata layers (records.py:40-63 of the reference)."""
    if not HAVE_REFERENCE or len(records) == 0:
        return
    try:
        from waveform_analysis.core.hardware.channel import HardwareChannel  # type: ignore
        from waveform_analysis.core.plugins.builtin.cpu.waveforms import _build_polarity_lookup  # type: ignore
    except Exception:
        return
    keys = np.unique(np.stack([records["board"].astype(np.int64), records["channel"].astype(np.int64)], axis=1), axis=0)
    try:
        lookup = _build_polarity_lookup(context, run_id, keys[:, 0], keys[:, 1])
    except Exception:
        return
    for b, c in keys.tolist():
        pol = lookup.get(HardwareChannel(int(b), int(c)), "unknown")
        if pol != "unknown":
            records["polarity"][(records["board"] == b) & (records["channel"] == c)] = pol


def _v1725_bundle(context: Any, run_id: str, plugin: Plugin):
    """raw_files (groups of .bin paths) -> records + wave_pool (records.py:133-152 of the reference)."""
    raw_files = context.get_data(run_id, "raw_files")
    paths, seen = [], set()
    for group in raw_files or []:
        for path in group or []:
            if path not in seen:
                seen.add(path)
                paths.append(path)
    dt_ns = context.get_config(plugin, "dt")
    if dt_ns is None:
        dt_ns = 4  # 250 MS/s (utils/formats/v1725.py:189)
    blobs = [np.fromfile(str(p), dtype=np.uint8) for p in paths]
    return ops.build_records_from_v1725(blobs, [str(p) for p in paths], int(dt_ns))


def build_bundle(context: Any, run_id: str, plugin: Plugin):
    cache = getattr(context, "_results", None)
    key = (run_id, _CACHE_ATTR)
    if isinstance(cache, dict) and key in cache:
        return cache[key]
    adapter_name = (context.get_config(plugin, "daq_adapter") or getattr(context, "config", {}).get("daq_adapter") or "vx2730").lower()
    if adapter_name == "v1725" and _seeded(context, run_id, "raw_arrays") is None:
        bundle = _v1725_bundle(context, run_id, plugin)
        _apply_polarity(context, run_id, bundle[0])
        if isinstance(cache, dict):
            cache[key] = bundle
        return bundle
    arrays, adapter = _read_raw_arrays(context, run_id, adapter_name)
    dt_ns = context.get_config(plugin, "dt")
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
        dt_ns = 1
    bs = context.get_config(plugin, "baseline_samples")
    if isinstance(bs, (list, tuple)):
        bl0, bl1 = int(bs[0]), int(bs[1])
    elif isinstance(bs, int):
        bl1 = bl0 + int(bs)
    # records.time = file epoch + timestamp // 1000 (records.py:168-180, records_builder.py:507-510)
    epoch_ns = None
    if adapter is not None and hasattr(adapter, "get_file_epoch"):
        from pathlib import Path

        raw_files = _seeded(context, run_id, "raw_files")
        if raw_files is None:
            try:
                raw_files = context.get_data(run_id, "raw_files")
            except Exception:
                raw_files = None
        first_file = next((group[0] for group in (raw_files or []) if group), None)
        if first_file is not None:
            try:
                epoch_ns = adapter.get_file_epoch(Path(first_file))
            except (FileNotFoundError, OSError):
                epoch_ns = None
    ts, boards, chans, samples = [], [], [], []
    for ch_idx, arr in arrays:
        t = arr[:, c_ts].astype(np.int64)
        ts.append(normalize(t, dt_ns=int(dt_ns)) if normalize is not None else t)
        boards.append(arr[:, c_board].astype(np.int16))
        chans.append(arr[:, c_chan].astype(np.int16))
        samples.append(arr[:, s0:].astype(np.int16))
    d_pool = None
    if not samples:
        bundle = (np.zeros(0, dtype=RECORDS_DTYPE), np.zeros(0, dtype=np.uint16))
    else:
        widths = {s.shape[1] for s in samples}
        if len(widths) != 1:  # channels with different waveform widths: ragged wave_pool
            bundle = ops.build_records_ragged(np.concatenate(ts), np.concatenate(boards), np.concatenate(chans), samples,
                                              dt_ns=int(dt_ns), baseline_window=(bl0, bl1), epoch_ns=epoch_ns)
        else:
            rec_h, pool_h, d_pool = ops.build_records(np.concatenate(ts), np.concatenate(boards), np.concatenate(chans), np.vstack(samples),
                                                      dt_ns=int(dt_ns), baseline_window=(bl0, bl1), epoch_ns=epoch_ns, return_device=True)
            bundle = (rec_h, pool_h)
    _apply_polarity(context, run_id, bundle[0])
    if d_pool is not None and len(bundle[0]):
        # the pool was gathered on the device: it stays there for the plugins downstream (records.py:441-464 shares
        # the in-memory bundle the same way); the records rows are re-read from the host AFTER the polarity was applied
        from .. import engine, residency

        if residency.fits_device(int(bundle[1].nbytes)):
            residency.adopt_run(run_id, bundle[0], bundle[1], "wave_pool", engine.DeviceRun.from_device_pool(bundle[0], d_pool))
    if isinstance(cache, dict):
        cache[key] = bundle
    return bundle


_RECORDS_OPTIONS = {
    "daq_adapter": Option(default="vx2730", type=str, help="DAQ adapter name for records bundle (e.g., 'vx2730', 'v1725')."),
    "channel_workers": Option(default=None, help="unused on the GPU", track=False),
    "channel_executor": Option(default="thread", type=str, help="unused on the GPU", track=False),
    "n_jobs": Option(default=None, type=int, help="unused on the GPU", track=False),
    "use_process_pool": Option(default=False, type=bool, help="unused on the GPU", track=False),
    "chunksize": Option(default=None, type=int, help="CSV read chunk size", track=False),
    "parse_engine": Option(default="auto", type=str, help="CSV engine: auto | polars | pyarrow | pandas", track=False),
    "records_part_size": Option(default=250_000, type=int, help="unused on the GPU (one global device sort)"),
    "dt": Option(default=None, type=int, help="Sample interval in ns for records.dt (defaults to adapter rate or 1ns)."),
    "baseline_samples": Option(default=None, type=None, help="Baseline range: int or (start, end) relative to samples_start."),
}


class B200RecordsPlugin(Plugin):
    provides = "records"
    depends_on = []
    description = "Build records (event index table) from the shared internal records bundle."
    version = "0.10.0"
    save_when = "always"
    uses_run_config = True
    output_dtype = RECORDS_DTYPE
    options = dict(_RECORDS_OPTIONS)

    def resolve_depends_on(self, context: Any, run_id: str | None = None) -> list[str]:
        return ["raw_files"]

    def compute(self, context: Any, run_id: str, **kwargs) -> np.ndarray:
        return build_bundle(context, run_id, self)[0]


class B200WavePoolPlugin(Plugin):
    provides = "wave_pool"
    depends_on = []
    description = "Build wave_pool from the shared internal records bundle."
    version = "0.10.0"
    save_when = "always"
    uses_run_config = True
    output_dtype = np.dtype(np.uint16)
    options = dict(_RECORDS_OPTIONS)

    def resolve_depends_on(self, context: Any, run_id: str | None = None) -> list[str]:
        return ["raw_files"]

    def compute(self, context: Any, run_id: str, **kwargs) -> np.ndarray:
        return build_bundle(context, run_id, self)[1]
