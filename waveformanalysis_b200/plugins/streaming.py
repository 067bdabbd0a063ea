"""The streaming side of the path on the B200 (reference: core/plugins/core/streaming.py:119-548 and
core/plugins/builtin/streaming/cpu/signal_peaks.py:35-406).

The reference's ``StreamingPlugin.compute`` walks a chunk iterator and calls ``compute_chunk`` for one chunk after the
other (optionally on a thread / process pool).  On the GPU the parallel resource is the device, so the backend here keeps
the protocol - chunks with ``main_start`` / ``main_end`` / ``segment_id`` metadata in, ``Chunk`` objects clipped to the
main range out, ``compute_chunk`` still callable on its own - and replaces the executor by a device pipeline of three slots:

    chunks k + 1, k + 2:  host rows + sample range -> H2D on the copy stream, kernels queued behind it   (``begin_chunk``)
    chunk k            :  kernels on the compute stream, rows D2H, clip to the main range, yield        (``end_chunk``)

``GpuStreamingMixin`` holds that pipeline.  Two plugins use it:

* ``B200SignalPeaksStreamPlugin``  (``signal_peaks_stream``): the reference's chunk iteration (per channel, dt segment,
  time break) with ``wfb_find_peaks`` per chunk instead of a Python loop over waveforms.
* ``B200HitThresholdStreamPlugin`` (``hit_threshold_stream``): records + wave_pool in time-ordered chunks with halo,
  the fused baseline -> hits -> basic_features pass (K2/K3) per chunk, hit rows clipped to the chunk's main records.

When the reference package is importable the classes derive from its ``StreamingPlugin`` (configuration through
``streaming_config``, ``Chunk`` validation); without it a small stand-in with the same attributes is used."""

from __future__ import annotations

from collections import deque
from types import SimpleNamespace
from typing import Any, Iterator

import numpy as np

from .. import engine, ops
from .. import _lib
from ..channel_config import per_channel_option
from ..dtypes import BASIC_FEATURES_DTYPE, HIT_DTYPE, RECORDS_DTYPE, THRESHOLD_HIT_DTYPE
from ..plugin_api import HAVE_REFERENCE, Option, Plugin, resolve_dt_config
from ..wave_source import WAVE_SOURCE_AUTO, load_wave_input, resolve_wave_input_spec

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

DEFAULT_BREAK_THRESHOLD_PS = 10_000_000_000_000  # core/processing/chunk.py:51

if HAVE_REFERENCE:
    from waveform_analysis.core.plugins.builtin.streaming.cpu.signal_peaks import SignalPeaksStreamPlugin as _PeaksBase  # type: ignore
    from waveform_analysis.core.plugins.core.streaming import StreamingPlugin as _StreamBase  # type: ignore
    from waveform_analysis.core.processing.chunk import TIMESTAMP_FIELD, Chunk  # type: ignore
else:
    TIMESTAMP_FIELD = "timestamp"

    def Chunk(**kw):  # noqa: N802 - stands in for core/processing/chunk.py:78-207
        kw.setdefault("metadata", {})
        return SimpleNamespace(**kw)

    class _StreamBase(Plugin):  # the attributes of StreamingPlugin the backend reads (streaming.py:154-176)
        chunk_size = 50000
        parallel = False
        time_field = TIMESTAMP_FIELD
        dt_field = "dt"
        length_field = "length"
        endtime_field = "endtime"
        dt = None
        output_time_field = TIMESTAMP_FIELD
        output_endtime_field = "endtime"
        output_data_kind = "stream"
        output_kind = "stream"
        required_halo_ns = 0
        required_halo_left_ns = 0
        required_halo_right_ns = 0
        clip_strict = False
        break_threshold_ps = DEFAULT_BREAK_THRESHOLD_PS

        def _get_required_halo(self) -> tuple[int, int]:  # streaming.py:318-324
            left, right = self.required_halo_left_ns or 0, self.required_halo_right_ns or 0
            if self.required_halo_ns:
                left, right = max(left, self.required_halo_ns), max(right, self.required_halo_ns)
            return int(left), int(right)

        def _postprocess_result(self, result: Any, input_chunk: Any):  # streaming.py:380-445 for instantaneous rows
            if result is None:
                return None
            main_start, main_end = input_chunk.metadata.get("main_start"), input_chunk.metadata.get("main_end")
            if not hasattr(result, "data"):
                result = Chunk(data=np.asarray(result), start=main_start, end=main_end, run_id=input_chunk.run_id, data_type=self.provides,
                               data_kind=self.output_data_kind, time_field=self.output_time_field,
                               metadata={"segment_id": input_chunk.metadata.get("segment_id")})
            if main_start is None or main_end is None or getattr(result.data, "dtype", None) is None or result.data.dtype.names is None:
                return result
            t = result.data[result.time_field]
            keep = (t >= main_start) & (t <= main_end) if self.clip_strict else (t > main_start) & (t < main_end)  # chunk.py:657-670
            if not keep.any():
                return None
            result.data = result.data[keep]
            result.start, result.end = int(main_start), int(main_end)
            result.metadata.update(main_start=main_start, main_end=main_end, segment_id=input_chunk.metadata.get("segment_id"))
            return result

        def _validate_chunk(self, chunk: Any) -> None:
            return None

    _PeaksBase = _StreamBase


class GpuStreamingMixin:
    """Device pipeline of ``pipeline_depth`` slots behind the StreamingPlugin protocol.  A subclass provides

    * ``begin_chunk(chunk, slots, context, run_id, **kw)`` - stage the chunk (``slots.stage``) and enqueue its kernels on
      ``slots.compute_stream``; returns a job object, or None to drop the chunk.  Nothing may wait for the device here.
    * ``end_chunk(job, chunk, context, run_id)`` - wait for the job and return the chunk's rows (or a Chunk, or None)."""

    pipeline_depth = 3  # device slots: one chunk computing / draining, two uploading or queued behind it
    parallel = False  # no executor: chunks overlap on the device, results come out in order
    executor_type = "thread"

    def _direct_slots(self) -> "engine.StreamSlots":
        slots = getattr(self, "_slots", None)
        if slots is None:
            slots = self._slots = engine.StreamSlots(self.pipeline_depth)
        return slots

    def compute_chunk(self, chunk: Any, context: Any, run_id: str, **kwargs):
        """One chunk, synchronously (the protocol's entry point; ``compute`` overlaps the chunks instead)."""
        job = self.begin_chunk(chunk, self._direct_slots(), context, run_id, **kwargs)
        return None if job is None else self.end_chunk(job, chunk, context, run_id)

    def compute(self, context: Any, run_id: str, show_progress: bool = False, progress_desc: str | None = None, **kwargs):
        if hasattr(self, "_collect_streaming_config"):  # the reference's per-run streaming configuration (streaming.py:469-477)
            streaming_config = kwargs.pop("streaming_config", None)
            self._apply_streaming_config(self._collect_streaming_config(context, streaming_config))
        kwargs.pop("executor_config", None)
        chunks = self._get_input_chunks(context, run_id, **kwargs)
        return self._pipeline(chunks, context, run_id, **kwargs)

    def _make_slots(self) -> "engine.StreamSlots":
        return engine.StreamSlots(self.pipeline_depth)

    def _pipeline(self, chunks: Iterator[Any], context: Any, run_id: str, **kwargs):
        slots = self._make_slots()
        self.stream_stats = {"chunks": 0, "bytes_uploaded": 0, "overlapped_chunks": 0}
        pending: deque = deque()  # chunks on the device, oldest first (at most pipeline_depth - 1 besides the one just begun)

        def emit(item):
            job, chunk = item
            result = None if job is None else self.end_chunk(job, chunk, context, run_id)
            result = self._postprocess_result(result, chunk)
            if result is not None:
                self._validate_chunk(result)
            return result

        try:
            for chunk in chunks:
                job = self.begin_chunk(chunk, slots, context, run_id, **kwargs)  # upload + launch of the newest chunk ...
                if job is not None and any(j is not None for j, _ in pending):
                    self.stream_stats["overlapped_chunks"] += 1
                pending.append((job, chunk))
                if len(pending) >= max(2, int(self.pipeline_depth)):
                    out = emit(pending.popleft())                                 # ... while the oldest one finishes
                    if out is not None:
                        yield out
            while pending:
                out = emit(pending.popleft())
                if out is not None:
                    yield out
        finally:
            if slots is not None:
                slots.close()  # also when the consumer stops early or a chunk raised: nothing stays in flight
                self.stream_stats.update(chunks=slots.chunks, bytes_uploaded=slots.bytes_uploaded)


class B200SignalPeaksStreamPlugin(GpuStreamingMixin, _PeaksBase):
    """Streaming peak detection: one device launch per chunk instead of a Python loop over waveforms."""

    provides = "signal_peaks_stream"
    depends_on = ["filtered_waveforms", "st_waveforms"]
    description = "Stream peak detection from filtered waveforms."
    version = "1.2.0"
    save_when = "never"
    output_dtype = None
    if not HAVE_REFERENCE:
        options = dict(_OPTIONS)
        output_data_kind = "peaks"

        def _get_input_chunks(self, context: Any, run_id: str, **kwargs):
            raise RuntimeError("signal_peaks_stream: the reference's chunk iteration (signal_peaks.py:113-232) is needed to walk a run; "
                               "call compute_chunk on your own chunks instead")

    def _load_config(self, context: Any) -> None:
        if HAVE_REFERENCE:
            super()._load_config(context)
            return
        for k in ("use_derivative", "height", "distance", "prominence", "width", "threshold", "height_method"):
            setattr(self, k, context.get_config(self, k))
        self.minmax_window_expand = max(0, int(context.get_config(self, "minmax_window_expand")))
        self.explicit_dt = resolve_dt_config(context, self, deprecated_keys=("sampling_interval_ns", "dt_ns"))

    def compute(self, context: Any, run_id: str, **kwargs):
        self._load_config(context)
        return GpuStreamingMixin.compute(self, context, run_id, **kwargs)

    def peak_options(self) -> dict:
        """find_peaks options of this run, under the names of the reference's ``_find_peaks_in_waveform`` call."""
        return dict(use_derivative=bool(self.use_derivative), height=float(self.height), distance=int(self.distance),
                    prominence=float(self.prominence), width=int(self.width),
                    threshold=None if self.threshold is None else float(self.threshold), height_method=str(self.height_method),
                    minmax_window_expand=int(self.minmax_window_expand))

    def begin_chunk(self, chunk: Any, slots: "engine.StreamSlots", context: Any, run_id: str, **kwargs):
        import torch

        st_chunk = chunk.data
        filtered_chunk = chunk.metadata.get("filtered_waveforms")
        if filtered_chunk is None or len(st_chunk) == 0:
            return None
        if not hasattr(self, "height_method"):
            self._load_config(context)
        if self.height_method not in ("minmax", "diff"):
            raise ValueError(f"不支持的峰高计算方法: {self.height_method}")  # the reference's message (peak_finding.py:611)
        inp = ops.peaks_stream_inputs(st_chunk, filtered_chunk, explicit_dt=self.explicit_dt, event_offset=int(chunk.metadata.get("event_offset", 0)),
                                      use_derivative=bool(self.use_derivative))
        if inp is None:
            return None
        rec, pool = inp
        lmax = int(rec["event_length"].max())
        if lmax <= 0:
            return None
        st = slots.stage(rec, pool)
        with torch.cuda.stream(slots.compute_stream):
            run = slots.device_run(st, lmax=lmax)
            opts = self.peak_options()
            opts["height_window_extension"] = opts.pop("minmax_window_expand")
            st.out = ops.find_peaks_launch(run, _lib.WAVE_AOS_F32_AS_F64, lmax, cumsum_diff=True, **opts)
            slots.mark_done(st)
        return (slots, st)

    def end_chunk(self, job, chunk: Any, context: Any, run_id: str):
        import torch

        slots, st = job
        with torch.cuda.stream(slots.compute_stream):
            peaks = ops.find_peaks_collect(st.out)
        st.out = {}
        if len(peaks) == 0:
            return None
        return Chunk(data=peaks, start=int(np.min(peaks["timestamp"])), end=int(np.max(peaks["timestamp"])), run_id=run_id,
                     data_type=self.provides, data_kind=self.output_data_kind, time_field=TIMESTAMP_FIELD)


class B200HitThresholdStreamPlugin(GpuStreamingMixin, _StreamBase):
    """``hit_threshold`` (and ``basic_features``) as a stream over time-ordered record chunks.

    Input: ``records`` + ``wave_pool`` (or ``wave_pool_filtered``), as the non-streaming plugins read them
    (``_wave_source.py:168-229``).  Chunks are ``chunk_size`` consecutive records of one time segment (segments end where
    the gap to the next record exceeds ``break_threshold_ps``, chunk.py:857-930), extended by the halo on both sides; per
    chunk the fused kernel computes the hit rows and the feature rows.  Output chunks carry the THRESHOLD_HIT rows of the
    chunk's MAIN records (rows of halo records are dropped: every hit of the run appears in exactly one chunk, which a
    clip of instantaneous rows to the open interval ``(main_start, main_end)`` - chunk.py:666-670 - would not guarantee)
    and, in ``metadata["basic_features"]``, the BASIC_FEATURES rows of the same records.  Concatenating the chunks gives
    the arrays of ``B200ThresholdHitPlugin`` / ``B200BasicFeaturesPlugin`` row for row."""

    provides = "hit_threshold_stream"
    depends_on = []
    description = "Threshold hits and basic features per time chunk of records, on the device."
    version = "0.1.0"
    save_when = "never"
    output_dtype = None
    output_data_kind = "hits"
    chunk_size = 262144
    time_field = TIMESTAMP_FIELD
    length_field = "event_length"
    dt_field = "dt"
    options = {
        "threshold": Option(default=10.0, type=float, help="hit threshold"),
        "use_filtered": Option(default=False, type=bool, help="use the filtered waveform source"),
        "wave_source": Option(default=WAVE_SOURCE_AUTO, type=str, help="auto|records (the stream reads the records source)"),
        "left_extension": Option(default=2, type=int, help="samples added left of the threshold region"),
        "right_extension": Option(default=2, type=int, help="samples added right of the threshold region"),
        "channel_config": Option(default=None, type=dict, help="per (board, channel) overrides: threshold, fixed_baseline"),
        "height_range": Option(default=(40, 90), help="sample range of the feature height"),
        "area_range": Option(default=(0, None), help="sample range of the feature area"),
        "with_features": Option(default=True, type=bool, help="also compute the BASIC_FEATURES rows of every chunk"),
    }

    def resolve_depends_on(self, context: Any, run_id: str | None = None) -> list[str]:
        return list(resolve_wave_input_spec(context, self).depends_on)

    # ---- chunk iteration ----------------------------------------------------------------------------------------
    def _get_input_chunks(self, context: Any, run_id: str, **kwargs) -> Iterator[Any]:
        wave_input = load_wave_input(context, self, run_id, needs_wave_samples=True)
        if not wave_input.spec.is_records:
            raise ValueError("hit_threshold_stream reads the records source (records + wave_pool)")
        records, pool = wave_input.records, wave_input.wave_pool
        if records is None or pool is None:
            raise ValueError("hit_threshold_stream failed to load records_view for records source")
        self._pool = pool
        self._run_cfg = self._load_run_config(context, run_id, records)
        return self._record_chunks(records, run_id)

    def _load_run_config(self, context: Any, run_id: str, records: np.ndarray) -> dict:
        threshold = float(context.get_config(self, "threshold"))
        names = records.dtype.names or ()
        for need in ("timestamp", "dt", "event_length", "wave_offset"):
            if need not in names:
                raise ValueError(f"[{self.provides}] records is missing required field '{need}'")
        boards = records["board"] if "board" in names else np.zeros(len(records), np.int16)
        channels = records["channel"] if "channel" in names else np.zeros(len(records), np.int16)
        cc = context.get_config(self, "channel_config")
        thr = per_channel_option(cc, run_id, boards, channels, "threshold", threshold)
        fixed = per_channel_option(cc, run_id, boards, channels, "fixed_baseline", None)
        lmax = engine.records_host_scan(records)["lmax"] if records.dtype == RECORDS_DTYPE else int(np.asarray(records["event_length"]).max(initial=0))
        return dict(threshold=threshold, left_extension=max(0, int(context.get_config(self, "left_extension"))),
                    right_extension=max(0, int(context.get_config(self, "right_extension"))),
                    rules=engine.make_rules({k: float(v) for k, v in thr.items() if float(v) != threshold},
                                            {k: float(v) for k, v in fixed.items() if v is not None}),
                    height_range=tuple(context.get_config(self, "height_range")), area_range=tuple(context.get_config(self, "area_range")),
                    with_features=bool(context.get_config(self, "with_features")),
                    lmax=lmax)  # the padded width is the run's, not the chunk's (hit_finder.py:364)

    def _record_chunks(self, records: np.ndarray, run_id: str) -> Iterator[Any]:
        n = len(records)
        if n == 0:
            return
        if records.dtype == RECORDS_DTYPE:  # one threaded pass in the library instead of a strided numpy pass per field
            scan = engine.records_host_scan(records, times=True)
            ts, end, dt_min = scan["ts"], scan["end"], scan["dt_min"]
        else:
            ts = np.asarray(records["timestamp"]).astype(np.int64)
            dt = np.asarray(records["dt"]).astype(np.int64)
            dt_min = int(dt.min())
            end = ts + np.maximum(np.asarray(records["event_length"]).astype(np.int64), 0) * dt * 1000  # ps (chunk.py:345-385)
        if dt_min <= 0:
            raise ValueError(f"[{self.provides}] records.dt must be positive for every row")
        run_end = np.maximum.accumulate(end)
        bounds = [0, n]
        if self.break_threshold_ps and self.break_threshold_ps > 0 and n > 1:
            gaps = ts[1:] - run_end[:-1]
            bounds = [0, *(np.flatnonzero(gaps > self.break_threshold_ps) + 1).tolist(), n]
        halo_l, halo_r = (int(h) * 1000 for h in self._get_required_halo())  # ns -> the ps of the time field
        size = max(1, int(self.chunk_size))
        for segment_id, (s0, s1) in enumerate(zip(bounds[:-1], bounds[1:])):
            seg_start, seg_end = int(ts[s0:s1].min()), int(end[s0:s1].max())
            for i in range(s0, s1, size):
                j = min(s1, i + size)
                main_start, main_end = int(ts[i:j].min()), int(end[i:j].max())
                lo, hi = i, j
                if halo_l or halo_r:
                    ext_start, ext_end = max(seg_start, main_start - halo_l), min(seg_end, main_end + halo_r)
                    touch = np.flatnonzero((end[s0:s1] > ext_start) & (ts[s0:s1] < ext_end))  # select_time_range, non-strict
                    lo, hi = min(i, s0 + int(touch[0])), max(j, s0 + int(touch[-1]) + 1)
                    # a record that only reaches into the halo starts before it: the chunk spans its rows (chunk.py:130-150)
                    ext_start, ext_end = min(ext_start, int(ts[lo:hi].min())), max(ext_end, int(end[lo:hi].max()))
                else:
                    ext_start, ext_end = main_start, main_end
                yield Chunk(data=records[lo:hi], start=ext_start, end=ext_end, run_id=run_id, data_type=self.provides,
                            time_field=TIMESTAMP_FIELD, dt_field="dt", length_field="event_length",
                            metadata={"main_start": main_start, "main_end": main_end, "segment_id": segment_id,
                                      "row_base": lo, "main_rows": (i - lo, j - lo)})

    # ---- device pipeline ----------------------------------------------------------------------------------------
    def begin_chunk(self, chunk: Any, slots: "engine.StreamSlots", context: Any, run_id: str, **kwargs):
        import torch

        rows = chunk.data
        if len(rows) == 0:
            return None
        cfg = getattr(self, "_run_cfg", None)
        pool = chunk.metadata.get("wave_pool", getattr(self, "_pool", None))
        if cfg is None or pool is None:  # compute_chunk called on its own: configuration from the chunk's rows
            wave_input = load_wave_input(context, self, run_id, needs_wave_samples=True)
            pool = self._pool = wave_input.wave_pool
            cfg = self._run_cfg = self._load_run_config(context, run_id, wave_input.records)
        rec = engine.packed_records(rows, None)
        scan = engine.records_host_scan(rec)
        lo, hi = scan["lo"], scan["hi"]
        if lo < 0 or hi > len(pool):
            raise ValueError("records reference samples outside wave_pool bounds")
        lo_al = lo & ~7  # record starts keep their 16-byte phase relative to the slot
        st = slots.stage(rec, pool[lo_al:hi])
        with torch.cuda.stream(slots.compute_stream):
            run = slots.device_run(st, lmax=cfg["lmax"], pool_base=lo_al, row_base=int(chunk.metadata.get("row_base", 0)))
            res = run.features_hits(features=cfg["with_features"], hits=True, threshold=cfg["threshold"], left_extension=cfg["left_extension"],
                                    right_extension=cfg["right_extension"], height_range=cfg["height_range"], area_range=cfg["area_range"],
                                    rules=cfg["rules"] if len(cfg["rules"]) else None)
            total_h = torch.empty(1, dtype=torch.int64, pin_memory=True)
            total_h.copy_(res["total"], non_blocking=True)
            st.out = dict(res=res, total_h=total_h)
            slots.mark_done(st)
        return (slots, st, cfg)

    def end_chunk(self, job, chunk: Any, context: Any, run_id: str):
        import torch

        slots, st, cfg = job
        slots.finish(st)
        res, run = st.out["res"], st.run
        with torch.cuda.stream(slots.compute_stream):
            run.check()
            total = int(st.out["total_h"][0])
            if total > res["cap"]:  # more hits than the 8-per-record buffer: once more with room
                res = run.features_hits(features=False, hits=True, threshold=cfg["threshold"], left_extension=cfg["left_extension"],
                                        right_extension=cfg["right_extension"], rules=cfg["rules"] if len(cfg["rules"]) else None, hit_cap=total)
                torch.cuda.current_stream().synchronize()
            hits = engine.to_host(res["hits"][: total * 60]).view(THRESHOLD_HIT_DTYPE)
            feats = engine.to_host(st.out["res"]["features"][: st.n * 36]).view(BASIC_FEATURES_DTYPE) if cfg["with_features"] else None
            torch.cuda.current_stream().synchronize()
        st.out = {}
        return self.main_rows_chunk(hits, feats, chunk, run_id)

    def main_rows_chunk(self, hits: np.ndarray, feats: np.ndarray | None, chunk: Any, run_id: str):
        """The output chunk: rows of the input chunk's main records (halo records belong to the neighbouring chunks)."""
        n = len(chunk.data)
        m0, m1 = chunk.metadata.get("main_rows", (0, n))
        if (m0, m1) != (0, n):
            if "record_id" in (chunk.data.dtype.names or ()):
                hits = hits[np.isin(hits["record_id"], np.asarray(chunk.data["record_id"])[m0:m1])]
            if feats is not None:
                feats = feats[m0:m1]
        return Chunk(data=hits, start=int(chunk.metadata.get("main_start", chunk.start)), end=int(chunk.metadata.get("main_end", chunk.end)),
                     run_id=run_id, data_type=self.provides, data_kind=self.output_data_kind, time_field=TIMESTAMP_FIELD, dt_field="dt",
                     metadata={"segment_id": chunk.metadata.get("segment_id"), "main_start": chunk.metadata.get("main_start"),
                               "main_end": chunk.metadata.get("main_end"), "basic_features": feats, "n_records": int(m1 - m0)})

    def _postprocess_result(self, result: Any, input_chunk: Any):
        # the rows are those of the chunk's main records already (see the class docstring); empty chunks are kept when
        # they carry feature rows
        if result is None:
            return None
        if len(result.data) == 0 and result.metadata.get("basic_features") is None:
            return None
        return result

    def _validate_chunk(self, chunk: Any) -> None:
        return None  # hit timestamps of one chunk are record-major, not monotonic (hit_finder.py:352-355)
