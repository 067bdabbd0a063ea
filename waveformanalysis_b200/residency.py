"""Device residency across plugin calls of one Context run.

The reference shares one in-memory records bundle between its ``records`` / ``wave_pool`` plugins and every
downstream plugin (core/plugins/builtin/cpu/records.py:209-305, 441-464 ``get_records_bundle``).  The B200
counterpart keeps the DEVICE copies: the first plugin that needs ``records`` + ``wave_pool`` (or
``wave_pool_filtered``) of a run uploads them once, every later plugin of the run finds them in HBM.  The
entry point is ``device_run``; results that one fused pass produces for a sibling plugin (``basic_features``
and ``hit_threshold`` come out of the same kernel) wait in ``put_rows`` / ``take_rows``.

An entry is keyed by (run_id, data name) and validated by a fingerprint of the host arrays the Context hands
over: address, size, dtype and a hash of evenly spaced probes of the contents, so that a result recomputed
under another configuration (or a different array under the same name) is never confused with the resident
copy.  A Context replaces results by ``np.memmap`` views of its cache files (core/context_execution.py:241-251):
the address changes, the probes do not, so the fingerprint compares contents when the address differs.

Nothing here computes; it only owns device tensors.  ``release()`` frees them (also reached through
``wfb_release_cache`` users: ``waveformanalysis_b200.release_device_cache``).
"""

from __future__ import annotations

import hashlib
import threading
from collections import OrderedDict

import numpy as np

_LOCK = threading.RLock()
_RUNS: "OrderedDict[tuple, dict]" = OrderedDict()   # (run_id, pool_name, signed) -> {fp, run, bytes}
_ROWS: "OrderedDict[tuple, dict]" = OrderedDict()   # (run_id, data_name) -> {sig, rows}
_N_PROBES = 128
_PROBE_BYTES = 64
MAX_FRACTION = 0.70  # of the device memory the resident pools may take before the oldest entries are dropped
STATS = {"uploads": 0, "hits": 0, "row_hits": 0, "bytes_uploaded": 0}


def _probe_hash(a: np.ndarray) -> str:
    """Hash of the array's size, dtype and _N_PROBES evenly spaced 64-byte probes (first and last block included)."""
    b = np.ascontiguousarray(a).view(np.uint8).reshape(-1) if a.size else np.zeros(0, np.uint8)
    h = hashlib.blake2b(digest_size=16)
    h.update(str((a.dtype.str if a.dtype.names is None else a.dtype.descr, a.shape)).encode())
    n = b.size
    if n <= _N_PROBES * _PROBE_BYTES:
        h.update(b.tobytes())
    else:
        step = (n - _PROBE_BYTES) // (_N_PROBES - 1)
        for k in range(_N_PROBES):
            o = k * step
            h.update(b[o:o + _PROBE_BYTES].tobytes())
    return h.hexdigest()


def fingerprint(*arrays: np.ndarray) -> tuple:
    return tuple((int(a.ctypes.data) if a.size else 0, int(a.nbytes), _probe_hash(a)) for a in arrays)


def _same(fp_a: tuple, fp_b: tuple) -> bool:
    """Equal sizes and probe hashes (the address is informational: memmap views of the same file differ in it)."""
    return len(fp_a) == len(fp_b) and all(x[1:] == y[1:] for x, y in zip(fp_a, fp_b))


_TOTAL: dict = {}


def _device_total() -> int:
    """HBM size of the current device (cudaMemGetInfo costs ~3 ms a call: asked once per device)."""
    import torch

    dev = torch.cuda.current_device()
    if dev not in _TOTAL:
        _TOTAL[dev] = int(torch.cuda.get_device_properties(dev).total_memory)
    return _TOTAL[dev]


def _device_budget() -> int:
    return int(_device_total() * MAX_FRACTION)


def _evict_for(nbytes: int) -> None:
    import torch

    budget = _device_budget()
    used = sum(e["bytes"] for e in _RUNS.values())
    while _RUNS and used + nbytes > budget:
        _, e = _RUNS.popitem(last=False)
        used -= e["bytes"]
    # memory torch has reserved but not handed out is reusable; only what OTHER users of the device hold is not
    reserved = torch.cuda.memory_reserved()
    inside = reserved - torch.cuda.memory_allocated()
    if inside < nbytes + (1 << 28) and _device_total() - reserved < nbytes + (1 << 30):
        free, _ = torch.cuda.mem_get_info()
        if free < nbytes + (1 << 30):
            torch.cuda.empty_cache()


def fits_device(nbytes: int) -> bool:
    """Whether a pool of ``nbytes`` can be made resident at all (else the chunked host pipeline is used)."""
    return nbytes < _device_total() * 0.45


def device_run(run_id: str, records: np.ndarray, pool: np.ndarray, pool_name: str, *, explicit_dt=None, signed: bool = False,
               clamp_lengths=None):
    """The DeviceRun (records metadata + sample pool in HBM) of (run_id, pool_name), uploaded on first use."""
    from . import engine

    key = (str(run_id), str(pool_name), bool(signed), None if clamp_lengths is None else _probe_hash(np.asarray(clamp_lengths)))
    fp = fingerprint(records, pool)
    with _LOCK:
        ent = _RUNS.get(key)
        if ent is not None and _same(ent["fp"], fp) and ent["explicit_dt"] == explicit_dt:
            _RUNS.move_to_end(key)
            STATS["hits"] += 1
            return ent["run"]
        if ent is not None:
            del _RUNS[key]
        nbytes = int(pool.nbytes + len(records) * (102 + 48))
        _evict_for(nbytes)
        run = engine.DeviceRun.from_host(records, pool, explicit_dt=explicit_dt, clamp_lengths=clamp_lengths)
        _RUNS[key] = {"fp": fp, "run": run, "bytes": nbytes, "explicit_dt": explicit_dt}
        STATS["uploads"] += 1
        STATS["bytes_uploaded"] += nbytes
        return run


def lookup_run(run_id: str, records: np.ndarray, pool: np.ndarray, pool_name: str, fp: tuple | None = None):
    """The resident DeviceRun of (run_id, pool_name) if it was made from these arrays, else None."""
    key = (str(run_id), str(pool_name), False, None)
    fp = fp or fingerprint(records, pool)
    with _LOCK:
        ent = _RUNS.get(key)
        if ent is not None and _same(ent["fp"], fp) and ent["explicit_dt"] is None:
            _RUNS.move_to_end(key)
            STATS["hits"] += 1
            return ent["run"]
    return None


def store_run(run_id: str, records: np.ndarray, pool: np.ndarray, pool_name: str, run, fp: tuple | None = None) -> None:
    """Register a run that a host pipeline call left resident (one upload, counted as such)."""
    key = (str(run_id), str(pool_name), False, None)
    nbytes = int(pool.nbytes + len(records) * (102 + 48))
    with _LOCK:
        _RUNS.pop(key, None)
        _evict_for(0)
        _RUNS[key] = {"fp": fp or fingerprint(records, pool), "run": run, "bytes": nbytes, "explicit_dt": None}
        STATS["uploads"] += 1
        STATS["bytes_uploaded"] += nbytes


def adopt_run(run_id: str, records: np.ndarray, pool: np.ndarray, pool_name: str, run, *, signed: bool = False) -> None:
    """Register a DeviceRun whose pool was PRODUCED on the device (records builder, filter) under the host arrays
    that were copied back from it, so that the next plugin finds the device copy instead of uploading the host one."""
    key = (str(run_id), str(pool_name), bool(signed), None)
    nbytes = int(pool.nbytes + len(records) * (102 + 48))
    with _LOCK:
        _RUNS.pop(key, None)
        _evict_for(nbytes)
        _RUNS[key] = {"fp": fingerprint(records, pool), "run": run, "bytes": nbytes, "explicit_dt": None}


def put_rows(run_id: str, data_name: str, signature: tuple, rows: np.ndarray) -> None:
    with _LOCK:
        _ROWS[(str(run_id), str(data_name))] = {"sig": signature, "rows": rows}
        while len(_ROWS) > 8:
            _ROWS.popitem(last=False)


def take_rows(run_id: str, data_name: str, signature: tuple):
    """Rows a sibling plugin's fused pass left for (run_id, data_name), if they were computed from the same inputs
    with the same configuration; the entry is consumed."""
    with _LOCK:
        ent = _ROWS.get((str(run_id), str(data_name)))
        if ent is None or ent["sig"] != signature:
            return None
        del _ROWS[(str(run_id), str(data_name))]
        STATS["row_hits"] += 1
        return ent["rows"]


def release(run_id: str | None = None) -> None:
    """Drop the resident copies (of one run, or all) and hand the memory back to the driver."""
    with _LOCK:
        for d in (_RUNS, _ROWS):
            for k in [k for k in d if run_id is None or k[0] == str(run_id)]:
                del d[k]
    if run_id is not None:
        return  # the blocks stay with torch's caching allocator: the next run of the same size reuses them
    try:
        import torch

        if torch.cuda.is_available():
            torch.cuda.empty_cache()
    except Exception:
        pass
