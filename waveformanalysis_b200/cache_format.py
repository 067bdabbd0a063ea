"""Write a result array straight into the reference's on-disk cache layout, so a Context finds it on the
next ``get_data`` without the save -> reload round trip (core/storage/memmap.py:175-203 paths, :303-311
metadata file, :313-423 metadata keys, :528-613 what the loader reads back).

Layout: ``work_dir/{run_id}/_cache/{key}.bin`` (the rows, raw) + ``{key}.json`` (count, dtype, itemsize,
storage_version, timestamp, shape, compressed, dtype_descr for structured dtypes, caller's extra metadata such
as the lineage).  Uncompressed, no checksum - the defaults of MemmapStorage.  Files appear atomically
(temporary name + rename), the metadata last: a reader never sees a half-written entry.
"""

from __future__ import annotations

import json
import os
import time
from typing import Any

import numpy as np

STORAGE_VERSION = "1.0.1"  # memmap.py:81


def cache_paths(work_dir: str, run_id: str, key: str, data_subdir: str = "_cache") -> tuple:
    root = os.path.join(work_dir, run_id, data_subdir)
    return os.path.join(root, f"{key}.bin"), os.path.join(root, f"{key}.json")


def write_cache_entry(work_dir: str, run_id: str, key: str, rows: np.ndarray, extra_metadata: dict | None = None,
                      data_subdir: str = "_cache") -> str:
    """Store ``rows`` (any contiguous numpy array, typically packed result rows copied back from the device) under
    ``key``.  Returns the path of the binary file.  An empty array writes nothing, like finalize_save (:326, :419)."""
    rows = np.ascontiguousarray(rows)
    bin_path, meta_path = cache_paths(work_dir, run_id, key, data_subdir)
    os.makedirs(os.path.dirname(bin_path), exist_ok=True)
    if rows.shape[0] == 0:
        return bin_path
    tmp = bin_path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(memoryview(rows).cast("B"))
    os.replace(tmp, bin_path)
    dt = rows.dtype
    meta: dict[str, Any] = {
        "count": int(rows.shape[0]),
        "dtype": dt.str,
        "itemsize": int(dt.itemsize),
        "storage_version": STORAGE_VERSION,
        "timestamp": time.time(),
        "shape": list(rows.shape),
        "compressed": False,
    }
    if dt.names is not None:
        meta["dtype_descr"] = dt.descr
    if extra_metadata:
        meta.update(extra_metadata)
    tmp = meta_path + ".tmp"
    with open(tmp, "w") as f:
        json.dump(meta, f, default=str)
    os.replace(tmp, meta_path)
    return bin_path


def read_cache_entry(work_dir: str, run_id: str, key: str, data_subdir: str = "_cache") -> np.ndarray | None:
    """Memory-map an entry written by either side (uncompressed entries only); None when absent."""
    bin_path, meta_path = cache_paths(work_dir, run_id, key, data_subdir)
    if not (os.path.exists(bin_path) and os.path.exists(meta_path)):
        return None
    with open(meta_path) as f:
        meta = json.load(f)
    if meta.get("compressed"):
        raise ValueError(f"cache entry {key} is compressed; read it through the reference's MemmapStorage")
    if "dtype_descr" in meta:
        dt = np.dtype([tuple(tuple(x) if isinstance(x, list) else x for x in item) for item in meta["dtype_descr"]])
    else:
        dt = np.dtype(meta["dtype"])
    shape = tuple(meta.get("shape") or (meta["count"],))
    return np.memmap(bin_path, dtype=dt, mode="r", shape=shape)
