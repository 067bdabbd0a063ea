"""CPU: libwfb200.so loads and exports every function include/wfb200.h declares."""

import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "wfb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wfb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    names = declared_functions()
    assert "wfb_features_hits" in names and "wfb_process_host" in names and len(names) >= 15


def test_library_exports_every_declared_symbol():
    from waveformanalysis_b200 import _lib, build

    build.build()
    lib = _lib.load()
    names = declared_functions()
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in wfb200.h but not exported: {missing}"
    assert set(names) == set(_lib.PROTOTYPES), set(names) ^ set(_lib.PROTOTYPES)
    assert lib.wfb_version() >= 100


def test_struct_layouts_match_header():
    import ctypes as C

    from waveformanalysis_b200 import _lib

    assert C.sizeof(_lib.RecMeta) == 48 and _lib.RecMeta.record_id.offset == 40
    assert C.sizeof(_lib.ChanRule) == 32
    assert _lib.FHParams.rules_dev.offset == 64 and C.sizeof(_lib.FHParams) == 96
    assert C.sizeof(_lib.FilterCfg) == 16 + 16 * 6 * 8 + 16 * 2 * 8


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np

    from waveformanalysis_b200 import engine
    from waveformanalysis_b200.dtypes import RECORDS_DTYPE

    with pytest.raises(RuntimeError, match="no CUDA device"):
        engine.process_host(np.zeros(1, RECORDS_DTYPE), np.zeros(4, np.uint16))


def test_v1725_host_scan_matches_oracle():
    """wfb_v1725_scan_host is a host function (no GPU needed): same index as the oracle's header walk,
    including streams cut in the middle of an event header, a channel header or a payload."""
    import numpy as np

    from oracle import np_oracle as O
    from waveformanalysis_b200 import ops
    from waveformanalysis_b200.synth import make_v1725_blob

    blob = make_v1725_blob(n_events=60, n_channels=16, seed=5, lengths=(2, 90), tie_every=4)
    for cut in (0, 1, 7, 15, 16, 17, 27, 28, 29, 31, 200, 1001, len(blob) - 1, len(blob)):
        part = blob[:cut]
        got, want = ops.v1725_scan(part), O.v1725_scan(part)
        for k in ("payload_offset", "n_samples", "channel", "timestamp", "baseline", "trunc"):
            assert np.array_equal(got[k], want[k]), (cut, k)
        assert got["n_samples_total"] == int(want["n_samples"].sum())
