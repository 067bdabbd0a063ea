import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden", "hotpath_golden.npz")

# floating-point bar from BASELINE.json north_star: rel 1e-5 / abs 1e-3 ADC*sample
RTOL = 1e-5
ATOL = 1e-3


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(GOLDEN, allow_pickle=False)


def assert_rows_match(got: np.ndarray, want: np.ndarray, *, float_exact=(), rtol=RTOL, atol=ATOL, what=""):
    """Structured-array comparison: integer/str fields bit-exact, float fields within tolerance."""
    assert got.dtype == want.dtype, f"{what}: dtype {got.dtype} != {want.dtype}"
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    for name in want.dtype.names:
        g, w = got[name], want[name]
        if w.dtype.kind in "iuUSb" or name in float_exact:
            if not np.array_equal(g, w, equal_nan=(w.dtype.kind == "f")):
                bad = np.flatnonzero(g != w)
                raise AssertionError(f"{what}.{name}: {bad.size} mismatches, first at {bad[:5]}: got {g[bad[:5]]} want {w[bad[:5]]}")
        else:
            ok = np.isclose(g, w, rtol=rtol, atol=atol, equal_nan=True)
            if not ok.all():
                bad = np.flatnonzero(~ok)
                raise AssertionError(f"{what}.{name}: {bad.size} out of tolerance, first at {bad[:5]}: got {g[bad[:5]]} want {w[bad[:5]]}")
