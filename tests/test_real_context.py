"""GPU: the B200 plugins inside a REAL reference Context (core/context_execution.py:140-183).

Two fresh interpreters run tests/real_context_run.py - once with the reference's CPU plugins, once with
``ctx.register(*b200_default(), allow_override=True)`` - on the same V1725 ``.bin`` raw files, and every
result pulled with ``ctx.get_data`` (records ... hit_grouped) is compared.  The reference comes from
``baseline/_ref`` on the GPU box (tools/install_reference.sh; git-ignored, travels with the snapshot)."""

import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ATOL, RTOL
from refctx import reference_root

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "tests", "real_context_run.py")


def _run(mode, tmp_path):
    out = os.path.join(str(tmp_path), f"{mode}.npz")
    res = subprocess.run([sys.executable, SCRIPT, mode, out, str(tmp_path)], capture_output=True, text=True, timeout=900)
    assert "REAL_CONTEXT_DONE" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
    return np.load(out, allow_pickle=False)


@pytest.mark.skipif(reference_root() is None, reason="reference package not installed (baseline/_ref)")
def test_records_route_through_a_real_context(tmp_path):
    cpu = _run("cpu", tmp_path)
    gpu = _run("b200", tmp_path)
    assert sorted(cpu.files) == sorted(gpu.files)
    for name in cpu.files:
        want, got = cpu[name], gpu[name]
        assert want.shape == got.shape and want.dtype == got.dtype, (name, want.shape, got.shape, want.dtype, got.dtype)
        if want.dtype.names:
            for f in want.dtype.names:
                w, g = want[f], got[f]
                if w.dtype.kind == "f":
                    assert np.allclose(g, w, rtol=RTOL, atol=ATOL, equal_nan=True), f"{name}.{f}"
                else:
                    assert np.array_equal(g, w), f"{name}.{f}"
        elif want.dtype.kind == "f":
            assert np.allclose(got, want, rtol=RTOL, atol=ATOL, equal_nan=True), name
        else:
            assert np.array_equal(got, want), name
    assert len(cpu["hit_threshold"]) > 1000 and len(cpu["hit_merged"]) < len(cpu["hit_threshold"])
