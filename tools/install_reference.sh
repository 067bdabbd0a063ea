#!/bin/bash
# Install the UNMODIFIED reference package into baseline/_ref (git-ignored, travels to the GPU box with the snapshot).
# bench.py --impl reference, bench.py's cpu_baseline leg and tests/test_real_context.py import it from there.
#
#   tools/install_reference.sh [/path/to/reference/checkout]      (default: /root/reference)
#
# The checkout is read-only and setuptools writes egg-info next to the sources, so the install runs from a copy under
# /tmp; --no-deps because the image already has numpy / scipy / pandas / numba-free fallbacks (resolution against the
# offline wheelhouse fails on the reference's numpy pin).  matplotlib is not in the image: tests/refctx.py stubs it.
set -euo pipefail
src="${1:-/root/reference}"
root="$(cd "$(dirname "$0")/.." && pwd)"
tmp="$(mktemp -d /tmp/refcopy.XXXXXX)"
cp -r "$src"/. "$tmp"/
rm -rf "$root/baseline/_ref"
mkdir -p "$root/baseline"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$root/baseline/_ref" "$tmp"
rm -rf "$tmp"
python - <<PY
import sys
sys.path.insert(0, "$root/tests")
from refctx import import_reference, reference_root
import os
os.environ["WFB_REFERENCE_ROOT"] = "$root/baseline/_ref"
print("reference importable from", reference_root(), "->", import_reference().__name__)
PY
