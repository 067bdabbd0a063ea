"""Locate and import the UNMODIFIED reference package for tests that compare against it.

Search order: ``$WFB_REFERENCE_ROOT``, ``/root/reference`` (the read-only checkout of the build
container), ``baseline/_ref`` (``pip install --target`` copy made by ``tools/install_reference.sh``;
git-ignored, but it travels to the GPU box).  matplotlib is absent from the image and imported
unconditionally by the reference (core/context.py:34), so it is stubbed before the import.
"""

from __future__ import annotations

import os
import sys
from unittest.mock import MagicMock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_MPL = ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.colors", "matplotlib.figure", "matplotlib.axes",
        "matplotlib.gridspec", "matplotlib.lines", "matplotlib.collections", "matplotlib.cm", "matplotlib.ticker", "matplotlib.dates")


def reference_root() -> str | None:
    for cand in (os.environ.get("WFB_REFERENCE_ROOT"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "waveform_analysis")):
            return cand
    return None


def import_reference():
    """Put the reference on sys.path (matplotlib stubbed) and import it; raises ImportError if absent."""
    root = reference_root()
    if root is None:
        raise ImportError("reference package not found (WFB_REFERENCE_ROOT, /root/reference, baseline/_ref)")
    for mod in _MPL:
        sys.modules.setdefault(mod, MagicMock())
    if root not in sys.path:
        sys.path.insert(0, root)
    import waveform_analysis  # noqa: F401

    return waveform_analysis
