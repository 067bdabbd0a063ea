"""In-tree build of libwfb200.so with nvcc for sm_100a (cross-compiles without a GPU).

    python -m waveformanalysis_b200.build [--force] [--verbose]

The shared library is written next to this file; it is git-ignored but travels to the GPU box.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libwfb200.so")
OBJ_DIR = os.path.join(HERE, "csrc", "_obj")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (needed to build libwfb200.so)")


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime() -> float:
    paths = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    paths.append(os.path.join(os.path.dirname(HERE), "include", "wfb200.h"))
    return max(os.path.getmtime(p) for p in paths)


STAMP = os.path.join(OBJ_DIR, "build_stamp.txt")


def _stamp(nvcc: str) -> str:
    """What the objects were built with: compiler version + flags.  A change rebuilds everything."""
    try:
        ver = subprocess.run([nvcc, "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-1]
    except OSError:
        ver = "unknown"
    return ver + "\n" + " ".join(NVCC_FLAGS) + "\n"


def build(force: bool = False, verbose: bool = False) -> str:
    try:
        nvcc = _nvcc()
    except RuntimeError:
        if os.path.exists(OUT):  # a box without the toolkit (the prebuilt library travelled with the snapshot)
            return OUT
        raise
    stamp = _stamp(nvcc)
    same_toolchain = os.path.exists(STAMP) and open(STAMP).read() == stamp
    if not force and same_toolchain and os.path.exists(OUT) and os.path.getmtime(OUT) >= _deps_mtime():
        return OUT
    os.makedirs(OBJ_DIR, exist_ok=True)
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", src, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
        if verbose:
            sys.stderr.write(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources()))) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [nvcc, "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    with open(STAMP, "w") as f:
        f.write(stamp)
    return OUT


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
