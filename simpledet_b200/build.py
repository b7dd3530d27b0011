"""In-tree build of the C-ABI library (nvcc, sm_100a only).  Used by __graft_entry__.build()."""
from __future__ import annotations

import glob
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_PKG, "libsimpledet_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # The coordinate / value paths must round exactly like the reference's CPU build:
    # never contract a*b+c into FMA behind our back (kernels that want FMA ask for it).
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
]


def sources() -> list[str]:
    return sorted(glob.glob(os.path.join(_PKG, "csrc", "*.cu")))


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(_PKG, "csrc", "*.cuh")) + glob.glob(
        os.path.join(_ROOT, "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile simpledet_b200/csrc/*.cu -> simpledet_b200/libsimpledet_b200.so."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    extra = os.environ.get("SDET_NVCC_EXTRA", "").split()  # e.g. -DSDET_RA_ABLATE for a profiling build
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-I", os.path.join(_ROOT, "include"), "-o", LIB_PATH, *sources()]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
