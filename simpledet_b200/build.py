"""In-tree build of the C-ABI library (nvcc, sm_100a only).  Used by __graft_entry__.build().

Every csrc/*.cu is compiled to an object in simpledet_b200/_obj/ (in parallel, re-used while the source, the
headers and the flags are unchanged) and linked into libsimpledet_b200.so.  A digest of all sources, headers
and flags is compiled into the library (sdet_build_digest()); `source_digest()` recomputes it from the tree so
smoke() / tests can prove the binary that was loaded was built from the checked-out sources."""
from __future__ import annotations

import concurrent.futures
import glob
import hashlib
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_PKG, "libsimpledet_b200.so")
OBJ_DIR = os.path.join(_PKG, "_obj")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # The coordinate / value paths must round exactly like the reference's CPU build:
    # never contract a*b+c into FMA behind our back (kernels that want FMA ask for it).
    "-fmad=false",
    "-Xcompiler", "-fPIC",
]
LINK_FLAGS = ["-shared", "-cudart", "static"]


def sources() -> list[str]:
    return sorted(glob.glob(os.path.join(_PKG, "csrc", "*.cu")))


def _headers() -> list[str]:
    return sorted(glob.glob(os.path.join(_PKG, "csrc", "*.cuh")) + glob.glob(os.path.join(_ROOT, "include", "*.h")))


def _extra() -> list[str]:
    return os.environ.get("SDET_NVCC_EXTRA", "").split()  # e.g. -DSDET_BAND_CONSUMERS=15 for an A/B build


def source_digest() -> str:
    """sha256 over every source, header and compile flag of the library (hex, 16 chars)."""
    h = hashlib.sha256()
    for f in sources() + _headers():
        h.update(os.path.relpath(f, _ROOT).encode())
        h.update(open(f, "rb").read())
    h.update(" ".join(NVCC_FLAGS + _extra()).encode())
    return h.hexdigest()[:16]


def _compile_one(nvcc: str, src: str, digest: str, verbose: bool) -> tuple[str, str]:
    hdr = hashlib.sha256()
    for f in [src] + _headers():
        hdr.update(open(f, "rb").read())
    hdr.update(" ".join(NVCC_FLAGS + _extra()).encode())
    if os.path.basename(src) == "lib.cu":
        hdr.update(digest.encode())  # lib.cu embeds the digest of the whole tree
    key = hdr.hexdigest()[:16]
    obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + "." + key + ".o")
    log = ""
    if not os.path.exists(obj):
        for old in glob.glob(os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".*.o")):
            os.remove(old)
        cmd = [nvcc, *NVCC_FLAGS, *_extra(), f'-DSDET_BUILD_DIGEST="{digest}"', "-I", os.path.join(_ROOT, "include"),
               "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        log = r.stderr
    return obj, log


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile simpledet_b200/csrc/*.cu -> simpledet_b200/libsimpledet_b200.so (no-op when up to date)."""
    digest = source_digest()
    stamp = os.path.join(OBJ_DIR, "linked.digest")
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if force and os.path.isdir(OBJ_DIR):
        shutil.rmtree(OBJ_DIR)
    os.makedirs(OBJ_DIR, exist_ok=True)
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        res = list(ex.map(lambda s: _compile_one(nvcc, s, digest, verbose), sources()))
    if verbose:
        print("".join(log for _, log in res))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", *LINK_FLAGS, "-o", LIB_PATH, *[o for o, _ in res]]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    open(stamp, "w").write(digest)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
