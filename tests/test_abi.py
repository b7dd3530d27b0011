"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and
exports every symbol include/simpledet_b200.h declares (no compute calls: there is no GPU here)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g

    g.build()
    from simpledet_b200.build import LIB_PATH

    return LIB_PATH


def _declared():
    src = open(os.path.join(ROOT, "include", "simpledet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    # every sdet_* entry point plus `_nms`, the reference's own C symbol kept for compatibility (gpu_nms.hpp:1-2)
    return sorted(set(re.findall(r"\b(sdet_[a-z0-9_]+|_nms)\s*\(", src)))


def test_header_symbols_exported(built):
    out = subprocess.run(["nm", "-D", "--defined-only", built], capture_output=True, text=True,
                         check=True).stdout
    exported = set(re.findall(r" T (sdet_[a-z0-9_]+|_nms)\b", out))
    declared = _declared()
    assert declared, "header parse found no functions"
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_ctypes_binding_covers_header(built):
    from simpledet_b200 import _lib

    L = _lib.lib()
    assert sorted(_lib.EXPORTED_SYMBOLS) == _declared()
    assert L.sdet_abi_version() == 4
    assert L.sdet_last_error() == b""
    assert L.sdet_launch_count() == 0


def test_sm100a_sass_present(built):
    out = subprocess.run(["cuobjdump", "-lelf", built], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_arg_validation_without_gpu(built):
    """Argument checks run before any CUDA call, so they are testable on CPU: the reference's
    CHECK failures become error codes + messages."""
    from simpledet_b200 import _lib

    L = _lib.lib()
    rc = L.sdet_roi_align_v2_forward(None, None, None, None, None, 1, 1, 1, 4, 4, 7, 7, 0.5, None, 0, None)
    assert rc == 1 and b"non-NULL" in L.sdet_last_error()
    rc = L.sdet_roi_align_v2_forward(8, 8, 8, None, None, 1, 1, 1, 4, 4, 0, 7, 0.5, None, 0, None)
    assert rc == 1 and b"pooled_size" in L.sdet_last_error()
    rc = L.sdet_roi_align_v2_forward(8, 8, 8, None, None, 1, 1, 1, 4, 4, 64, 7, 0.5, None, 0, None)
    assert rc == 2
    rc = L.sdet_roi_align_v2_forward(8, 8, 8, None, None, 1, 1, 1, 4, 4, 7, 7, 1.5, None, 0, None)
    assert rc == 1 and b"spatial_scale" in L.sdet_last_error()


def test_ops_refuse_cpu_tensors(built):
    import torch

    from simpledet_b200 import ops

    with pytest.raises(RuntimeError, match="CUDA-only"):
        ops.ROIAlign_v2(torch.zeros(1, 1, 4, 4), torch.zeros(1, 1, 4), (2, 2), 1.0)


def test_ctypes_arity_matches_header(built):
    """Every binding in simpledet_b200/_lib.py has exactly as many argtypes as the C declaration."""
    from simpledet_b200 import _lib

    src = open(os.path.join(ROOT, "include", "simpledet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = dict(re.findall(r"\b(sdet_[a-z0-9_]+|_nms)\s*\(([^;{]*?)\)\s*;", src, flags=re.S))
    assert set(decls) == set(_lib._SIGNATURES)
    for name, params in decls.items():
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(_lib._SIGNATURES[name]), f"{name}: header has {n} parameters, binding has {len(_lib._SIGNATURES[name])}"
