"""INTEGRATION.md §1's MXNet-side stub (tests/c_abi/mxnet_roi_align_stub.cc) must compile against the reference's own
roi_align_v2-inl.h, the MXNet stand-in of oracle/shim and include/simpledet_b200.h - so the documented binding can
not drift from the C ABI (round 1's snippet passed 14 of 16 arguments).  CPU only; needs /root/reference."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/operator_cxx/contrib"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference headers not present")
def test_stub_compiles_against_header_and_reference(tmp_path):
    gxx = shutil.which("g++")
    assert gxx
    shim = os.path.join(ROOT, "oracle", "shim")
    cmd = [gxx, "-std=c++14", "-fsyntax-only", "-w", "-I", REF, "-I", shim, "-I", os.path.join(shim, "l1"),
           "-I", os.path.join(shim, "l1", "l2"), "-I", os.path.join(ROOT, "include"), "-include",
           os.path.join(shim, "mxnet_shim.h"), os.path.join(ROOT, "tests", "c_abi", "mxnet_roi_align_stub.cc")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_integration_md_quotes_the_compiled_stub():
    """The C++ block of INTEGRATION.md §1 is the stub file, line for line."""
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    stub = open(os.path.join(ROOT, "tests", "c_abi", "mxnet_roi_align_stub.cc")).read()
    body = stub[stub.index("#define SDET_CALL"):]
    assert body.strip() in md
