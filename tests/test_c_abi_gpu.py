"""The drop-in boundary used the way an MXNet maintainer would: a plain C program (tests/c_abi/
roialign_main.c) links libsimpledet_b200.so, owns its buffers through the CUDA runtime and calls
sdet_roi_align_v2_forward.  Its output must be bit-identical to the oracle."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle
from oracle import np_ops
from simpledet_b200 import build

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_plain_c_caller(tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    gcc = shutil.which("gcc")
    assert gcc, "gcc not found"
    lib = build.build()
    exe = str(tmp_path / "roialign_main")
    subprocess.run([gcc, "-O1", "-o", exe, os.path.join(HERE, "c_abi", "roialign_main.c"), "-I", os.path.join(ROOT, "include"),
                    "-I", os.path.join(cuda_home, "include"), lib, "-L", os.path.join(cuda_home, "lib64"), "-lcudart",
                    "-Wl,-rpath," + os.path.dirname(lib), "-Wl,-rpath," + os.path.join(cuda_home, "lib64")], check=True)
    rng = np.random.default_rng(3)
    B, N, C, H, W, PH, PW, scale = 2, 37, 24, 40, 52, 7, 7, 1 / 16
    data = rng.standard_normal((B, C, H, W)).astype(np.float32)
    xy = rng.uniform(0, [W * 16 * 0.8, H * 16 * 0.8], (B, N, 2))
    rois = np.concatenate([xy, xy + rng.uniform(8, 300, (B, N, 2))], 2).astype(np.float32)
    with open(tmp_path / "in.bin", "wb") as f:
        f.write(np.array([B, N, C, H, W, PH, PW], np.int32).tobytes())
        f.write(np.float32(scale).tobytes())
        f.write(data.tobytes())
        f.write(rois.tobytes())
    r = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    assert "launches=" in r.stdout and int(r.stdout.split("launches=")[1]) >= 1
    got = np.fromfile(tmp_path / "out.bin", np.float32).reshape(3, B, N, C, PH, PW)
    out, ax, ay = oracle.roi_align_v2_forward(data, rois, (PH, PW), scale)
    assert np.array_equal(got[0], out.reshape(B, N, C, PH, PW))
    assert np.array_equal(got[1], ax.reshape(B, N, C, PH, PW)) and np.array_equal(got[2], ay.reshape(B, N, C, PH, PW))


def test_reference_nms_symbol(tmp_path):
    """`_nms` (operator_py/cython/gpu_nms.hpp:1-2): a C program that only knows the reference's prototype gets the
    keep list of greedy NMS with `IoU > thresh` (nms_kernel.cu:71) - compared with the oracle's greedy NMS."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    gcc = shutil.which("gcc")
    lib = build.build()
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    exe = str(tmp_path / "nms_main")
    subprocess.run([gcc, "-O1", "-o", exe, os.path.join(HERE, "c_abi", "nms_main.c"), lib,
                    "-Wl,-rpath," + os.path.dirname(lib), "-Wl,-rpath," + os.path.join(cuda_home, "lib64")], check=True)
    rng = np.random.default_rng(5)
    for n, dim, thresh in ((1500, 5, 0.5), (300, 6, 0.3), (1, 5, 0.7)):
        xy = rng.uniform(0, 600, (n, 2))
        boxes = np.concatenate([xy, xy + rng.uniform(10, 200, (n, 2)), np.sort(rng.uniform(0, 1, (n, 1)), 0)[::-1]], 1)
        boxes = np.concatenate([boxes, rng.uniform(0, 1, (n, dim - 5))], 1).astype(np.float32)
        with open(tmp_path / "in.bin", "wb") as f:
            f.write(np.array([n, dim], np.int32).tobytes())
            f.write(np.float32(thresh).tobytes())
            f.write(boxes.tobytes())
        r = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr + r.stdout
        got = np.fromfile(tmp_path / "out.bin", np.int32)
        # oracle: operator_py/nms.py `nms` keeps ovr <= thresh, i.e. suppresses IoU > thresh like nms_kernel.cu:71;
        # on score-sorted input its result rows are the kept boxes in order
        kept_rows = np_ops.py_nms(np.ascontiguousarray(boxes[:, :5]), thresh)
        keep = [int(np.where((boxes[:, :5] == r).all(1))[0][0]) for r in kept_rows]
        assert got[0] == len(keep) and np.array_equal(got[1:], np.asarray(keep, np.int32))
