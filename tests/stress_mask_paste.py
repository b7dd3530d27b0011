#!/usr/bin/env python
"""Stress check of the mask-paste kernels' source against the reference itself, at test size (needs /root/reference
and cv2; CPU only): 120 random detections on an 800 x 1333 image through (a) the reference's own segm_results
(models/maskrcnn/utils.py:26-67) on the installed cv2 with a stand-in pycocotools encoder and (b) the host emulation of
mask_paste.cu's kernels (tests/c_abi/mask_paste_emul.cc) under ops._segm_results_impl; compares the RLE strings.
    python tests/stress_mask_paste.py SEED   (test infrastructure: it imports the oracle)"""
import ctypes
import os
import subprocess
import sys
import tempfile
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import oracle
from oracle import np_ops
from simpledet_b200 import ops
pm = types.ModuleType("pycocotools"); mk = types.ModuleType("pycocotools.mask")
def encode(arr):
    h, w, n = arr.shape
    return [{"size": [h, w], "counts": np_ops.rle_to_string(oracle.rle_encode(np.ascontiguousarray(arr[:, :, i])))} for i in range(n)]
mk.encode = encode; pm.mask = mk
sys.modules["pycocotools"], sys.modules["pycocotools.mask"] = pm, mk
sys.path.insert(0, "/root/reference")
from models.maskrcnn.utils import segm_results
so = os.path.join(tempfile.mkdtemp(), "libemul.so")
subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-std=c++17", "-I", os.path.join(ROOT, "simpledet_b200", "csrc"),
                "-o", so, os.path.join(ROOT, "tests", "c_abi", "mask_paste_emul.cc")], check=True)
emul = ctypes.CDLL(so)
from test_mask_paste_host import run_emulated
rng = np.random.default_rng(int(sys.argv[1]))
im_h, im_w, n, k, m = 800, 1333, 120, 3, 28
xy = rng.uniform(-30, [im_w - 40, im_h - 40], (n, 2)); wh = rng.uniform(35, [900, 700], (n, 2))
box = np.concatenate([xy, xy + wh], 1).astype(np.float32)
z = rng.standard_normal((n, k, 7, 7)).astype(np.float32)
masks = np.ascontiguousarray(1/(1+np.exp(-1.5*np.kron(z, np.ones((4,4),np.float32)))) + rng.normal(0,0.02,(n,k,28,28)), np.float32)
cls = rng.integers(0, k, n).astype(np.int32)
t=time.time(); want = segm_results(box, cls, masks, im_h, im_w); t1=time.time()-t
t=time.time(); got = run_emulated(emul, box, cls, masks, im_h, im_w); t2=time.time()-t
bad = sum(g["counts"] != w["counts"] for g, w in zip(got, want))
print("seed", sys.argv[1], "mismatching detections:", bad, "of", n, "| reference %.1fs, emulated kernels + host logic %.1fs" % (t1, t2))
