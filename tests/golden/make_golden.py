#!/usr/bin/env python
"""Generates tests/golden/*.npz by RUNNING THE REFERENCE'S OWN CODE in this container:
  - operator_py/cython/{bbox,cpu_nms}.pyx compiled by oracle/build_ref.py (oracle/_ref)
  - operator_py/bbox_transform.py and operator_py/nms.py, exec'd from /root/reference with the
    compiled Cython injected for their relative imports and `np.float = float` (they use the alias
    NumPy removed, bbox_transform.py:20,92,140).
Nothing from the reference is copied into the repo; only the input/output vectors are committed.
Run:  python tests/golden/make_golden.py     (needs /root/reference)"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

REF = "/root/reference/operator_py"


def load_reference():
    cy = oracle.ref_cython()
    assert cy is not None, "run oracle/build_ref.py first"
    np.float = float
    pkg = types.ModuleType("operator_py")
    pkg.__path__ = [REF]
    sys.modules["operator_py"] = pkg
    cpkg = types.ModuleType("operator_py.cython")
    cpkg.__path__ = []
    sys.modules["operator_py.cython"] = cpkg
    sys.modules["operator_py.cython.bbox"] = cy["bbox"]
    sys.modules["operator_py.cython.cpu_nms"] = cy["cpu_nms"]
    mods = {}
    for name in ("bbox_transform", "nms"):
        m = types.ModuleType(f"operator_py.{name}")
        m.__package__ = "operator_py"
        src = open(os.path.join(REF, name + ".py")).read()
        exec(compile(src, os.path.join(REF, name + ".py"), "exec"), m.__dict__)
        mods[name] = m
    return cy, mods


def boxes(rng, n, size=600.0):
    xy = rng.uniform(0, size, (n, 2))
    wh = rng.uniform(2, 250, (n, 2))
    return np.concatenate([xy, xy + wh], 1).astype(np.float32)


def main():
    cy, m = load_reference()
    bt, nm = m["bbox_transform"], m["nms"]
    rng = np.random.default_rng(20260922)
    out = {}
    # --- Cython IoU (float32 with the generated-C double promotions)
    b, q = boxes(rng, 257), boxes(rng, 19)
    out["overlaps_boxes"], out["overlaps_query"] = b, q
    out["overlaps_out"] = cy["bbox"].bbox_overlaps_cython(b, q)
    # --- greedy / soft NMS (distinct scores: the tie order of argsort()[::-1] is unspecified)
    d = np.concatenate([boxes(rng, 400, 300.0), rng.permutation(400)[:, None].astype(np.float32) / 400 + 1e-3], 1)
    d = d.astype(np.float32)
    out["nms_dets"] = d
    out["greedy_keep_0.5"] = cy["cpu_nms"].greedy_nms(d, np.float32(0.5))
    for method in (0, 1, 2):
        bx, idx = cy["cpu_nms"].soft_nms(d, np.float32(0.5), np.float32(0.3), np.float32(0.05), np.uint8(method))
        out[f"soft_boxes_{method}"], out[f"soft_inds_{method}"] = bx, idx
    out["py_nms_0.5"] = nm.nms(d, 0.5)
    out["soft_wrapper_linear"] = nm.cython_soft_nms_wrapper(0.5)(d)
    # --- numpy box transforms (float64)
    ex, gt = boxes(rng, 64).astype(np.float64), boxes(rng, 64).astype(np.float64)
    out["xf_ex"], out["xf_gt"] = ex, gt
    out["nonlinear_transform"] = bt.nonlinear_transform(ex, gt)
    deltas = rng.standard_normal((64, 4 * 5)) * 0.7
    deltas[0, 2::4] = 6.0  # exceeds BBOX_XFORM_CLIP
    out["xf_deltas"] = deltas
    out["nonlinear_pred"] = bt.nonlinear_pred(ex.astype(np.float32), deltas)
    out["iou_pred"] = bt.iou_pred(ex.astype(np.float32), deltas)
    out["clip_boxes"] = bt.clip_boxes(out["nonlinear_pred"].copy(), (400, 500))
    out["flip_boxes"] = bt.flip_boxes(ex, 640)
    np.savez_compressed(os.path.join(HERE, "reference_python_ops.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_python_ops.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
