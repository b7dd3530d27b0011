#!/usr/bin/env python
"""Generates tests/golden/reference_bbox_target.npz by running the reference's OWN CustomOp
(operator_py/bbox_target.py: BboxTargetOperator.forward, unmodified; the Detectron-style sampler of the crowdhuman
detector, models/crowdhuman/builder.py:380-396) on the numpy stand-in for mx.nd of make_golden_customops.py, with the
reference's own compiled Cython `bbox_overlaps_cython` (oracle/_ref, oracle/build_ref.py) behind
operator_py/detectron_bbox_utils.py.  The operator samples with the GLOBAL numpy RNG (`numpy.random.choice`):
each case records the seed it ran under.  Run:  python tests/golden/make_golden_bbox_target.py   (needs /root/reference)"""
import importlib.util
import glob
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
from make_golden_customops import ND, make_mx  # noqa: E402

REF = "/root/reference"


def _load_ref_cython(name):
    path = glob.glob(os.path.join(ROOT, "oracle", "_ref", name + ".*.so"))[0]
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def cases():
    """(name, proposals (B,K,4), gt (B,M,5), kwargs).  Every image has at least `image_rois` candidates (with fewer
    the reference returns ragged lists that np.array() cannot stack)."""
    rng = np.random.default_rng(11)
    out = []
    for name, B, K, M, nvalid, kw in (
            ("c81", 2, 300, 20, (7, 12), dict(num_class=81, add_gt_to_proposal=True, image_rois=64, fg_fraction=0.25,
                                              fg_thresh=0.5, bg_thresh_hi=0.5, bg_thresh_lo=0.0,
                                              bbox_target_std=(0.1, 0.1, 0.2, 0.2))),
            ("agnostic", 2, 700, 16, (5, 9), dict(num_class=2, add_gt_to_proposal=True, image_rois=128, fg_fraction=0.25,
                                                  fg_thresh=0.5, bg_thresh_hi=0.5, bg_thresh_lo=0.05,
                                                  bbox_target_std=(0.1, 0.1, 0.2, 0.2))),
            ("nogt_fgshort", 1, 200, 8, (3,), dict(num_class=4, add_gt_to_proposal=False, image_rois=32, fg_fraction=0.5,
                                                   fg_thresh=0.93, bg_thresh_hi=0.6, bg_thresh_lo=0.0,
                                                   bbox_target_std=(1.0, 1.0, 1.0, 1.0)))):
        gt = np.full((B, M, 5), -1, np.float32)
        prop = np.zeros((B, K, 4), np.float32)
        for b in range(B):
            n = nvalid[b]
            xy = rng.uniform(0, 500, (n, 2))
            wh = rng.uniform(40, 260, (n, 2))
            gt[b, :n, :4] = np.concatenate([xy, xy + wh], 1)
            gt[b, :n, 4] = rng.integers(1, max(2, kw["num_class"] if kw["num_class"] > 2 else 2), n)
            kp = K - 25                                       # the tail stays zero: y2 == 0 marks padding
            src = rng.integers(0, n, kp)
            jit = rng.normal(0, 1, (kp, 4)) * np.tile(wh[src], 2) * rng.choice([0.04, 0.15, 0.6], (kp, 1))
            p = gt[b, src, :4] + jit
            p = np.stack([np.minimum(p[:, 0], p[:, 2]), np.minimum(p[:, 1], p[:, 3]),
                          np.maximum(p[:, 0], p[:, 2]), np.maximum(p[:, 1], p[:, 3])], 1)
            prop[b, :kp] = np.clip(p, 0, 799).astype(np.float32)
            prop[b, :kp, 3] = np.maximum(prop[b, :kp, 3], 1.0)
            prop[b, 3] = prop[b, 2]                           # a duplicated proposal (IoU tie between rows)
            prop[b, 4] = gt[b, 0, :4]                         # IoU exactly 1 with gt 0
            if n > 1:
                gt[b, 1, :4] = gt[b, 0, :4]                   # two identical gt boxes: argmax takes the first
        out.append((name, prop, gt, kw))
    return out


def main():
    mx = make_mx()
    sys.modules["mxnet"] = mx
    if not hasattr(np, "float"):
        np.float = float  # the reference predates NumPy 1.24
    sys.path.insert(0, REF)
    import operator_py  # noqa: F401
    pkg = types.ModuleType("operator_py.cython")
    pkg.__path__ = []
    sys.modules["operator_py.cython"] = pkg
    sys.modules["operator_py.cython.bbox"] = _load_ref_cython("bbox")
    sys.modules["operator_py.cython.cpu_nms"] = _load_ref_cython("cpu_nms")
    from operator_py.bbox_target import BboxTargetOperator

    d = {}
    for i, (name, prop, gt, kw) in enumerate(cases()):
        seed = 100 + i
        op = BboxTargetOperator(kw["num_class"], kw["add_gt_to_proposal"], kw["image_rois"], kw["fg_fraction"],
                                kw["fg_thresh"], kw["bg_thresh_hi"], kw["bg_thresh_lo"], kw["bbox_target_std"])
        B, R, C = prop.shape[0], kw["image_rois"], kw["num_class"]
        out = [ND(np.zeros((B, R, 4), np.float32)), ND(np.zeros((B, R), np.float32)),
               ND(np.zeros((B, R, 4 * C), np.float32)), ND(np.zeros((B, R, 4 * C), np.float32))]
        np.random.seed(seed)
        op.forward(True, ["write"] * 4, [ND(prop), ND(gt)], out, [])
        d[f"{name}_prop"], d[f"{name}_gt"], d[f"{name}_seed"] = prop, gt, seed
        for k, v in kw.items():
            d[f"{name}_kw_{k}"] = np.asarray(v)
        for k, o in zip(("rois", "label", "target", "weight"), out):
            d[f"{name}_{k}"] = o.a
        lab = out[1].a
        print(name, "fg per image:", [(lab[b] > 0).sum() for b in range(B)], "of", R)
    d["names"] = np.array([c[0] for c in cases()])
    np.savez_compressed(os.path.join(HERE, "reference_bbox_target.npz"), **d)


if __name__ == "__main__":
    main()
