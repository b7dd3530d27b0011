#!/usr/bin/env python
"""Generates tests/golden/*.npz by RUNNING THE REFERENCE'S OWN CODE in this container:
  - operator_py/cython/{bbox,cpu_nms}.pyx compiled by oracle/build_ref.py (oracle/_ref)
  - operator_py/bbox_transform.py and operator_py/nms.py, exec'd from /root/reference with the
    compiled Cython injected for their relative imports and `np.float = float` (they use the alias
    NumPy removed, bbox_transform.py:20,92,140).
Nothing from the reference is copied into the repo; only the input/output vectors are committed.
  - core/detection_input.py AnchorTarget2D and models/FPN/input.py PyramidAnchorTarget2D, exec'd the same
    way with DEBUG=True (deterministic sub-sampling) -> reference_anchor_target.npz
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


def load_anchor_targets():
    """core/detection_input.py (AnchorTarget2D) and models/FPN/input.py (PyramidAnchorTarget2D), exec'd
    with empty stand-ins for cv2 / mxnet (imported at module scope there, unused by these classes)."""
    for name in ("cv2", "mxnet"):
        sys.modules.setdefault(name, types.ModuleType(name))
    mx = sys.modules["mxnet"]
    if not hasattr(mx, "io"):  # `class Loader(mx.io.DataIter)` is defined at module scope
        mx.io = types.SimpleNamespace(DataIter=object)
    core = types.ModuleType("core")
    core.__path__ = []
    sys.modules["core"] = core
    di = types.ModuleType("core.detection_input")
    path = "/root/reference/core/detection_input.py"
    exec(compile(open(path).read(), path, "exec"), di.__dict__)
    sys.modules["core.detection_input"] = di
    fi = types.ModuleType("models.FPN.input")
    path = "/root/reference/models/FPN/input.py"
    exec(compile(open(path).read(), path, "exec"), fi.__dict__)
    return di.AnchorTarget2D, fi.PyramidAnchorTarget2D


class _NS:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def anchor_param(stride, short, long, scales, aspects, border, pos, neg, min_pos, num, frac):
    return _NS(generate=_NS(stride=stride, short=short, long=long, scales=scales, aspects=aspects),
               assign=_NS(allowed_border=border, pos_thr=pos, neg_thr=neg, min_pos_thr=min_pos),
               sample=_NS(image_anchor=num, pos_fraction=frac))


def gt_boxes(rng, n, max_gt, h, w):
    g = np.full((max_gt, 5), -1, np.float32)
    xy = rng.uniform(0, [w * 0.7, h * 0.7], (n, 2))
    wh = rng.uniform(12, [w * 0.5, h * 0.5], (n, 2))
    g[:n, :4] = np.concatenate([xy, np.minimum(xy + wh, [w - 1, h - 1])], 1)
    g[:n, 4] = rng.integers(1, 81, n)
    return g


def anchor_goldens():
    """Reference AnchorTarget2D / PyramidAnchorTarget2D with DEBUG=True (deterministic sub-sampling:
    the FIRST surplus indices are disabled, detection_input.py:487-494)."""
    A2D, P2D = load_anchor_targets()
    rng = np.random.default_rng(777)
    out = {}
    # single level (C4 style), both orientations, allowed_border 0
    cfg = dict(stride=16, short=12, long=18, scales=(2, 4, 8), aspects=(0.5, 1.0, 2.0), border=0, pos=0.7, neg=0.3,
               min_pos=0.0, num=64, frac=0.5)
    for tag, (h, w) in (("h", (180.0, 280.0)), ("v", (288.0, 170.0)), ("fgcap", (190.0, 270.0))):
        if tag == "fgcap":  # more positives than pos_fraction * image_anchor: the fg branch of _sample_anchor
            cfg = dict(cfg, num=8, pos=0.5)
        op = A2D(anchor_param(**cfg))
        op.DEBUG = True
        gt = gt_boxes(rng, 5, 8, h, w)
        rec = {"im_info": np.array([h, w, 1.0], np.float32), "gt_bbox": gt}
        lab, tgt, wgt = op.apply(rec)
        out[f"a2d_{tag}_im_info"], out[f"a2d_{tag}_gt"] = rec["im_info"], gt
        out[f"a2d_{tag}_label"], out[f"a2d_{tag}_target"], out[f"a2d_{tag}_weight"] = lab, tgt, wgt
    # pyramid (FPN style), allowed_border 9999, plus an image without gt
    pcfg = dict(stride=(4, 8, 16, 32), short=(40, 20, 10, 5), long=(60, 30, 15, 8), scales=(8,),
                aspects=(0.5, 1.0, 2.0), border=9999, pos=0.7, neg=0.3, min_pos=0.0, num=256, frac=0.5)
    for tag, (h, w), ngt in (("h", (150.0, 236.0), 7), ("v", (240.0, 158.0), 3), ("empty", (160.0, 240.0), 0)):
        op = P2D(anchor_param(**pcfg))
        op.anchor_target_2d.DEBUG = True
        gt = gt_boxes(rng, ngt, 10, h, w)
        rec = {"im_info": np.array([h, w, 1.5], np.float32), "gt_bbox": gt}
        lab, tgt, wgt = op.apply(rec)
        out[f"p2d_{tag}_im_info"], out[f"p2d_{tag}_gt"] = rec["im_info"], gt
        out[f"p2d_{tag}_label"], out[f"p2d_{tag}_target"], out[f"p2d_{tag}_weight"] = lab, tgt, wgt
    np.savez_compressed(os.path.join(HERE, "reference_anchor_target.npz"), **out)
    print("wrote reference_anchor_target.npz", {k: (v.shape, v.dtype) for k, v in out.items() if "label" in k or "target" in k})


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
    out["selfoverlaps_out"] = cy["bbox_self"].bbox_selfoverlaps_cython(b, q)
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
    d6 = np.concatenate([d, rng.integers(0, 7, (400, 1)).astype(np.float32)], 1)
    out["set_nms_dets"] = d6
    out["set_nms_0.4"] = nm.set_nms(d6, 0.4)
    out["weighted_nms_0.3_0.6"] = nm.py_weighted_nms(d, 0.3, 0.6)
    out["weighted_nms_0.5_0.5"] = nm.py_weighted_nms(d, 0.5, 0.5)
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
    # --- box voting (float32 dets; every scoring method)
    bt.bbox_overlaps = cy["bbox"].bbox_overlaps_cython
    top = d[np.argsort(-d[:, 4])[:60]].copy()
    out["vote_top"] = top
    for meth, beta in (("ID", 1.0), ("AVG", 1.0), ("IOU_AVG", 1.0), ("GENERALIZED_AVG", 2.0), ("QUASI_SUM", 0.5),
                       ("TEMP_AVG", 0.7)):
        out[f"vote_{meth}"] = bt.box_voting(top, d, 0.5, meth, beta)
    np.savez_compressed(os.path.join(HERE, "reference_python_ops.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_python_ops.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
    anchor_goldens()
