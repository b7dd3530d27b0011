#!/usr/bin/env python
"""Generates tests/golden/reference_decode_retina.npz by running the reference's OWN CustomOp
(models/retinanet/decode_retina.py: DecodeRetinaOperator.forward, unmodified) on the numpy stand-in for mx.nd of
make_golden_customops.py.  Run:  python tests/golden/make_golden_decode_retina.py   (needs /root/reference)"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_customops import ND, make_mx  # noqa: E402

REF = "/root/reference"


def main():
    mx = make_mx()
    mx.io = types.SimpleNamespace(DataIter=object, DataBatch=object, DataDesc=object)  # core/detection_input.py:579 subclasses it
    sys.modules["mxnet"] = mx
    if not hasattr(np, "float"):
        np.float = float  # the reference predates NumPy 1.24
    cy = types.ModuleType("operator_py.cython.bbox")
    cy.bbox_overlaps_cython = lambda a, b: (_ for _ in ()).throw(RuntimeError("not used"))
    sys.path.insert(0, REF)
    import operator_py  # noqa: F401
    pkg = types.ModuleType("operator_py.cython")
    pkg.__path__ = []
    sys.modules["operator_py.cython"] = pkg
    sys.modules["operator_py.cython.bbox"] = cy
    from models.retinanet.decode_retina import DecodeRetinaOperator

    rng = np.random.default_rng(7)
    stride, scales, ratios = (8, 16, 32, 64, 128), (4 * 2 ** 0, 4 * 2 ** (1 / 3), 4 * 2 ** (2 / 3)), (0.5, 1, 2)
    A, K, top, thresh = 9, 5, 40, 0.05
    H0, W0 = 24, 40
    shapes = [(max(1, -(-H0 * 8 // s)), max(1, -(-W0 * 8 // s))) for s in stride]
    cls, reg = [], []
    for h, w in shapes:
        z = rng.standard_normal((1, A * K, h, w)).astype(np.float32) * 1.5 - 3.0
        cls.append((1 / (1 + np.exp(-z))).astype(np.float32))
        r = (rng.standard_normal((1, A * 4, h, w)) * 0.5).astype(np.float32)
        r[0, 2, 0, 0] = 9.0  # exercises the BBOX_XFORM_CLIP clamp
        reg.append(r)
    cls[0][0, 3, 2, 5] = cls[0][0, 7, 1, 1] = 0.875  # a score tie
    cls[1][:] = 0.01                                  # a level with no candidate above the threshold
    im_info = np.array([[H0 * 8 - 3, W0 * 8 - 5, 1.0]], np.float32)
    op = DecodeRetinaOperator(stride, scales, ratios, top, thresh)
    out = [ND(np.zeros((1, top * len(stride), 4), np.float32)), ND(np.zeros((1, top * len(stride), K + 1), np.float32))]
    op.forward(False, ["write", "write"], [ND(c) for c in cls] + [ND(r) for r in reg] + [ND(im_info)], out, [])
    d = {"stride": np.array(stride), "scales": np.array(scales), "ratios": np.array(ratios), "top": top, "thresh": thresh,
         "im_info": im_info, "boxes": out[0].a, "scores": out[1].a}
    for i, (c, r) in enumerate(zip(cls, reg)):
        d[f"cls{i}"], d[f"reg{i}"] = c, r
    np.savez_compressed(os.path.join(HERE, "reference_decode_retina.npz"), **d)
    print("rows with a score:", int((out[1].a.sum(-1) > 0).sum()), "of", top * len(stride))


if __name__ == "__main__":
    main()
