#!/usr/bin/env python
"""Generates tests/golden/reference_maskiou_compute.npz by running the reference's OWN CustomOp
(models/msrcnn/maskiou_compute.py: MaskIoUComputeOperator.forward, unmodified; Mask Scoring R-CNN's MaskIoU target, the
consumer of ProposalMaskTarget(output_ratio=True)) on the numpy stand-in for mx.nd of make_golden_customops.py.
Run:  python tests/golden/make_golden_maskiou.py   (needs /root/reference)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_customops import ND, make_mx  # noqa: E402


def main():
    sys.modules["mxnet"] = make_mx()
    for k, v in (("bool", bool), ("float", float)):       # the reference predates NumPy 1.24
        if not hasattr(np, k):
            setattr(np, k, v)
    sys.path.insert(0, "/root/reference")
    from models.msrcnn.maskiou_compute import MaskIoUComputeOperator

    rng = np.random.default_rng(21)
    R, M = 96, 28
    logits = rng.normal(0.3, 1.0, (R, M, M)).astype(np.float32)
    target = (rng.random((R, M, M)) < 0.4).astype(np.float32)
    target[5] = 0                                           # empty target: union clamps to >= 1
    target[6] = -1                                          # a background slot's ignore mask
    logits[7] = -3                                          # empty prediction
    ratio = rng.uniform(0.2, 1.0, (R,)).astype(np.float32)
    ratio[8] = 1.0
    inds = rng.integers(0, 5, (R,)).astype(np.float32)      # class per slot; 0 = background: weight 0
    out = [ND(np.zeros((R, 1), np.float32)), ND(np.zeros((R, 1), np.float32))]
    MaskIoUComputeOperator().forward(True, ["write"] * 2, [ND(logits), ND(target), ND(ratio), ND(inds)], out, [])
    np.savez_compressed(os.path.join(HERE, "reference_maskiou_compute.npz"), logits=logits, target=target, ratio=ratio,
                        inds=inds, iou=out[0].a, weight=out[1].a)
    print("iou range", out[0].a.min(), out[0].a.max(), "positive slots", int(out[1].a.sum()))


if __name__ == "__main__":
    main()
