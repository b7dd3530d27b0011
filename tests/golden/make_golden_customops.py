#!/usr/bin/env python
"""Generates tests/golden/reference_customops.npz by EXEC'ING THE REFERENCE'S OWN Python CustomOps
(models/FPN/assign_layer_fpn.py, models/FPN/get_top_proposal.py, models/maskrcnn/bbox_post_processing.py) under a
small stand-in for `mxnet` whose `mx.nd` functions are numpy float32 (MXNet computes these elementwise ops in the
array's dtype; scalars are cast to it) - the operators' forward() methods run unmodified on it.
mx.nd.argsort(is_ascend=False) is a STABLE descending sort here (ties: lower index first).
Run:  python tests/golden/make_golden_customops.py     (needs /root/reference)"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from simpledet_b200 import synth  # noqa: E402


class ND:
    """The slice of mx.nd.NDArray these three operators touch, over a numpy array."""

    def __init__(self, a):
        self.a = np.asarray(a)

    @property
    def shape(self):
        return self.a.shape

    def asnumpy(self):
        return np.array(self.a)

    def astype(self, t):
        return ND(self.a.astype(t))

    def _b(self, o):
        return o.a if isinstance(o, ND) else self.a.dtype.type(o)

    def __getitem__(self, k):
        if isinstance(k, ND):
            k = k.a.astype(np.int64)
        elif isinstance(k, tuple):
            k = tuple(x.a.astype(np.int64) if isinstance(x, ND) else x for x in k)
        return ND(self.a[k])

    def __setitem__(self, k, v):
        self.a[k] = v.a if isinstance(v, ND) else v

    def __add__(self, o): return ND(self.a + self._b(o))
    def __sub__(self, o): return ND(self.a - self._b(o))
    def __mul__(self, o): return ND(self.a * self._b(o))
    def __truediv__(self, o): return ND(self.a / self._b(o))
    def __rpow__(self, o): return ND(np.power(self.a.dtype.type(o), self.a))
    def __eq__(self, o): return ND((self.a == self._b(o)).astype(self.a.dtype))
    __radd__, __rmul__ = __add__, __mul__


def make_mx():
    mx = types.ModuleType("mxnet")
    nd = types.SimpleNamespace(
        sqrt=lambda x: ND(np.sqrt(x.a)), floor=lambda x: ND(np.floor(x.a)), log2=lambda x: ND(np.log2(x.a)),
        clip=lambda x, lo, hi: ND(np.clip(x.a, x.a.dtype.type(lo), x.a.dtype.type(hi))),
        zeros_like=lambda x: ND(np.zeros_like(x.a)), expand_dims=lambda x, axis: ND(np.expand_dims(x.a, axis)),
        broadcast_like=lambda x, y: ND(np.broadcast_to(x.a, y.a.shape).copy()),
        where=lambda c, x, y: ND(np.where(c.a != 0, x.a, y.a)),
        argsort=lambda x, is_ascend=True: ND(np.argsort(x.a if is_ascend else -x.a, kind="stable").astype(np.float32)),
        stack=lambda *xs: ND(np.stack([x.a for x in xs])))

    class CustomOp:
        def assign(self, dst, req, src):
            dst[:] = src

    class CustomOpProp:
        def __init__(self, need_top_grad=False):
            pass

    mx.nd = nd
    mx.operator = types.SimpleNamespace(CustomOp=CustomOp, CustomOpProp=CustomOpProp, register=lambda name: (lambda c: c))
    return mx


def load(path, extra=None):
    m = types.ModuleType("ref_" + os.path.basename(path)[:-3])
    m.__dict__.update(extra or {})
    exec(compile(open(path).read(), path, "exec"), m.__dict__)
    return m


def main():
    sys.modules["mxnet"] = make_mx()
    # bbox_post_processing.py imports py_nms_wrapper from operator_py.nms: the real one (make_golden.py's loader)
    sys.path.insert(0, HERE)
    import make_golden as mg

    _, mods = mg.load_reference()
    ref = "/root/reference"
    g = {}
    rng = np.random.default_rng(0)
    # assign_layer_fpn
    al = load(f"{ref}/models/FPN/assign_layer_fpn.py")
    rois = synth.random_rois(rng, 2, 400, min_side=4, max_side=900)
    rois[0, :4] = [[0, 0, 0, 0], [10, 10, 233, 233], [10, 10, 232.99, 233], [0, 0, 1332, 799]]
    op = al.AssignLayerFPNOperator((4, 8, 16, 32), 224, 4)
    outs = [ND(np.zeros_like(rois)) for _ in range(4)]
    op.forward(False, ["write"] * 4, [ND(rois)], outs, [])
    g.update(al_rois=rois, **{f"al_out{i}": o.a for i, o in enumerate(outs)})
    # get_top_proposal
    gt = load(f"{ref}/models/FPN/get_top_proposal.py")
    boxes = synth.random_rois(rng, 2, 500)
    scores = rng.uniform(0, 1, (2, 500, 1)).astype(np.float32)
    scores[0, 100:110] = scores[0, 99]   # ties
    ob, os_ = ND(np.zeros((2, 200, 4), np.float32)), ND(np.zeros((2, 200, 1), np.float32))
    gt.GetTopProposalOperator(200).forward(False, ["write"] * 2, [ND(boxes), ND(scores)], [ob, os_], [])
    g.update(gt_boxes=boxes, gt_scores=scores, gt_out_boxes=ob.a, gt_out_scores=os_.a)
    # BboxPostProcessing
    sys.modules["operator_py.nms"] = mods["nms"]
    bp = load(f"{ref}/models/maskrcnn/bbox_post_processing.py")
    B, N, K = 2, 300, 6
    cls_score = rng.uniform(0, 1, (B, N, K)).astype(np.float32) ** 3
    bb = synth.random_rois(rng, B, N)
    bbox = np.concatenate([bb + rng.uniform(-3, 3, bb.shape).astype(np.float32) for _ in range(K)], 2)
    outs = [ND(np.zeros((B, 50, 1), np.float32)), ND(np.zeros((B, 50, 4), np.float32)), ND(np.zeros((B, 50, 1), np.float32))]
    bp.BboxPostProcessingOperator(50, 0.3, "nms", 0.5).forward(False, ["write"] * 3, [ND(cls_score), ND(bbox)], outs, [])
    g.update(bp_cls_score=cls_score, bp_bbox=bbox, bp_score=outs[0].a, bp_box=outs[1].a, bp_cls=outs[2].a)
    path = os.path.join(HERE, "reference_customops.npz")
    np.savez_compressed(path, **g)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
