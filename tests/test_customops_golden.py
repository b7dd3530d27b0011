"""The oracle's restatements of the reference's Python CustomOps (FPN level assignment, get_top_proposal, the
BboxPostProcessing selection) against vectors produced by exec'ing the reference's own files under an mx.nd stand-in
(tests/golden/make_golden_customops.py).  CPU only."""
import os

import numpy as np

import oracle
from oracle import np_ops

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_customops.npz"))


def test_assign_layer_fpn():
    rois = G["al_rois"]
    idx = oracle.fpn_assign_levels(rois, (4, 8, 16, 32)).reshape(rois.shape[:2])
    for i in range(4):
        want = np.where((idx == i)[..., None], rois, np.float32(0))
        assert np.array_equal(want, G[f"al_out{i}"]), i
    # every roi lands on exactly one level
    assert ((idx >= 0) & (idx < 4)).all()


def test_get_top_proposal():
    ob, os_ = np_ops.get_top_proposal(G["gt_boxes"], G["gt_scores"], 200)
    assert np.array_equal(ob, G["gt_out_boxes"]) and np.array_equal(os_, G["gt_out_scores"])


def bbox_post_processing_oracle(cls_score, bbox, max_det, min_score, thr):
    """models/maskrcnn/bbox_post_processing.py:6-32 on top of the oracle's per-class NMS (np_ops.do_nms)."""
    B = cls_score.shape[0]
    score = np.zeros((B, max_det, 1), np.float32)
    box = np.zeros((B, max_det, 4), np.float32)
    cls = np.full((B, max_det, 1), -1, np.float32)
    for b in range(B):
        per = np_ops.do_nms(cls_score[b][:, 1:], bbox[b][:, 4:] if bbox.shape[2] != 4 else bbox[b], thr, min_score)
        rows = np.vstack([np.hstack((d, np.full((d.shape[0], 1), c, np.float32))) for c, d in per.items()])
        top = np.argsort(rows[:, 4])[::-1][:max_det]
        n = len(top)
        box[b, :n], score[b, :n, 0], cls[b, :n, 0] = rows[top, :4], rows[top, 4], rows[top, 5]
    return score, box, cls


def test_bbox_post_processing():
    s, b, c = bbox_post_processing_oracle(G["bp_cls_score"], G["bp_bbox"], 50, 0.3, 0.5)
    assert np.array_equal(s, G["bp_score"]) and np.array_equal(b, G["bp_box"]) and np.array_equal(c, G["bp_cls"])


def test_maskiou_compute_host_composition(monkeypatch):
    """ops.maskiou_compute (torch elementwise + reductions) against the reference's CustomOp run unmodified
    (tests/golden/make_golden_maskiou.py), on CPU tensors with only the CUDA tensor check lifted; its device twin is in
    tests/test_zz_late_gpu.py."""
    import torch

    from simpledet_b200 import ops

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_maskiou_compute.npz"))
    monkeypatch.setattr(ops, "_dev", lambda t, name, dtype=torch.float32: t.contiguous())
    iou, w = ops.OPS["maskiou_compute"](*[torch.from_numpy(g[k]) for k in ("logits", "target", "ratio", "inds")])
    assert np.array_equal(iou.numpy(), g["iou"]) and np.array_equal(w.numpy(), g["weight"])
    iou2, _ = ops.maskiou_compute(torch.from_numpy(g["logits"]), torch.from_numpy(g["target"]),
                                  torch.from_numpy(g["ratio"]).reshape(-1, 1), torch.from_numpy(g["inds"]))
    assert torch.equal(iou, iou2)
