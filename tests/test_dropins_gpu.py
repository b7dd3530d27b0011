"""Drop-in callables and CustomOp twins on the device (SURVEY §8b secondary APIs) against the oracle / goldens."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import np_ops
from simpledet_b200 import ops, synth
from test_customops_golden import G, bbox_post_processing_oracle

pytestmark = pytest.mark.gpu


def test_ops_registry_names():
    for name in ("assign_layer_fpn", "get_top_proposal", "BboxPostProcessing", "gpu_nms", "greedy_nms",
                 "bbox_overlaps_cython", "soft_nms", "_contrib_ROIAlign_v2", "ProposalTarget", "decode_retina"):
        assert callable(ops.OPS[name]), name


def test_gpu_nms_symbol_through_python(cuda):
    rng = np.random.default_rng(0)
    n = 800
    xy = rng.uniform(0, 500, (n, 2))
    dets = np.concatenate([xy, xy + rng.uniform(10, 150, (n, 2)), rng.permutation(n)[:, None] / n], 1).astype(np.float32)
    keep = ops.gpu_nms(dets, 0.5)
    want = np_ops.py_nms(dets, 0.5)  # nms.py keeps ovr <= thresh == nms_kernel.cu's `>` suppression
    assert np.array_equal(dets[keep], want)


def test_greedy_nms(cuda):
    rng = np.random.default_rng(1)
    n = 600
    xy = rng.uniform(0, 400, (n, 2))
    dets = np.concatenate([xy, xy + rng.uniform(10, 150, (n, 2)), rng.permutation(n)[:, None] / n], 1).astype(np.float32)
    got = ops.greedy_nms(torch.from_numpy(dets).to(cuda), 0.5).cpu().numpy()
    assert np.array_equal(got, oracle.greedy_nms(dets, 0.5))


def test_assign_layer_fpn(cuda):
    rois = G["al_rois"]
    outs, lv = ops.assign_layer_fpn(torch.from_numpy(rois).to(cuda), return_levels=True)
    for i, o in enumerate(outs):
        assert np.array_equal(o.cpu().numpy(), G[f"al_out{i}"]), i
    assert np.array_equal(lv.cpu().numpy(), oracle.fpn_assign_levels(rois, (4, 8, 16, 32)).reshape(rois.shape[:2]))


def test_get_top_proposal_golden(cuda):
    ob, os_ = ops.get_top_proposal(torch.from_numpy(G["gt_boxes"]).to(cuda), torch.from_numpy(G["gt_scores"]).to(cuda), 200)
    assert np.array_equal(ob.cpu().numpy(), G["gt_out_boxes"]) and np.array_equal(os_.cpu().numpy(), G["gt_out_scores"])


def test_bbox_post_processing(cuda):
    s, b, c = ops.BboxPostProcessing(torch.from_numpy(G["bp_cls_score"]).to(cuda), torch.from_numpy(G["bp_bbox"]).to(cuda),
                                     max_det_per_image=50, min_det_score=0.3, nms_thr=0.5)
    assert np.array_equal(s.cpu().numpy(), G["bp_score"])
    assert np.array_equal(b.cpu().numpy(), G["bp_box"])
    assert np.array_equal(c.cpu().numpy(), G["bp_cls"])


def test_decode_retina_against_the_reference_customop(cuda):
    """models/retinanet/decode_retina.py run unmodified produced tests/golden/reference_decode_retina.npz.  The
    reference leaves each level's rows in np.argpartition order: compare the padded outputs as sets of rows, level by
    level (a level's rows are contiguous; levels with no candidate contribute none), boxes to float32 rounding of
    the float32 exp."""
    import os

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_decode_retina.npz"))
    L = len(g["stride"])
    cls = [torch.from_numpy(g[f"cls{i}"]).to(cuda) for i in range(L)]
    reg = [torch.from_numpy(g[f"reg{i}"]).to(cuda) for i in range(L)]
    boxes, scores = ops.OPS["decode_retina"](cls, reg, torch.from_numpy(g["im_info"]).to(cuda),
                                             stride=tuple(int(s) for s in g["stride"]), scales=tuple(g["scales"]),
                                             ratios=tuple(g["ratios"]), per_level_top_n=int(g["top"]),
                                             thresh=float(g["thresh"]))
    b, s = boxes.cpu().numpy()[0], scores.cpu().numpy()[0]
    rb, rs = g["boxes"][0], g["scores"][0]
    assert b.shape == rb.shape and s.shape == rs.shape
    n = int((rs.sum(-1) > 0).sum())
    assert int((s.sum(-1) > 0).sum()) == n and not s[n:].any() and not b[n:].any()
    assert not s[:, 0].any()                                # no background scores
    top, at = int(g["top"]), 0
    for i in range(L):                                      # rows of level i: how many passed its threshold
        thr = float(g["thresh"]) if g["stride"][i] != g["stride"].max() else 0.0
        k = min(top, int((g[f"cls{i}"] > thr).sum()))
        key = lambda bb, ss: np.lexsort((bb[:, 3], bb[:, 2], bb[:, 1], bb[:, 0], ss.argmax(-1), ss.max(-1)))
        mine, ref = key(b[at:at + k], s[at:at + k]), key(rb[at:at + k], rs[at:at + k])
        np.testing.assert_array_equal(s[at:at + k][mine], rs[at:at + k][ref])
        np.testing.assert_allclose(b[at:at + k][mine], rb[at:at + k][ref], rtol=2e-6, atol=2e-4)
        at += k
    assert at == n
