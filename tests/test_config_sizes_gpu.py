"""Parity at the FULL sizes of BASELINE.json configs[2..4]: the oracle's C restatements
finish these in seconds, so the comparison is direct; where it is not (DCN's numpy oracle) a sampled
subset plus size-independent properties are used."""
import numpy as np
import pytest
import torch

import oracle
from oracle import np_ops
from simpledet_b200 import ops

pytestmark = pytest.mark.gpu


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def test_config2_retina_focal_full_size(cuda):
    """retina_r50v1_fpn_1x: data (2, 200700, 80), label (2, 200700)."""
    rng = np.random.default_rng(202)
    B, N, K = 2, 200700, 80
    data = (rng.standard_normal((B, N, K), dtype=np.float32) * 2 - 3)
    label = np.zeros((B, N), np.float32)
    for b in range(B):
        pos = rng.choice(N, 400, replace=False)
        label[b, pos] = rng.integers(1, K + 1, 400)
        label[b, rng.choice(N, 5000, replace=False)] = -1
    d = _t(data, cuda).requires_grad_(True)
    out = ops.FocalLoss(d, _t(label, cuda), alpha=0.25, gamma=2.0, normalization="valid", grad_scale=1.0)
    out.backward(torch.ones_like(out))
    g = d.grad.cpu().numpy()
    rg = oracle.focal_loss_backward(out.detach().cpu().numpy(), label, 0.25, 2.0, 1.0, "valid", None)
    np.testing.assert_allclose(g, rg, rtol=1e-4, atol=1e-9)
    assert (g[label == -1] == 0).all() and np.isfinite(g).all()


def test_config2_retina_proposals_p3_full_size(cuda):
    """GenAnchor + GenProposalRetina on P3 of an 800x1333 image: 720 x 100 x 167 = 12 M (anchor, class) pairs."""
    rng = np.random.default_rng(203)
    B, A, K, H, W, stride = 2, 9, 80, 100, 167, 8
    cls = (rng.random((B, A * K, H, W), dtype=np.float32) ** 6)          # sigmoid-like: mostly tiny
    deltas = (rng.standard_normal((B, 4 * A, H, W), dtype=np.float32) * 0.3)
    im_info = np.array([[800, 1333, 1.0], [768, 1280, 1.2]], np.float32)
    scales = tuple(4 * 2 ** (i / 3) for i in range(3))
    anchors_gpu = ops.GenAnchor(_t(cls[:1, :9], cuda), scales=scales, ratios=(0.5, 1, 2), feature_stride=stride)
    anchors = oracle.gen_anchor(H, W, stride, scales, (0.5, 1, 2))
    assert np.array_equal(anchors_gpu.cpu().numpy(), anchors)
    kw = dict(num_anchors=A, rpn_pre_nms_top_n=1000, rpn_min_size=0, thresh=0.05, anchor_mean=(0, 0, 0, 0),
              anchor_std=(1, 1, 1, 1))
    rb, rs = oracle.gen_proposal_retina(cls, deltas, im_info, anchors, **kw)
    gb, gs = ops.GenProposalRetina(_t(cls, cuda), _t(deltas, cuda), _t(im_info, cuda), anchors_gpu,
                                   feature_stride=stride, **kw)
    assert np.array_equal(gs.cpu().numpy(), rs)
    np.testing.assert_allclose(gb.cpu().numpy(), rb, rtol=1e-5, atol=1e-3)
    top = rs.max(-1)
    assert (np.diff(top, axis=1) <= 0).all() and (top[:, -1] > 0.05).all()   # 1000 survivors, sorted


def test_config3_mask_target_full_size(cuda):
    """mask_r50v1_fpn_1x: rois (2, 2000, 4), gt (2, 100, 5), polys (2, 100, 2500), 512 rois/img, 28x28."""
    from tests.test_mask_target_gpu import _scene
    rng = np.random.default_rng(204)
    B, R, G, PL, IR, M = 2, 2000, 100, 2500, 512, 28
    rois, gt, polys = _scene(rng, B, R, G, PL)
    pr = rng.integers(0, 2 ** 32, (B, 4, R + G), dtype=np.uint64).astype(np.uint32)
    ref = oracle.proposal_mask_target(rois, gt, polys, pr, 81, IR, M, fg_fraction=0.25, fg_thresh=0.5,
                                      bg_thresh_hi=0.5, bg_thresh_lo=0.0)
    res = ops.ProposalMaskTarget(_t(rois, cuda), _t(gt, cuda), _t(polys, cuda), 81, B, IR, M, 0.5, 0.5, 0.0, False,
                                 priorities=_t(pr.astype(np.int64), cuda))
    for k in range(4):
        a, b_ = res[k].cpu().numpy(), ref[k]
        assert np.array_equal(a, b_) if k < 2 else np.allclose(a, b_, rtol=1e-5, atol=1e-6), k
    assert res[4].shape == (B, 128, M, M) and np.array_equal(res[4].cpu().numpy(), ref[5])


def test_config4_dcn_c4_and_soft_nms_full_size(cuda):
    """dcn faster_r50v1_fpn + soft-NMS: DCN on (2, 256, 50, 84) with 4 deformable groups; soft-NMS linear on
    80 classes x 1000 boxes x 2 images."""
    rng = np.random.default_rng(205)
    B, C, H, W, dg = 2, 256, 50, 84, 4
    data = rng.standard_normal((B, C, H, W), dtype=np.float32)
    offset = (rng.standard_normal((B, dg * 18, H, W), dtype=np.float32) * 2)
    weight = (rng.standard_normal((C, C, 3, 3), dtype=np.float32) * 0.02)
    y = ops.DeformableConvolution(_t(data, cuda), _t(offset, cuda), _t(weight, cuda), kernel=(3, 3), pad=(1, 1),
                                  num_filter=C, num_deformable_group=dg, no_bias=True)
    sub = [0, 63, 64, 200, 255]                                    # channels of three different groups
    col = np_ops.deformable_im2col(data[:, sub], offset.reshape(B, dg, 18, H, W)[:, [0, 0, 1, 3, 3]].reshape(B, 5 * 18, H, W),
                                   (3, 3), (1, 1), (1, 1), (1, 1), 5)
    # the op's columns for those channels, recovered through a one-hot weight
    pick = np.zeros((len(sub) * 9, C, 3, 3), np.float32)
    for i, c in enumerate(sub):
        for t in range(9):
            pick[i * 9 + t, c, t // 3, t % 3] = 1
    got = ops.DeformableConvolution(_t(data, cuda), _t(offset, cuda), _t(pick, cuda), kernel=(3, 3), pad=(1, 1),
                                    num_deformable_group=dg, no_bias=True)
    np.testing.assert_allclose(got.reshape(B, len(sub) * 9, H * W).cpu().numpy(), col, rtol=1e-5, atol=1e-5)
    assert y.shape == (B, C, H, W) and torch.isfinite(y).all()
    P, m = 160, 1000
    dets = np.zeros((P, m, 5), np.float32)
    for p in range(P):
        xy = rng.uniform(0, 1100, (m, 2))
        dets[p] = np.concatenate([xy, xy + rng.uniform(8, 300, (m, 2)), rng.permutation(m)[:, None] / m + 1e-3], 1)
    counts = rng.integers(0, m + 1, P).astype(np.int32)
    ob, oi, oc = ops.soft_nms_batched(_t(dets, cuda), 0.5, 0.5, 0.001, 1, counts=_t(counts, cuda))
    ob, oi, oc = ob.cpu().numpy(), oi.cpu().numpy(), oc.cpu().numpy()
    for p in range(0, P, 7):
        rb, ri = oracle.soft_nms(dets[p, : counts[p]], 0.5, 0.5, 0.001, 1)
        assert oc[p] == len(ri) and np.array_equal(ob[p, : oc[p]], rb) and np.array_equal(oi[p, : oc[p]], ri), p
