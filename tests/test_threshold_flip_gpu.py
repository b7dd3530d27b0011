"""IoU-vs-threshold decisions at 1-ulp resolution.

A decoded Proposal box goes through `exp`, which CUDA and libm evaluate differently in the last ulps, so the
proposal BOXES are compared with a tolerance (tests/test_boxes_gpu.py) - a box that moved by 1 ulp could flip an
`IoU >= thr` decision downstream.  This file pins the decision itself:
  * NMS comparators on boxes whose IoU is EXACTLY representable: `>=` (Proposal_v3, greedy_nms) suppresses at
    IoU == thr, `>` (_contrib_NMS, gpu_nms, nms.py) keeps; thresholds one ulp either side flip it - on the device
    exactly as in the oracle;
  * Proposal_v3 with deltas whose exp is exact on both sides (dw = dh = 0) is bit-identical end to end, boxes
    included, so any difference elsewhere is the exp and nothing else;
  * on the random case the boxes differ by < 1e-3 px and no IoU of the pre-NMS set lies within 1e-5 of the
    threshold - the keep sets agree."""
import numpy as np
import pytest
import torch

import oracle
from simpledet_b200 import ops

pytestmark = pytest.mark.gpu


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _exact_iou_boxes():
    # +1 pixel convention: [0,0,9,9] has area 100.  IoU with the first box: 50/100 = 0.5, 25/100 = 0.25, 75/100 = 0.75
    return np.array([[0, 0, 9, 9, 0.9], [0, 0, 9, 4, 0.8], [0, 0, 4, 4, 0.7], [100, 100, 109, 109, 0.6],
                     [100, 100, 109, 106.5, 0.5]], np.float32)  # last: inter 75, union 100 -> 0.75


@pytest.mark.parametrize("thr_base", [0.5, 0.25, 0.75])
def test_nms_comparators_at_exact_iou(cuda, thr_base):
    dets = _exact_iou_boxes()
    for thr in (np.nextafter(np.float32(thr_base), np.float32(0)), np.float32(thr_base),
                np.nextafter(np.float32(thr_base), np.float32(1))):
        for ge in (True, False):
            keep, nkeep = ops.nms_sorted(_t(dets[None], cuda), float(thr), ge=ge)
            got = keep[0, :int(nkeep[0])].cpu().numpy().tolist()
            # oracle: sequential greedy NMS with the same comparator, float32 IoU with the +1 convention
            want = []
            for i in range(len(dets)):
                ok = True
                for j in want:
                    iw = min(dets[i, 2], dets[j, 2]) - max(dets[i, 0], dets[j, 0]) + np.float32(1)
                    ih = min(dets[i, 3], dets[j, 3]) - max(dets[i, 1], dets[j, 1]) + np.float32(1)
                    if iw > 0 and ih > 0:
                        inter = np.float32(iw * ih)
                        a_i = np.float32((dets[i, 2] - dets[i, 0] + 1) * (dets[i, 3] - dets[i, 1] + 1))
                        a_j = np.float32((dets[j, 2] - dets[j, 0] + 1) * (dets[j, 3] - dets[j, 1] + 1))
                        iou = np.float32(inter / np.float32(a_i + a_j - inter))
                        if (iou >= thr) if ge else (iou > thr):
                            ok = False
                            break
                if ok:
                    want.append(i)
            assert got == want, (thr, ge, got, want)
    # the flip itself: at IoU == thr the two comparators disagree, one ulp above thr they agree (both keep)
    k_ge, n_ge = ops.nms_sorted(_t(dets[None], cuda), float(thr_base), ge=True)
    k_gt, n_gt = ops.nms_sorted(_t(dets[None], cuda), float(thr_base), ge=False)
    assert int(n_gt[0]) > int(n_ge[0])


def test_proposal_v3_is_bit_exact_when_exp_is_exact(cuda):
    """dw = dh = 0 everywhere: exp(0) = 1 on both sides, so boxes, scores and the NMS decisions are bit-identical;
    dx, dy are random."""
    rng = np.random.default_rng(3)
    B, A, H, W = 2, 3, 25, 42
    cls = rng.random((B, 2 * A, H, W), dtype=np.float32)
    deltas = (rng.standard_normal((B, 4 * A, H, W)) * 0.2).astype(np.float32)
    deltas[:, 2::4] = 0
    deltas[:, 3::4] = 0
    im_info = np.array([[800, 1333, 1.0], [700, 1200, 1.1]], np.float32)
    kw = dict(feature_stride=32, scales=(8,), ratios=(0.5, 1, 2), rpn_pre_nms_top_n=1000, rpn_post_nms_top_n=300,
              threshold=0.7, rpn_min_size=0)
    rb, rs = oracle.proposal_v3(cls, deltas, im_info, **kw)
    gb, gs = ops.Proposal_v3(_t(cls, cuda), _t(deltas, cuda), _t(im_info, cuda), output_score=True, **kw)
    assert np.array_equal(gs.cpu().numpy(), rs)
    assert np.array_equal(gb.cpu().numpy(), rb)


def test_random_proposals_have_no_iou_near_the_threshold(cuda):
    """With random deltas the device boxes differ from the oracle's by the exp's last ulps only (< 1e-3 px); the keep
    sets agree because no pair's IoU sits within 1e-5 of the threshold."""
    rng = np.random.default_rng(4)
    B, A, H, W = 1, 3, 50, 84
    cls = rng.random((B, 2 * A, H, W), dtype=np.float32)
    deltas = (rng.standard_normal((B, 4 * A, H, W)) * 0.3).astype(np.float32)
    im_info = np.array([[800, 1333, 1.0]], np.float32)
    kw = dict(feature_stride=16, scales=(8,), ratios=(0.5, 1, 2), rpn_pre_nms_top_n=2000, rpn_post_nms_top_n=1000,
              threshold=0.7, rpn_min_size=0)
    rb, rs = oracle.proposal_v3(cls, deltas, im_info, **kw)
    gb, gs = ops.Proposal_v3(_t(cls, cuda), _t(deltas, cuda), _t(im_info, cuda), output_score=True, **kw)
    assert np.array_equal(gs.cpu().numpy(), rs)           # same boxes survive, in the same order
    assert np.abs(gb.cpu().numpy() - rb).max() < 1e-3
    kept = rb[0][rs[0, :, 0] > 0][:400]
    iou = ops.bbox_overlaps(_t(kept, cuda), _t(kept, cuda)).cpu().numpy()
    near = np.abs(iou - 0.7) < 1e-5
    assert not near.any()
