"""ProposalMaskTarget: device rasteriser vs the oracle restatement of cocoapi's rleFrPoly +
convertPoly2Mask (parity unpinned: maskApi.c is not vendored in the reference).  Masks are integer
(0/1/-1) outputs: compared exactly, with the same injected shuffle priorities on both sides."""
import numpy as np
import pytest
import torch

import oracle
from simpledet_b200 import ops

pytestmark = pytest.mark.gpu


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


from simpledet_b200.synth import mask_scene as _scene  # noqa: E402


@pytest.mark.parametrize("M", [14, 28])
def test_mask_target_exact(cuda, M):
    rng = np.random.default_rng(M)
    B, R, G, PL, IR = 2, 400, 12, 2500, 128
    rois, gt, polys = _scene(rng, B, R, G, PL)
    pr = rng.integers(0, 2 ** 32, (B, 4, R + G), dtype=np.uint64).astype(np.uint32)
    ref = oracle.proposal_mask_target(rois, gt, polys, pr, 81, IR, M, fg_fraction=0.25, fg_thresh=0.5,
                                      bg_thresh_hi=0.5, bg_thresh_lo=0.0)
    res = ops.ProposalMaskTarget(_t(rois, cuda), _t(gt, cuda), _t(polys, cuda), 81, B, IR, M, 0.5, 0.5, 0.0, False,
                                 priorities=_t(pr.astype(np.int64), cuda))
    assert len(res) == 5 and res[4].shape == (B, 32, M, M)
    assert np.array_equal(res[0].cpu().numpy(), ref[0]) and np.array_equal(res[1].cpu().numpy(), ref[1])
    got, want = res[4].cpu().numpy(), ref[5]
    assert set(np.unique(got)) <= {-1.0, 0.0, 1.0}
    assert np.array_equal(got, want), f"{(got != want).sum()} mask pixels differ"
    assert (want == 1).sum() > 100  # the scene really has masks


def test_poly2mask_shapes(cuda):
    """Known shapes through the op: an axis-aligned half box and a triangle."""
    rois = np.zeros((1, 8, 4), np.float32)
    rois[0, 0] = [10, 20, 110, 120]
    gt = np.full((1, 2, 5), -1, np.float32)
    gt[0, 0] = [10, 20, 110, 120, 5]
    polys = np.full((1, 2, 40), -1, np.float32)
    polys[0, 0, :11] = [5, 1, 8, 10, 20, 60, 20, 60, 120, 10, 120, ][:11]
    pr = np.zeros((1, 3, 10), np.uint32)
    res = ops.ProposalMaskTarget(_t(rois, cuda), _t(gt, cuda), _t(polys, cuda), 81, 1, 4, 28, 0.5, 0.5, 0.0, False,
                                 priorities=_t(pr.astype(np.int64), cuda))
    m = res[4].cpu().numpy()[0, 0]
    want = oracle.poly2mask(rois[0, 0], polys[0, 0], 28)
    assert np.array_equal(m, want) and m[:, :14].sum() == 28 * 14 and m[:, 14:].sum() == 0


def test_mask_target_filter_scales(cuda):
    """TridentNet form (4 inputs): gt boxes outside valid_ranges are not appended as candidates."""
    rng = np.random.default_rng(9)
    B, R, G, PL, IR, M = 2, 300, 12, 2500, 64, 14
    rois, gt, polys = _scene(rng, B, R, G, PL)
    vr = np.array([[0, 150], [120, 1e5]], np.float32)
    pr = rng.integers(2 ** 20, 2 ** 32, (B, 4, R + G), dtype=np.uint64).astype(np.uint32)
    # candidates = the R-20 non-padding rois, then the appended gt boxes: those win every shuffle
    pr[:, :, R - 20:R - 20 + G] = np.arange(G).astype(np.uint32)  # (a shuffle orders by ascending priority)
    ref = oracle.proposal_mask_target(rois, gt, polys, pr, 81, IR, M, fg_fraction=0.25, fg_thresh=0.5,
                                      bg_thresh_hi=0.5, bg_thresh_lo=0.0, valid_ranges=vr, filter_scales=True)
    plain = oracle.proposal_mask_target(rois, gt, polys, pr, 81, IR, M, fg_fraction=0.25, fg_thresh=0.5,
                                        bg_thresh_hi=0.5, bg_thresh_lo=0.0)
    assert not np.array_equal(ref[0], plain[0])  # the filter changes the candidate set
    res = ops.ProposalMaskTarget(_t(rois, cuda), _t(gt, cuda), _t(polys, cuda), 81, B, IR, M, 0.5, 0.5, 0.0, False,
                                 priorities=_t(pr.astype(np.int64), cuda), filter_scales=True, num_args=4,
                                 valid_ranges=_t(vr, cuda))
    assert np.array_equal(res[0].cpu().numpy(), ref[0]) and np.array_equal(res[1].cpu().numpy(), ref[1])
    assert np.array_equal(res[4].cpu().numpy(), ref[5])


def test_polygon_with_more_edges_than_one_batch(cuda):
    """A 700-vertex polygon (COCO has them): poly_mask_kernel walks its edges in batches of 512 and hands the last
    point of a batch to the next one.  Exact against the oracle rasteriser."""
    nv = 700
    ang = np.linspace(0, 2 * np.pi, nv, endpoint=False)
    rad = 45 + 12 * np.sin(9 * ang) + 3 * np.cos(31 * ang)
    xs, ys = 160 + rad * np.cos(ang), 140 + rad * np.sin(ang) * 0.8
    PL = 2 + 1 + 2 * nv + 5
    polys = np.full((1, 2, PL), -1, np.float32)
    polys[0, 0, :3] = [3, 1, 2 * nv]
    polys[0, 0, 3:3 + 2 * nv] = np.stack([xs, ys], 1).reshape(-1)
    gt = np.full((1, 2, 5), -1, np.float32)
    gt[0, 0] = [xs.min(), ys.min(), xs.max(), ys.max(), 3]
    rois = np.zeros((1, 8, 4), np.float32)
    rois[0, 0] = gt[0, 0, :4] + [2, -3, -4, 5]
    pr = np.zeros((1, 3, 10), np.uint32)
    for M in (28, 56):
        res = ops.ProposalMaskTarget(_t(rois, cuda), _t(gt, cuda), _t(polys, cuda), 81, 1, 4, M, 0.5, 0.5, 0.0, False,
                                     priorities=_t(pr.astype(np.int64), cuda))
        kept = res[0].cpu().numpy()[0, 0]
        want = oracle.poly2mask(kept, polys[0, 0], M)
        got = res[4].cpu().numpy()[0, 0]
        assert np.array_equal(got, want) and 0.2 < want.mean() < 0.9
