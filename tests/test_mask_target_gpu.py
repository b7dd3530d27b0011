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


@pytest.mark.parametrize("M", [14, 28])
def test_mask_target_output_ratio(cuda, M):
    """Mask Scoring R-CNN form (output_iou + output_ratio, models/msrcnn/builder.py:219-239): seven outputs; the mask
    follows the ratio variant's double vertex transform and mask_ratio is the count ratio on the reference's integer
    rasters - computed on the device from sorted toggle positions, no raster is ever materialised.  The oracle's
    restatement is pinned against the reference operator (tests/test_oracle_ref_cxx.py)."""
    rng = np.random.default_rng(100 + M)
    B, R, G, PL, IR = 2, 400, 12, 2500, 128
    rois, gt, polys = _scene(rng, B, R, G, PL)
    if M == 28:
        rois = (np.round(rois * 4) / 4).astype(np.float32)    # quarter-pixel corners, some negative: int() truncates
    pr = rng.integers(0, 2 ** 32, (B, 4, R + G), dtype=np.uint64).astype(np.uint32)
    ref = oracle.proposal_mask_target(rois, gt, polys, pr, 81, IR, M, fg_fraction=0.25, fg_thresh=0.5,
                                      bg_thresh_hi=0.5, bg_thresh_lo=0.0, output_ratio=True)
    res = ops.ProposalMaskTarget(_t(rois, cuda), _t(gt, cuda), _t(polys, cuda), 81, B, IR, M, 0.5, 0.5, 0.0, False,
                                 output_iou=True, output_ratio=True, priorities=_t(pr.astype(np.int64), cuda))
    assert len(res) == 7 and res[5].shape == (B, 32, M, M) and res[6].shape == (B, 32)
    for i in (0, 1, 3, 4):
        assert np.array_equal(res[i].cpu().numpy(), ref[i]), i
    np.testing.assert_allclose(res[2].cpu().numpy(), ref[2], rtol=1e-5, atol=1e-6)   # log() targets: 1 ulp (as elsewhere)
    assert np.array_equal(res[5].cpu().numpy(), ref[5])
    got, want = res[6].cpu().numpy(), ref[6]
    assert np.array_equal(got, want), (got - want)[got != want]
    assert ((want > 0.05) & (want < 0.999)).sum() > 10          # real partial overlaps, not a trivial vector


def test_mask_ratio_overlapping_segments_and_empty_rows(cuda):
    """Two overlapping segments (the union is counted once), a polygon that sticks far out of the roi (small ratio),
    and rows past the foreground count (ratio 0, mask -1)."""
    sq = lambda x1, y1, x2, y2: [x1, y1, x2, y1, x2, y2, x1, y2]
    polys = np.full((1, 3, 60), -1, np.float32)
    row = [7, 2, 8, 8] + sq(100.3, 80.2, 220.7, 190.4) + sq(180.5, 150.1, 300.9, 260.6)
    polys[0, 0, :len(row)] = row
    tri = [40.2, 300.7, 900.4, 320.1, 470.3, 700.9]
    polys[0, 1, :3 + len(tri)] = [9, 1, len(tri)] + tri
    gt = np.full((1, 3, 5), -1, np.float32)
    gt[0, 0] = [100.3, 80.2, 300.9, 260.6, 7]
    gt[0, 1] = [40.2, 300.7, 900.4, 700.9, 9]
    rois = np.zeros((1, 8, 4), np.float32)
    rois[0, 0] = [120.5, 95.25, 290.0, 240.75]
    rois[0, 1] = [60.0, 310.0, 700.5, 690.0]
    rois[0, 2] = [98.0, 77.0, 303.0, 262.0]
    pr = np.zeros((1, 3, 11), np.uint32)
    pr[0, :, :] = np.arange(11)
    kw = dict(fg_fraction=0.5, fg_thresh=0.5, bg_thresh_hi=0.5, bg_thresh_lo=0.0)
    ref = oracle.proposal_mask_target(rois, gt, polys, pr, 81, 16, 28, output_ratio=True, **kw)
    res = ops.ProposalMaskTarget(_t(rois, cuda), _t(gt, cuda), _t(polys, cuda), 81, 1, 16, 28, 0.5, 0.5, 0.0, False,
                                 fg_fraction=0.5, output_iou=True, output_ratio=True,
                                 priorities=_t(pr.astype(np.int64), cuda))
    assert np.array_equal(res[0].cpu().numpy(), ref[0])
    assert np.array_equal(res[5].cpu().numpy(), ref[5]) and np.array_equal(res[6].cpu().numpy(), ref[6])
    r = ref[6][0]
    nfg = int((ref[1][0] > 0).sum())
    assert nfg == 5 and r.shape == (8,) and (r[:nfg] > 0.3).all() and (r[:nfg] <= 1).all() and not r[nfg:].any()
    assert (ref[5][0, nfg:] == -1).all()
