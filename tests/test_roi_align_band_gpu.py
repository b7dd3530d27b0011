"""Band-stationary RoIAlign_v2 forward (roi_align_band.cu: bulk-TMA staged feature bands, persistent CTAs)
against the oracle and against the per-roi kernel, through the C ABI.  Everything is BIT-EXACT: both
kernels round like the reference's CPU build (roi_align_v2-inl.h:61-153)."""
import numpy as np
import pytest
import torch

import oracle
from simpledet_b200 import ops, synth

pytestmark = pytest.mark.gpu

BAND = 2  # path_used code of the band-stationary kernel


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _single(data, rois, pooled, scale, dev, expect_band=True):
    d, r = _t(data, dev), _t(rois, dev)
    out, _, _, used = ops.roi_align_v2_raw(d, r, pooled, scale, with_argmax=False, path=2, return_path=True)
    if expect_band:
        assert used == BAND, f"band path not taken (path_used={used})"
    else:
        assert used == 1
    ref, _, _ = oracle.roi_align_v2_forward(data, rois, pooled, scale)
    o = out.cpu().numpy()
    assert np.array_equal(o, ref), f"band kernel differs from the oracle at {np.argwhere(o != ref)[:5]}"
    per, _, _, used1 = ops.roi_align_v2_raw(d, r, pooled, scale, with_argmax=False, path=1, return_path=True)
    assert used1 == 1
    assert torch.equal(per, out)


def test_config1_band(cuda):
    data, rois, pooled, scale = synth.config1(0)
    _single(data, rois, pooled, scale, cuda)


@pytest.mark.parametrize("pooled", [(7, 7), (14, 14), (3, 5), (1, 1), (16, 16)])
@pytest.mark.parametrize("C", [4, 36, 64])
def test_random_shapes_band(cuda, pooled, C):
    rng = np.random.default_rng(C * 100 + pooled[0])
    data = rng.standard_normal((2, C, 50, 84)).astype(np.float32)
    rois = synth.random_rois(rng, 2, 60, 800, 1333)
    _single(data, rois, pooled, 1 / 16, cuda)


def test_edge_cases_band(cuda):
    """The edge set of test_roi_align_gpu.py on a level whose planes are only 8-byte aligned
    (25*42 % 4 == 2: odd channels are staged with an 8-byte shift)."""
    rng = np.random.default_rng(7)
    data = rng.standard_normal((1, 8, 25, 42)).astype(np.float32)
    rois = np.array([[
        [0, 0, 0, 0], [-500, -400, -100, -50], [5000, 4000, 6000, 5000], [0, 0, 1343, 799],
        [96, 96, 96 + 7 * 48, 96 + 7 * 48], [100, 100, 100.2, 100.2], [100, 100, 101.5, 250],
        [64, 64, 64.96, 64.96], [300, 200, 100, 50], [np.nan, 10, 200, 300], [1200, 700, 1400, 900],
        [-30, -30, 60, 60], [10, 10, 700, 40], [10, 10, 40, 700],
    ]], np.float32)
    _single(data, rois, (7, 7), 1 / 32, cuda)
    _single(data, rois, (14, 14), 1 / 32, cuda)


def test_ties_band(cuda):
    data = np.ones((1, 4, 50, 50), np.float32)
    data[:, 1] = 0.0
    data[:, 2, ::2] = 2.0
    data[:, 3] = -1.0
    rois = synth.random_rois(np.random.default_rng(2), 1, 64, 800, 800)
    _single(data, rois, (7, 7), 1 / 16, cuda)


def test_crowded_band_many_groups(cuda):
    """400 rois on a small map: every band holds far more items than one unit's table (30), so
    bands are staged once per item group; images > 1."""
    rng = np.random.default_rng(11)
    data = rng.standard_normal((3, 16, 30, 40)).astype(np.float32)
    rois = synth.random_rois(rng, 3, 400, 480, 640, min_side=24, max_side=300)
    _single(data, rois, (7, 7), 1 / 16, cuda)
    _single(data, rois, (14, 14), 1 / 16, cuda)


def test_unsupported_levels_fall_back(cuda):
    """Odd H*W (planes not 8-byte aligned) and very wide maps are served by the per-roi kernel."""
    rng = np.random.default_rng(5)
    data = rng.standard_normal((1, 8, 25, 25)).astype(np.float32)
    rois = synth.random_rois(rng, 1, 30, 400, 400)
    _single(data, rois, (7, 7), 1 / 16, cuda, expect_band=False)
    data = rng.standard_normal((1, 4, 12, 400)).astype(np.float32)
    rois = synth.random_rois(rng, 1, 30, 96, 3200)
    _single(data, rois, (7, 7), 1 / 8, cuda, expect_band=False)


def _fpn(B, N, C, pooled, dev, seed, check_oracle):
    rng = np.random.default_rng(seed)
    shapes = synth.fpn_shapes()
    feats_np = [rng.standard_normal((B, C, h, w)).astype(np.float32) for h, w in shapes]
    rois_np = synth.random_rois(rng, B, N)
    feats = [_t(f, dev) for f in feats_np]
    rois = _t(rois_np, dev)
    out, _, _, lv, used = ops.fpn_roi_align_raw(feats, rois, synth.FPN_STRIDES, pooled, with_argmax=False, path=2,
                                                return_path=True)
    assert used == BAND
    per, _, _, lv1, used1 = ops.fpn_roi_align_raw(feats, rois, synth.FPN_STRIDES, pooled, with_argmax=False, path=1,
                                                  return_path=True)
    assert used1 == 1
    assert torch.equal(lv, lv1)
    assert torch.equal(out, per), "band kernel differs from the per-roi kernel"
    if check_oracle:
        ref, rl = oracle.fpn_roi_align_v2_forward(feats_np, rois_np, synth.FPN_STRIDES, (pooled, pooled))
        assert np.array_equal(lv.cpu().numpy(), rl)
        assert np.array_equal(out.cpu().numpy(), ref)


def test_fpn_target_shape_small_c(cuda):
    _fpn(1, 512, 16, 14, cuda, 0, True)


def test_fpn_bench_shape_small_c(cuda):
    _fpn(2, 1000, 8, 7, cuda, 1, True)


def test_fpn_target_shape_full(cuda):
    """North-star shape 512 rois x 256 ch x 14x14: band kernel == per-roi kernel (itself oracle-checked)."""
    _fpn(1, 512, 256, 14, cuda, 0, False)


def test_fpn_bench_shape_full(cuda):
    _fpn(2, 1000, 256, 7, cuda, 1, False)
