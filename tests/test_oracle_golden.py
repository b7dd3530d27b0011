"""Pins the oracle: (1) the reference's only known-answer vector for this path — the ROIPooling
docstring example (operator_cxx/roi_pooling_v1.cc:265-285); (2) a second independent pure-Python
restatement of ROIAlign_v2 (tests/pyref.py) on a tiny case; (3) structural properties."""
import os

import numpy as np
import pytest

import oracle
import oracle.np_ops
from tests import pyref


def test_roipooling_docstring_example():
    x = np.arange(48, dtype=np.float32).reshape(1, 1, 8, 6)
    rois = np.array([[0, 0, 0, 4, 4]], np.float32)
    out, idx = oracle.roi_pool_v1_forward(x, rois, (2, 2), 1.0)
    assert out.reshape(2, 2).tolist() == [[14.0, 16.0], [26.0, 28.0]]
    assert idx.reshape(2, 2).tolist() == [[14.0, 16.0], [26.0, 28.0]]  # ramp: value == index
    out, _ = oracle.roi_pool_v1_forward(x, rois, (2, 2), 0.7)
    assert out.reshape(2, 2).tolist() == [[7.0, 9.0], [19.0, 21.0]]


def test_roi_align_oracle_matches_python_restatement():
    rng = np.random.default_rng(3)
    data = rng.standard_normal((2, 3, 9, 11)).astype(np.float32)
    rois = np.array([[[8, 4, 100, 90], [0, 0, 0, 0], [30.5, 20.25, 31, 21], [-50, -40, 400, 300],
                      [16, 16, 48, 48]],
                     [[5, 5, 6, 6], [160, 100, 175, 140], [3.3, 7.7, 90.1, 60.9], [170, 0, 176, 144],
                      [0, 140, 170, 143]]], np.float32)
    out, ax, ay = oracle.roi_align_v2_forward(data, rois, (3, 4), 1 / 16)
    N = rois.shape[1]
    for n in range(2 * N):
        for c in range(3):
            for ph in range(3):
                for pw in range(4):
                    v, x, y = pyref.roi_align_v2_element(data, rois, N, 3, 4, 1 / 16, n, c, ph, pw)
                    idx = (n // N, n % N, c, ph, pw)
                    assert out[idx] == v and ax[idx] == x and ay[idx] == y, (idx, out[idx], v)


def test_roi_align_zero_roi_is_all_empty():
    data = np.ones((1, 2, 6, 6), np.float32)
    out, ax, ay = oracle.roi_align_v2_forward(data, np.zeros((1, 2, 4), np.float32), (7, 7), 0.25)
    assert (out == 0).all() and (ax == -1).all() and (ay == -1).all()


def test_roi_align_constant_map_gives_constant():
    data = np.full((1, 1, 20, 20), 3.0, np.float32)
    rois = np.array([[[8, 8, 60, 70]]], np.float32)
    out, ax, ay = oracle.roi_align_v2_forward(data, rois, (7, 7), 0.25)
    np.testing.assert_allclose(out, 3.0, rtol=1e-6)
    assert (ax >= 2).all() and (ax <= 15).all() and (ay >= 2).all() and (ay <= 17.5).all()


def test_roi_align_backward_conserves_mass():
    """Each output gradient is scattered with 4 bilinear weights that sum to 1."""
    data, rois = np.random.default_rng(0).standard_normal((1, 4, 12, 12)).astype(np.float32), None
    rois = np.array([[[4, 4, 40, 40], [0, 0, 0, 0], [10, 2, 30, 44]]], np.float32)
    out, ax, ay = oracle.roi_align_v2_forward(data, rois, (4, 4), 0.25)
    g = np.random.default_rng(1).standard_normal(out.shape).astype(np.float32)
    grad = oracle.roi_align_v2_backward(g, ax, ay, data.shape)
    np.testing.assert_allclose(grad.sum(), g[ax != -1].sum(), rtol=1e-4)
    acc = oracle.roi_align_v2_backward(g, ax, ay, data.shape, accumulate_into=np.ones_like(data))
    np.testing.assert_allclose(acc, grad + 1, rtol=1e-6, atol=1e-6)


def test_fpn_assign_levels_boundaries():
    # sqrt(area) with the +1 convention: w = x2-x1+1.  224 -> k0=4 -> stride 16 (index 2)
    def box(side):
        return [0, 0, side - 1, side - 1]

    rois = np.array([box(16), box(111), box(112), box(223), box(224), box(447), box(448), box(2000)],
                    np.float32)
    idx = oracle.fpn_assign_levels(rois, (4, 8, 16, 32))
    assert idx.tolist() == [0, 0, 1, 1, 2, 2, 3, 3]
    # numpy float32 restatement of the mx.nd expression
    r = np.random.default_rng(0).uniform(0, 800, (1000, 4)).astype(np.float32)
    r[:, 2:] += r[:, :2]
    area = (r[:, 2] - r[:, 0] + np.float32(1)) * (r[:, 3] - r[:, 1] + np.float32(1))
    lv = np.clip(np.floor(np.float32(4) + np.log2(np.sqrt(area) / np.float32(224) + np.float32(1e-6))),
                 2, 5)
    assert (oracle.fpn_assign_levels(r, (4, 8, 16, 32)) == (lv - 2).astype(np.int32)).all()


def test_roi_pool_backward_routes_to_argmax():
    rng = np.random.default_rng(5)
    data = rng.standard_normal((2, 3, 10, 12)).astype(np.float32)
    rois = np.array([[0, 0, 0, 20, 18], [1, 4, 4, 23, 19], [1, 30, 30, 2, 2]], np.float32)
    out, idx = oracle.roi_pool_v1_forward(data, rois, (3, 3), 0.5)
    flat = data.reshape(2, 3, -1)
    for r in range(3):
        b = int(rois[r, 0])
        for c in range(3):
            for k, a in enumerate(idx[r, c].ravel()):
                if a >= 0:
                    assert flat[b, c, int(a)] == out[r, c].ravel()[k]
    g = np.ones_like(out)
    grad = oracle.roi_pool_v1_backward(g, idx, rois, data.shape)
    assert grad.sum() == (idx >= 0).sum()


# ---- AnchorTarget2D / PyramidAnchorTarget2D: oracle restatement vs the reference classes (DEBUG mode) ----
_A2D = dict(strides=(16,), shorts=(12,), longs=(18,), scales=(2, 4, 8), aspects=(0.5, 1.0, 2.0), allowed_border=0,
            neg_thr=0.3, pos_thr=0.7, min_pos_thr=0.0, image_anchor=64, pos_fraction=0.5)
_P2D = dict(strides=(4, 8, 16, 32), shorts=(40, 20, 10, 5), longs=(60, 30, 15, 8), scales=(8,),
            aspects=(0.5, 1.0, 2.0), allowed_border=9999, neg_thr=0.3, pos_thr=0.7, min_pos_thr=0.0,
            image_anchor=256, pos_fraction=0.5)
ANCHOR_CASES = [("a2d_h", _A2D), ("a2d_v", _A2D), ("a2d_fgcap", dict(_A2D, image_anchor=8, pos_thr=0.5)),
                ("p2d_h", _P2D), ("p2d_v", _P2D), ("p2d_empty", _P2D)]


@pytest.mark.parametrize("tag,cfg", ANCHOR_CASES)
def test_anchor_target_matches_reference_classes(tag, cfg):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_anchor_target.npz"))
    lab, tgt, wgt = oracle.np_ops.anchor_target(g[f"{tag}_im_info"], g[f"{tag}_gt"], **cfg)
    assert np.array_equal(lab, g[f"{tag}_label"])
    assert np.array_equal(wgt.reshape(-1), g[f"{tag}_weight"].reshape(-1))
    assert np.array_equal(tgt.reshape(-1), g[f"{tag}_target"].reshape(-1))
