"""AnchorTarget2D / PyramidAnchorTarget2D on the GPU vs (a) the goldens produced by the reference
classes themselves (tests/golden/reference_anchor_target.npz, DEBUG sub-sampling) and (b) the oracle
restatement with injected priorities at the real FPN size."""
import os

import numpy as np
import pytest
import torch

import oracle
import oracle.np_ops
from simpledet_b200 import ops
from tests.test_oracle_golden import ANCHOR_CASES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


def _run(cuda, im_info, gt, cfg, priorities):
    kw = dict(cfg)
    st, sh, lo = kw.pop("strides"), kw.pop("shorts"), kw.pop("longs")
    pr = None if priorities is None else torch.from_numpy(priorities.astype(np.int64).astype(np.int32)).to(cuda)
    lab, tgt, wgt = ops.PyramidAnchorTarget2D(torch.from_numpy(im_info).to(cuda), torch.from_numpy(gt).to(cuda),
                                              st, sh, lo, priorities=pr, **kw)
    return lab.cpu().numpy(), tgt.cpu().numpy(), wgt.cpu().numpy()


@pytest.mark.parametrize("tag,cfg", ANCHOR_CASES)
def test_matches_reference_goldens(cuda, tag, cfg):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_anchor_target.npz"))
    n = g[f"{tag}_label"].size
    # batch of two copies: also checks the per-image indexing
    im = np.stack([g[f"{tag}_im_info"]] * 2)
    gt = np.stack([g[f"{tag}_gt"]] * 2)
    pr = np.tile(np.arange(n, dtype=np.int64), (2, 1))  # DEBUG mode: the first surplus indices are disabled
    lab, tgt, wgt = _run(cuda, im, gt, cfg, pr)
    for b in range(2):
        assert np.array_equal(lab[b], g[f"{tag}_label"])
        assert np.array_equal(wgt[b].reshape(-1), g[f"{tag}_weight"].reshape(-1))
        # float64 log on the device vs numpy's: equal after the float32 cast up to 1 ulp
        np.testing.assert_allclose(tgt[b].reshape(-1), g[f"{tag}_target"].reshape(-1), rtol=2e-7, atol=1e-7)


def test_fpn_full_size_random_priorities(cuda):
    """faster_r50v1_fpn_1x sizes: 5 levels, 267 069 anchors per image, ragged gt counts."""
    rng = np.random.default_rng(31)
    cfg = dict(strides=(4, 8, 16, 32, 64), shorts=(200, 100, 50, 25, 13), longs=(334, 167, 84, 42, 21), scales=(8,),
               aspects=(0.5, 1.0, 2.0), allowed_border=9999, neg_thr=0.3, pos_thr=0.7, min_pos_thr=0.0,
               image_anchor=256, pos_fraction=0.5)
    ims = np.array([[800, 1333, 1.6], [1333, 800, 1.6], [736, 1200, 1.5]], np.float32)
    G = 40
    gts = np.full((3, G, 5), -1, np.float32)
    for b, n in enumerate((40, 9, 0)):
        h, w = ims[b, :2]
        xy = rng.uniform(0, [w * 0.8, h * 0.8], (n, 2))
        wh = rng.uniform(16, [w * 0.4, h * 0.4], (n, 2))
        gts[b, :n, :4] = np.concatenate([xy, np.minimum(xy + wh, [w - 1, h - 1])], 1)
        gts[b, :n, 4] = 1
    gts[0, [3, 17]] = -1  # holes in the padded gt list
    N = 3 * sum(s * l for s, l in zip(cfg["shorts"], cfg["longs"]))
    pr = rng.integers(0, 2 ** 31 - 1, (3, N), dtype=np.int64)
    lab, tgt, wgt = _run(cuda, ims, gts, cfg, pr)
    for b in range(3):
        rl, rt, rw = oracle.np_ops.anchor_target(ims[b], gts[b], priorities=pr[b], **cfg)
        assert np.array_equal(lab[b], rl)
        assert np.array_equal(wgt[b], rw)
        np.testing.assert_allclose(tgt[b], rt, rtol=2e-7, atol=1e-7)
        assert (rl == 1).sum() + (rl == 0).sum() == 256


def test_philox_priorities_respect_quotas(cuda):
    rng = np.random.default_rng(5)
    cfg = dict(ANCHOR_CASES[2][1])  # fg-capped single level
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_anchor_target.npz"))
    im, gt = g["a2d_fgcap_im_info"][None], g["a2d_fgcap_gt"][None]
    full = oracle.np_ops.anchor_target(im[0], gt[0], **dict(cfg, image_anchor=10 ** 6))[0]  # no sub-sampling
    lab1 = _run(cuda, im, gt, cfg, None)[0][0]
    assert (lab1 == 1).sum() == 4 and (lab1 == 0).sum() == 4
    assert np.all(full[lab1 == 1] == 1) and np.all(full[lab1 == 0] == 0)  # a subset of the unsampled labels
    del rng
