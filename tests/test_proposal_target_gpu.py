"""ProposalTarget on the device vs the oracle.  With the SAME shuffle priorities injected on both
sides every output must be bit-identical except the two log() targets (CUDA logf vs glibc logf,
<= 2 ulp -> rtol 1e-5).  With the on-device Philox priorities the priorities are read back and fed
to the oracle (same bar), plus distribution / determinism invariants."""
import numpy as np
import pytest
import torch

import oracle
from simpledet_b200 import ops

pytestmark = pytest.mark.gpu


def _scene(rng, B, R, G, n_gt, n_valid, jitter=25.0):
    gt = np.full((B, G, 5), -1, np.float32)
    rois = np.zeros((B, R, 4), np.float32)
    for b in range(B):
        k = n_gt[b]
        xy = rng.uniform(0, 600, (k, 2))
        wh = rng.uniform(30, 250, (k, 2))
        gt[b, :k, :4] = np.concatenate([xy, xy + wh], 1)
        gt[b, :k, 4] = rng.integers(1, 81, k)
        m = n_valid[b]
        if k:
            near = gt[b, rng.integers(0, k, m // 2), :4] + rng.normal(0, jitter, (m // 2, 4))
        else:
            near = np.zeros((0, 4))
        far_xy = rng.uniform(0, 700, (m - len(near), 2))
        far = np.concatenate([far_xy, far_xy + rng.uniform(10, 300, (m - len(near), 2))], 1)
        allr = np.concatenate([near, far])
        allr[:, 3] = np.maximum(allr[:, 3], 1.0)  # valid rois have y2 > 0
        rois[b, :m] = allr[rng.permutation(m)]
    return rois, gt


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _compare(res, ref, IR):
    o_rois, o_lab, o_tgt, o_wgt, o_iou, kept = [x.cpu().numpy() for x in res]
    r_rois, r_lab, r_tgt, r_wgt, r_iou, r_kept = ref
    assert np.array_equal(kept, r_kept)
    assert np.array_equal(o_rois, r_rois) and np.array_equal(o_lab, r_lab)
    assert np.array_equal(o_iou, r_iou) and np.array_equal(o_wgt, r_wgt)
    assert np.array_equal(o_tgt != 0, r_tgt != 0)
    np.testing.assert_allclose(o_tgt, r_tgt, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("agnostic,without_gt", [(False, False), (True, False), (False, True)])
def test_injected_priorities_bit_exact(cuda, agnostic, without_gt):
    rng = np.random.default_rng(3)
    B, R, G, IR = 3, 2000, 100, 512
    rois, gt = _scene(rng, B, R, G, n_gt=[7, 40, 1], n_valid=[2000, 1500, 300])
    pr = rng.integers(0, 2 ** 32, (B, 5, R + G), dtype=np.uint64).astype(np.uint32)
    nc = 2 if agnostic else 81
    kw = dict(fg_fraction=0.25, fg_thresh=0.5, bg_thresh_hi=0.5, bg_thresh_lo=0.0,
              proposal_without_gt=without_gt, class_agnostic=agnostic)
    ref = oracle.proposal_target(rois, gt, pr, nc, IR, **kw)
    res = ops.ProposalTarget(_t(rois, cuda), _t(gt, cuda), nc, B, IR, kw["fg_thresh"], kw["bg_thresh_hi"],
                             kw["bg_thresh_lo"], without_gt, fg_fraction=0.25, class_agnostic=agnostic,
                             priorities=_t(pr.astype(np.int64), cuda), return_debug=True)
    _compare((res[0], res[1], res[2], res[3], res[4], res[5]), ref, IR)


def test_shortages_and_padding_rounds(cuda):
    """fg shortage (no shuffle, index order), bg shortage -> negative padding, tiny negative list ->
    many padding rounds, no gt at all, bg_lo > 0 so that bg != neg."""
    rng = np.random.default_rng(9)
    B, R, G, IR = 4, 64, 8, 48
    rois, gt = _scene(rng, B, R, G, n_gt=[2, 1, 0, 3], n_valid=[64, 5, 30, 20], jitter=4.0)
    pr = rng.integers(0, 2 ** 32, (B, 4, R + G), dtype=np.uint64).astype(np.uint32)
    kw = dict(fg_fraction=0.25, fg_thresh=0.5, bg_thresh_hi=0.5, bg_thresh_lo=0.1, proposal_without_gt=False,
              class_agnostic=False)
    ref = oracle.proposal_target(rois, gt, pr, 81, IR, **kw)
    res = ops.ProposalTarget(_t(rois, cuda), _t(gt, cuda), 81, B, IR, 0.5, 0.5, 0.1, False,
                             priorities=_t(pr.astype(np.int64), cuda), return_debug=True)
    _compare(res[:6], ref, IR)


def test_device_philox_sampling(cuda):
    rng = np.random.default_rng(5)
    B, R, G, IR = 2, 2000, 100, 512
    rois, gt = _scene(rng, B, R, G, n_gt=[12, 30], n_valid=[2000, 1800])
    args = (_t(rois, cuda), _t(gt, cuda), 81, B, IR, 0.5, 0.5, 0.0, False)
    r1 = ops.ProposalTarget(*args, seed=42, return_debug=True)
    r2 = ops.ProposalTarget(*args, seed=42, return_debug=True)
    r3 = ops.ProposalTarget(*args, seed=43, return_debug=True)
    assert all(torch.equal(a, b) for a, b in zip(r1, r2)), "same seed must reproduce"
    assert not torch.equal(r1[5], r3[5]), "a different seed must sample differently"
    # feed the priorities the kernel drew to the oracle: full parity under device RNG
    used = r1[6].cpu().numpy().astype(np.int64) & 0xFFFFFFFF
    ref = oracle.proposal_target(rois, gt, used.astype(np.uint32), 81, IR, 0.25, 0.5, 0.5, 0.0)
    _compare(r1[:6], ref, IR)
    # invariants of the reference's sampling
    lab, kept = r1[1].cpu().numpy(), r1[5].cpu().numpy()
    iou = r1[4].cpu().numpy()
    for b in range(B):
        nfg = int((lab[b] > 0).sum())
        assert nfg <= 128 and (lab[b][nfg:] == 0).all()
        assert (iou[b][:nfg] >= 0.5).all() and (iou[b][nfg:] < 0.5).all()
        assert len(set(kept[b][:nfg].tolist())) == nfg
    # the Philox priorities look uniform
    u = used[:, :2].ravel() / 2.0 ** 32
    assert abs(u.mean() - 0.5) < 0.02 and abs(u.std() - 12 ** -0.5) < 0.02


def test_visible_outputs_follow_output_iou(cuda):
    rng = np.random.default_rng(6)
    rois, gt = _scene(rng, 1, 100, 10, n_gt=[3], n_valid=[100])
    a = ops.ProposalTarget(_t(rois, cuda), _t(gt, cuda), 81, 1, 32, 0.5, 0.5, 0.0, False, seed=1)
    b = ops.ProposalTarget(_t(rois, cuda), _t(gt, cuda), 81, 1, 32, 0.5, 0.5, 0.0, False, seed=1, output_iou=True)
    assert len(a) == 4 and len(b) == 5 and b[4].shape == (1, 32)
    assert a[2].shape == (1, 32, 324)


@pytest.mark.parametrize("image_rois,filter_scales", [(128, True), (-1, False), (-1, True)])
def test_proposal_target_v2(cuda, image_rois, filter_scales):
    rng = np.random.default_rng(17)
    B, R, G = 2, 600, 30
    rois, gt = _scene(rng, B, R, G, n_gt=[9, 4], n_valid=[600, 350])
    vr = np.array([[0, 120], [90, 1e4]], np.float32)  # image 0 drops big gts, image 1 drops small ones
    pr = rng.integers(0, 2 ** 32, (B, 4, R + G), dtype=np.uint64).astype(np.uint32)
    ref = oracle.proposal_target(rois, gt, pr, 81, image_rois, 0.25, 0.5, 0.5, 0.0, valid_ranges=vr,
                                 filter_scales=filter_scales)
    res = ops.ProposalTarget_v2(_t(rois, cuda), _t(gt, cuda), _t(vr, cuda), 81, B, image_rois, 0.5, 0.5, 0.0, False,
                                filter_scales=filter_scales, priorities=_t(pr.astype(np.int64), cuda),
                                return_debug=True)
    IR = R if image_rois == -1 else image_rois
    assert res[0].shape == (B, IR, 4)
    _compare(res[:6], ref, IR)
