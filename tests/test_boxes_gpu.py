"""DecodeBBox / Proposal_v3 / _contrib_NMS / batched greedy NMS: CUDA vs oracle through the C ABI.

Index outputs (which boxes are selected, in which order, which are suppressed) must be exact;
box coordinates are floats that pass through expf (CUDA's and glibc's differ by <= 2 ulp), so
they are compared with rtol 1e-5 (north_star tolerance: 1e-4 relative)."""
import numpy as np
import pytest
import torch

import oracle
from simpledet_b200 import ops

pytestmark = pytest.mark.gpu


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _boxes(rng, n, size=600):
    xy = rng.uniform(0, size, (n, 2))
    wh = rng.uniform(4, 200, (n, 2))
    return np.concatenate([xy, xy + wh], 1).astype(np.float32)


@pytest.mark.parametrize("agnostic", [True, False])
@pytest.mark.parametrize("kind", ["xywh", "xyxy"])
def test_decode_bbox(cuda, agnostic, kind):
    rng = np.random.default_rng(1)
    B, N, K = 2, 300, 81
    rois = np.stack([_boxes(rng, N), _boxes(rng, N)])
    deltas = (rng.standard_normal((B, N, 4 * K)) * 0.5).astype(np.float32)
    im_info = np.array([[800, 1333, 1.5], [600, 700, 1.0]], np.float32)
    mean, std = (0.0, 0.1, -0.05, 0.0), (0.1, 0.1, 0.2, 0.2)
    ref = oracle.decode_bbox(rois, deltas, im_info, mean, std, agnostic, kind)
    out = ops.DecodeBBox(_t(rois, cuda), _t(deltas, cuda), _t(im_info, cuda), mean, std, agnostic, kind)
    assert out.shape == ref.shape
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-5, atol=1e-4)
    if kind == "xyxy":  # no exp on this path: bit-exact
        assert np.array_equal(out.cpu().numpy(), ref)


def test_decode_bbox_defaults_are_class_agnostic(cuda):
    rois = torch.tensor([[[10.0, 10, 50, 60]]], device=cuda)
    d = torch.zeros((1, 1, 8), device=cuda)
    out = ops.DecodeBBox(rois, d, torch.tensor([[100.0, 100, 1]], device=cuda))
    assert out.shape == (1, 1, 4) and torch.equal(out, rois)


@pytest.mark.parametrize("n,thr,ge", [(1, 0.5, True), (63, 0.5, True), (64, 0.7, False), (65, 0.3, True),
                                       (1000, 0.7, True), (2000, 0.5, False), (6000, 0.7, True)])
def test_nms_sorted_exact(cuda, n, thr, ge):
    rng = np.random.default_rng(n)
    P = 3
    dets = np.zeros((P, n, 5), np.float32)
    for p in range(P):
        b = _boxes(rng, n, 300 + 200 * p)
        s = np.sort(rng.uniform(0, 1, n).astype(np.float32))[::-1]
        dets[p] = np.concatenate([b, s[:, None]], 1)
    keep, nkeep = ops.nms_sorted(_t(dets, cuda), thr, ge)
    keep, nkeep = keep.cpu().numpy(), nkeep.cpu().numpy()
    for p in range(P):
        # oracle: greedy over the given order with the same comparator
        if ge:
            ref = oracle.greedy_nms(dets[p], thr, order=np.arange(n))
        else:
            o, _ = oracle.contrib_nms(dets[p][None], n, n, thr, already_sorted=True)
            ref = None
            kept_boxes = o[0][: int(nkeep[p])]
            assert np.array_equal(kept_boxes, dets[p][keep[p, : nkeep[p]], :4])
            assert int(nkeep[p]) == n or np.all(o[0][int(nkeep[p]):] == 0)
        if ref is not None:
            assert int(nkeep[p]) == len(ref)
            assert np.array_equal(keep[p, : nkeep[p]], ref)


def test_nms_sorted_counts_and_identical_boxes(cuda):
    """All-identical boxes: IoU == 1 -> only the first survives; ragged counts per problem."""
    n = 200
    dets = np.tile(np.array([[10, 10, 50, 50, 0.5]], np.float32), (2, n, 1))
    dets[1, :, :4] += np.arange(n, dtype=np.float32)[:, None] * 100  # disjoint -> all kept
    counts = np.array([n, 77], np.int32)
    keep, nkeep = ops.nms_sorted(_t(dets, cuda), 0.5, True, counts=_t(counts, cuda))
    assert nkeep.cpu().tolist() == [1, 77]
    assert keep[1, :77].cpu().tolist() == list(range(77))


@pytest.mark.parametrize("is_train", [False, True])
@pytest.mark.parametrize("hw,stride,pre,post", [((50, 84), 16, 1000, 1000), ((25, 42), 32, 2000, 2000),
                                               ((100, 167), 8, 2000, 300),
                                               ((200, 336), 4, 1000, 1000)])  # chunked pre-selection
def test_proposal_v3(cuda, is_train, hw, stride, pre, post):
    rng = np.random.default_rng(pre + hw[0])
    B, A = 2, 3
    H, W = hw
    cls = rng.uniform(0, 1, (B, 2 * A, H, W)).astype(np.float32)
    # ties in the scores exercise the stable order (index ascending)
    cls[:, A:, : H // 3] = np.round(cls[:, A:, : H // 3], 2)
    deltas = (rng.standard_normal((B, 4 * A, H, W)) * 0.3).astype(np.float32)
    im_info = np.array([[H * stride - 7, W * stride - 11, 1.0], [H * stride * 0.8, W * stride * 0.9, 1.6]], np.float32)
    kw = dict(feature_stride=stride, scales=(8,), ratios=(0.5, 1, 2), rpn_pre_nms_top_n=pre,
              rpn_post_nms_top_n=post, threshold=0.7, rpn_min_size=16 if stride > 8 else 0, is_train=is_train)
    ro, rs, rdets, rkeep, rnk = oracle.proposal_v3(cls, deltas, im_info, debug=True, **kw)
    out, sc = ops.Proposal_v3(_t(cls, cuda), _t(deltas, cuda), _t(im_info, cuda), output_score=True, **kw)
    assert out.shape == ro.shape and sc.shape == rs.shape
    # scores identify the selected anchors exactly (they are copied, not computed)
    assert np.array_equal(sc.cpu().numpy(), rs)
    np.testing.assert_allclose(out.cpu().numpy(), ro, rtol=1e-5, atol=1e-3)


def test_proposal_v3_iou_loss(cuda):
    """IoUPredKernel path: additive decode and the padded-cell mask applied BEFORE the sort (both
    the single-CTA and the chunked selection)."""
    rng = np.random.default_rng(77)
    for (H, W, stride, pre) in ((38, 50, 16, 1500), (120, 160, 4, 1000)):
        B, A = 2, 3
        cls = rng.uniform(0, 1, (B, 2 * A, H, W)).astype(np.float32)
        deltas = (rng.standard_normal((B, 4 * A, H, W)) * 6).astype(np.float32)
        im_info = np.array([[H * stride - 70, W * stride - 90, 1.0], [H * stride, W * stride, 1.3]], np.float32)
        kw = dict(feature_stride=stride, scales=(8,), ratios=(0.5, 1, 2), rpn_pre_nms_top_n=pre,
                  rpn_post_nms_top_n=600, threshold=0.7, rpn_min_size=4, iou_loss=True)
        ro, rs = oracle.proposal_v3(cls, deltas, im_info, **kw)
        out, sc = ops.Proposal_v3(_t(cls, cuda), _t(deltas, cuda), _t(im_info, cuda), output_score=True, **kw)
        assert np.array_equal(sc.cpu().numpy(), rs)
        np.testing.assert_allclose(out.cpu().numpy(), ro, rtol=1e-6, atol=1e-4)


def test_proposal_v3_degenerate_harness(cuda):
    """detection_infer_speed.py: zero weights -> constant fg prob 0.5, zero deltas,
    im_info=(400, 666.5, 2): every score ties, order must be anchor-index order."""
    B, A, H, W = 1, 3, 50, 84
    cls = np.full((B, 2 * A, H, W), 0.5, np.float32)
    deltas = np.zeros((B, 4 * A, H, W), np.float32)
    im_info = np.array([[400, 666.5, 2.0]], np.float32)
    kw = dict(feature_stride=16, scales=(8,), ratios=(0.5, 1, 2), rpn_pre_nms_top_n=1000,
              rpn_post_nms_top_n=1000, threshold=0.7, rpn_min_size=0)
    ro, rs = oracle.proposal_v3(cls, deltas, im_info, **kw)
    out, sc = ops.Proposal_v3(_t(cls, cuda), _t(deltas, cuda), _t(im_info, cuda), output_score=True, **kw)
    assert np.array_equal(sc.cpu().numpy(), rs)
    assert np.array_equal(out.cpu().numpy(), ro)  # exp(0) = 1 exactly on both sides


@pytest.mark.parametrize("sorted_in", [False, True])
def test_contrib_nms(cuda, sorted_in):
    rng = np.random.default_rng(5)
    B, count = 2, 3000
    props = np.zeros((B, count, 5), np.float32)
    for b in range(B):
        s = rng.uniform(0, 1, count).astype(np.float32)
        s[::7] = 0.25  # ties
        if sorted_in:
            s = np.sort(s)[::-1]
        props[b] = np.concatenate([_boxes(rng, count), s[:, None]], 1)
    ro, rs = oracle.contrib_nms(props, 2000, 500, 0.6, already_sorted=sorted_in)
    out, sc = ops.NMS(_t(props, cuda), 2000, 500, 0.6, output_score=True, already_sorted=sorted_in)
    assert np.array_equal(out.cpu().numpy(), ro) and np.array_equal(sc.cpu().numpy(), rs)
    # post > pre: rows beyond min(post, pre) are untouched by the reference (NaN in the oracle)
    ro, rs = oracle.contrib_nms(props, 100, 300, 0.6, already_sorted=sorted_in)
    out, sc = ops.NMS(_t(props, cuda), 100, 300, 0.6, output_score=True, already_sorted=sorted_in)
    assert np.array_equal(out.cpu().numpy()[:, :100], ro[:, :100]) and np.isnan(ro[:, 100:]).all()


def test_get_top_proposal(cuda):
    from oracle import np_ops

    rng = np.random.default_rng(3)
    B, M = 2, 5000
    boxes = np.stack([_boxes(rng, M), _boxes(rng, M)])
    scores = rng.uniform(0, 1, (B, M, 1)).astype(np.float32)
    scores[:, ::5] = 0.5  # ties -> lower index first
    rb, rs = np_ops.get_top_proposal(boxes, scores, 1000)
    ob, os_ = ops.get_top_proposal(_t(boxes, cuda), _t(scores, cuda), 1000)
    assert np.array_equal(ob.cpu().numpy(), rb) and np.array_equal(os_.cpu().numpy(), rs)


def test_multiclass_nms_matches_do_nms(cuda):
    from oracle import np_ops

    rng = np.random.default_rng(4)
    B, N, K = 2, 300, 9
    score = rng.uniform(0, 1, (B, N, K)).astype(np.float32) ** 3  # distinct scores, many below 0.05
    bbox = np.stack([np.concatenate([_boxes(rng, N, 200) for _ in range(K)], 1) for _ in range(B)])
    dets, counts, keep, nkeep, src = ops.multiclass_nms(_t(score, cuda), _t(bbox, cuda), 0.5, 0.05)
    dets, counts, keep, nkeep, src = [x.cpu().numpy() for x in (dets, counts, keep, nkeep, src)]
    for b in range(B):
        ref = np_ops.do_nms(score[b], bbox[b], 0.5, 0.05)
        for cid in range(K):
            p = b * K + cid
            got = dets[p][keep[p, : nkeep[p]]]
            assert counts[p] == (score[b, :, cid] > 0.05).sum()
            assert np.array_equal(got, ref[cid]), (b, cid)
            # candidates trace back to their roi
            assert np.array_equal(score[b, src[p, : counts[p]], cid], dets[p, : counts[p], 4])


def test_proposal_v3_fpn_equals_per_level_concat(cuda):
    rng = np.random.default_rng(21)
    B, A = 2, 3
    strides = (4, 8, 16, 32, 64)
    shapes = [(-(-320 // s), -(-416 // s)) for s in strides]
    cls = [rng.uniform(0, 1, (B, 2 * A, h, w)).astype(np.float32) for h, w in shapes]
    dl = [(rng.standard_normal((B, 4 * A, h, w)) * 0.3).astype(np.float32) for h, w in shapes]
    im_info = np.array([[320, 416, 1.0], [300, 400, 1.3]], np.float32)
    kw = dict(scales=(8,), ratios=(0.5, 1, 2), rpn_pre_nms_top_n=500, rpn_post_nms_top_n=200, threshold=0.7,
              rpn_min_size=0)
    ref = [oracle.proposal_v3(c, d, im_info, feature_stride=s, **kw) for c, d, s in zip(cls, dl, strides)]
    rb = np.concatenate([r[0] for r in ref], 1)
    rs = np.concatenate([r[1] for r in ref], 1)
    out, sc = ops.Proposal_v3_fpn([_t(c, cuda) for c in cls], [_t(d, cuda) for d in dl], _t(im_info, cuda),
                                  strides, **kw)
    assert np.array_equal(sc.cpu().numpy(), rs)
    np.testing.assert_allclose(out.cpu().numpy(), rb, rtol=1e-5, atol=1e-3)


def test_proposal_v3_fpn_training_with_a_small_level(cuda):
    """FPN training: P6 has fewer anchors than rpn_post_nms_top_n (819 < 2000 at 800x1333).  Every level keeps `post`
    rows; a level writes min(post, its anchors) rows (kept boxes, then wrap-around padding like proposal_v3.cu:368-390)
    and the rest are zero rows, which ProposalTarget ignores."""
    rng = np.random.default_rng(22)
    B, A = 2, 3
    strides = (8, 16, 32, 64)
    shapes = [(-(-256 // s), -(-384 // s)) for s in strides]          # 32x48 ... 4x6 (72 anchors)
    cls = [rng.uniform(0, 1, (B, 2 * A, h, w)).astype(np.float32) for h, w in shapes]
    dl = [(rng.standard_normal((B, 4 * A, h, w)) * 0.3).astype(np.float32) for h, w in shapes]
    im_info = np.array([[256, 384, 1.0], [250, 380, 1.3]], np.float32)
    post = 150
    kw = dict(scales=(8,), ratios=(0.5, 1, 2), rpn_pre_nms_top_n=400, rpn_post_nms_top_n=post, threshold=0.7,
              rpn_min_size=0, is_train=True)
    out, sc = ops.Proposal_v3_fpn([_t(c, cuda) for c in cls], [_t(d, cuda) for d in dl], _t(im_info, cuda), strides, **kw)
    out, sc = out.cpu().numpy(), sc.cpu().numpy()
    assert out.shape == (B, len(strides) * post, 4)
    for l, (c, d, s) in enumerate(zip(cls, dl, strides)):
        rb, rs = oracle.proposal_v3(c, d, im_info, feature_stride=s, **kw)   # (B, min(post, anchors), .)
        n = rb.shape[1]
        assert n == min(post, A * c.shape[2] * c.shape[3])
        blk, sblk = out[:, l * post:(l + 1) * post], sc[:, l * post:(l + 1) * post]
        assert np.array_equal(sblk[:, :n], rs), l
        np.testing.assert_allclose(blk[:, :n], rb, rtol=1e-5, atol=1e-3)
        assert not blk[:, n:].any() and not sblk[:, n:].any()
    assert (shapes[-1][0] * shapes[-1][1] * A) < post                  # the case is exercised


@pytest.mark.parametrize("version,is_train,filt", [(1, False, False), (1, True, False), (2, False, True)])
def test_proposal_legacy(cuda, version, is_train, filt):
    """_contrib_Proposal / _contrib_Proposal_v2: filter and padded-cell mask BEFORE the sort."""
    rng = np.random.default_rng(40 + version)
    B, A, H, W, stride = 2, 12, 38, 50, 16
    cls = rng.uniform(0, 1, (B, 2 * A, H, W)).astype(np.float32)
    deltas = (rng.standard_normal((B, 4 * A, H, W)) * 0.4).astype(np.float32)
    deltas[:, 2::4] -= 1.0  # shrink many boxes below rpn_min_size
    im_info = np.array([[H * stride - 40, W * stride - 64, 1.0], [H * stride, W * stride, 1.5]], np.float32)
    vr = np.array([[0, 200], [60, 1e5]], np.float32)
    kw = dict(feature_stride=stride, scales=(2, 4, 8, 16), ratios=(0.5, 1, 2), rpn_pre_nms_top_n=3000,
              rpn_post_nms_top_n=400, threshold=0.7, rpn_min_size=16)
    ro, rs = oracle.proposal_legacy(cls, deltas, im_info, version=version, valid_ranges=vr, is_train=is_train,
                                    filter_scales=filt, **kw)
    if version == 1:
        out, sc = ops.Proposal(_t(cls, cuda), _t(deltas, cuda), _t(im_info, cuda), output_score=True,
                               is_train=is_train, **kw)
    else:
        out, sc = ops.Proposal_v2(_t(cls, cuda), _t(deltas, cuda), _t(im_info, cuda), _t(vr, cuda),
                                  output_score=True, filter_scales=filt, **kw)
    assert out.shape == ro.shape
    assert np.array_equal(sc.cpu().numpy(), rs)
    np.testing.assert_allclose(out.cpu().numpy(), ro, rtol=1e-5, atol=1e-3)


def test_gen_anchor(cuda):
    for (H, W, stride) in [(100, 168, 8), (7, 11, 128), (1, 1, 64)]:
        scales = tuple(4 * 2 ** (i / 3) for i in range(3))  # retinanet: 3 octave scales (non-integer doubles)
        ref = oracle.gen_anchor(H, W, stride, scales, (0.5, 1, 2))
        got = ops.GenAnchor(torch.empty((1, 9, H, W), device=cuda), scales=scales, ratios=(0.5, 1, 2),
                            feature_stride=stride)
        assert np.array_equal(got.cpu().numpy(), ref)


def test_gen_proposal(cuda):
    rng = np.random.default_rng(51)
    B, A, H, W, stride = 2, 9, 25, 34, 16
    cls = rng.uniform(0, 1, (B, 2 * A, H, W)).astype(np.float32)
    deltas = (rng.standard_normal((B, 4 * A, H, W)) * 0.4).astype(np.float32)
    deltas[:, 2::4] -= 0.8
    im_info = np.array([[H * stride - 30, W * stride - 50, 1.0], [H * stride, W * stride, 2.0]], np.float32)
    anchors = oracle.gen_anchor(H, W, stride, (2, 4, 8), (0.5, 1, 2))
    for pre in (2000, A * H * W + 50):
        ref = oracle.gen_proposal(cls, deltas, im_info, anchors, feature_stride=stride, rpn_pre_nms_top_n=pre,
                                  rpn_min_size=8)
        got = ops.GenProposal(_t(cls, cuda), _t(deltas, cuda), _t(im_info, cuda), _t(anchors, cuda),
                              feature_stride=stride, rpn_pre_nms_top_n=pre, rpn_min_size=8).cpu().numpy()
        assert np.array_equal(got[..., 4], ref[..., 4])
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("thresh,K", [(0.05, 80), (0.0, 80), (0.3, 1)])
def test_gen_proposal_retina(cuda, thresh, K):
    rng = np.random.default_rng(52)
    B, A, H, W, stride = 2, 9, 13, 21, 32
    cls = (rng.uniform(0, 1, (B, A * K, H, W)) ** 4).astype(np.float32)  # mostly below thresh, like sigmoid outputs
    cls[0, :4] = cls[0, 4:8]  # exact score ties -> order by reference index
    deltas = (rng.standard_normal((B, 4 * A, H, W)) * 0.5).astype(np.float32)
    im_info = np.array([[H * stride - 20, W * stride - 40, 1.0], [H * stride, W * stride, 1.6]], np.float32)
    scales = tuple(4 * 2 ** (i / 3) for i in range(3))
    anchors = oracle.gen_anchor(H, W, stride, scales, (0.5, 1, 2))
    kw = dict(num_anchors=A, rpn_pre_nms_top_n=1000, rpn_min_size=40, thresh=thresh,
              anchor_mean=(0.0, 0.1, 0.0, -0.1), anchor_std=(0.1, 0.1, 0.2, 0.2))
    rb, rs = oracle.gen_proposal_retina(cls, deltas, im_info, anchors, **kw)
    gb, gs = ops.GenProposalRetina(_t(cls, cuda), _t(deltas, cuda), _t(im_info, cuda), _t(anchors, cuda),
                                   feature_stride=stride, **kw)
    assert gs.shape == rs.shape == (B, 1000, K + 1)
    assert np.array_equal(gs.cpu().numpy(), rs)
    np.testing.assert_allclose(gb.cpu().numpy(), rb, rtol=1e-5, atol=1e-3)
    assert (rs > 0).any()


def test_gen_proposal_retina_few_survivors(cuda):
    """Fewer survivors than rpn_pre_nms_top_n (and none at all): remaining rows are all-zero."""
    rng = np.random.default_rng(53)
    B, A, K, H, W = 2, 9, 80, 4, 6
    cls = rng.uniform(0, 0.04, (B, A * K, H, W)).astype(np.float32)
    cls[0, 7, 1, 2] = 0.9
    cls[0, 85, 3, 5] = 0.5
    deltas = np.zeros((B, 4 * A, H, W), np.float32)
    im_info = np.array([[512, 768, 1.0]] * 2, np.float32)
    anchors = oracle.gen_anchor(H, W, 128, (4, 5, 6), (0.5, 1, 2))
    kw = dict(num_anchors=A, rpn_pre_nms_top_n=1000, rpn_min_size=0, thresh=0.05)
    rb, rs = oracle.gen_proposal_retina(cls, deltas, im_info, anchors, **kw)
    gb, gs = ops.GenProposalRetina(_t(cls, cuda), _t(deltas, cuda), _t(im_info, cuda), _t(anchors, cuda), **kw)
    assert np.array_equal(gs.cpu().numpy(), rs) and np.array_equal(gb.cpu().numpy(), rb)
    assert np.count_nonzero(rs) == 2


def test_gen_proposal_retina_dense_and_tied_scores(cuda):
    """The histogram pre-filter's corner cases: (a) nearly every pair passes the threshold and the scores sit in a few
    exact values, so the cut bin holds far more than `pre` keys and the order is decided by the reference index;
    (b) scores far above the binned range (> 1) share the top bin; (c) thresh = 0."""
    rng = np.random.default_rng(54)
    B, A, K, H, W, stride = 2, 9, 8, 11, 17, 16
    cls = rng.choice(np.array([0.25, 0.5, 0.75], np.float32), (B, A * K, H, W))
    cls[1] = (rng.uniform(0, 1, (A * K, H, W)) ** 2).astype(np.float32)
    cls[1, :3] *= 1000.0
    deltas = (rng.standard_normal((B, 4 * A, H, W)) * 0.3).astype(np.float32)
    im_info = np.array([[H * stride, W * stride, 1.0]] * B, np.float32)
    anchors = oracle.gen_anchor(H, W, stride, (4, 5, 6), (0.5, 1, 2))
    for thresh, pre in ((0.05, 300), (0.0, 1000), (0.6, 50)):
        kw = dict(num_anchors=A, rpn_pre_nms_top_n=pre, rpn_min_size=0, thresh=thresh)
        rb, rs = oracle.gen_proposal_retina(cls, deltas, im_info, anchors, **kw)
        gb, gs = ops.GenProposalRetina(_t(cls, cuda), _t(deltas, cuda), _t(im_info, cuda), _t(anchors, cuda),
                                       feature_stride=stride, **kw)
        assert np.array_equal(gs.cpu().numpy(), rs), (thresh, pre)
        np.testing.assert_allclose(gb.cpu().numpy(), rb, rtol=1e-5, atol=1e-3)


def test_final_detections(cuda):
    """detection_test.py:233-291: per-class NMS then sorted(result, key=score)[-max_det:] per image,
    including equal scores across classes (the later class survives the cut)."""
    from oracle import np_ops
    rng = np.random.default_rng(61)
    B, N, K = 2, 300, 12
    score = rng.uniform(0, 1, (B, N, K)).astype(np.float32) ** 3
    score[:, :, 7] = score[:, :, 2]          # ties between classes 2 and 7
    score[1, :, 4] = 0.01                    # a class without detections
    xy = rng.uniform(0, 400, (B, N, 1, 2))
    wh = rng.uniform(10, 200, (B, N, K, 2))
    bbox = np.concatenate([np.broadcast_to(xy, (B, N, K, 2)) + rng.uniform(-5, 5, (B, N, K, 2)), xy + wh], 3)
    bbox = bbox.reshape(B, N, K * 4).astype(np.float32)
    for max_det in (100, 7, 5000):
        out, cnt = ops.final_detections(_t(score, cuda), _t(bbox, cuda), 0.5, 0.05, max_det)
        out, cnt = out.cpu().numpy(), cnt.cpu().numpy()
        for b in range(B):
            ref = np_ops.final_detections(score[b], bbox[b], 0.5, 0.05, max_det)
            assert cnt[b] == ref.shape[0]
            assert np.array_equal(out[b, :cnt[b]], ref), (max_det, b)
            assert np.all(out[b, cnt[b]:, 5] == -1)
