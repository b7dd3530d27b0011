"""The oracle against committed vectors produced by the reference's own operator_cxx sources (compiled unmodified,
oracle/build_ref_cxx.py; generator tests/golden/make_golden_cxx.py).  Runs anywhere: needs neither /root/reference
nor the compiled library.  Everything is compared bit for bit."""
import os

import numpy as np

import oracle
from test_oracle_ref_cxx import (BASE, FOCAL_MODES, MASK_KW, MASK_RATIO_CASES, NMS_CASES, mask_ratio_case, PROPOSAL_V3_CASES, RETINA_CASES, focal_case, nms_case,
                                 oracle_under_constant_rand, retina_case, rpn_case, sigmoid_ce_case)

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_cxx_ops.npz"))


def test_roi_align_v2_forward_backward():
    o, x, y = oracle.roi_align_v2_forward(G["ra_data"], G["ra_rois"], (7, 7), 1 / 16)
    assert np.array_equal(o, G["ra_out"]) and np.array_equal(x, G["ra_ax"]) and np.array_equal(y, G["ra_ay"])
    gd = oracle.roi_align_v2_backward(G["ra_ograd"], G["ra_ax"], G["ra_ay"], G["ra_data"].shape)
    assert np.array_equal(gd, G["ra_gdata"])
    for p in (7, 14):
        o, x, y = oracle.roi_align_v2_forward(G["ra_edge_data"], G["ra_edge_rois"], (p, p), 1 / 32)
        assert np.array_equal(o, G[f"ra_edge{p}_out"]) and np.array_equal(x, G[f"ra_edge{p}_ax"])
        assert np.array_equal(y, G[f"ra_edge{p}_ay"])


def test_roi_pooling_v1():
    o, i = oracle.roi_pool_v1_forward(G["rp_data"], G["rp_rois"], (7, 7), 1 / 16)
    assert np.array_equal(o, G["rp_out"]) and np.array_equal(i, G["rp_idx"])


def test_decode_bbox():
    for ag in (True, False):
        for ty in ("xywh", "xyxy"):
            o = oracle.decode_bbox(G["db_rois"], G["db_deltas"], G["db_info"], class_agnostic=ag, bbox_decode_type=ty)
            assert np.array_equal(o, G[f"db_{int(ag)}_{ty}"])


def test_gen_anchor():
    assert np.array_equal(oracle.gen_anchor(10, 17, 8, (4.0, 5.04, 6.35), (0.5, 1.0, 2.0)), G["ga_out"].reshape(-1, 4))


def test_proposal_target():
    o = oracle_under_constant_rand(G["pt_rois"], G["pt_gt"], dict(BASE, image_rois=32), 0)
    for i in range(5):
        assert np.array_equal(o[i], G[f"pt_out{i}"]), i


def test_focal_loss_and_bbox_norm():
    for m, kw in enumerate(FOCAL_MODES):
        data, label, ograd = focal_case(30 + m)
        out = oracle.sigmoid(data)
        assert np.array_equal(out, G[f"fl{m}_out"])
        gd = oracle.focal_loss_backward(out, label, kw["alpha"], kw["gamma"], kw["grad_scale"], kw["normalization"],
                                        ograd if kw["out_grad"] else None)
        assert np.array_equal(gd, G[f"fl{m}_gdata"]), m
    assert np.array_equal(oracle.bbox_norm_backward(G["bn_gout"], G["bn_label"]), G["bn_gdata"])


def test_sigmoid_cross_entropy():
    data, label = sigmoid_ce_case()
    assert np.array_equal(oracle.sigmoid_ce_forward(data, label), G["sce_out"])
    assert np.array_equal(oracle.sigmoid_ce_backward(data, label, 0.37), G["sce_gdata"])


def test_proposal_v3():
    for ci, c in enumerate(PROPOSAL_V3_CASES):
        c = dict(c)
        kw = c.pop("kw")
        cls, reg, info = rpn_case(**c)
        for tr in (False, True):
            for iou in (False, True):
                r, sc = oracle.proposal_v3(cls, reg, info, is_train=tr, iou_loss=iou, **kw)
                assert np.array_equal(r, G[f"p3_{ci}_{int(tr)}_{int(iou)}_out"]), (ci, tr, iou)
                assert np.array_equal(sc, G[f"p3_{ci}_{int(tr)}_{int(iou)}_score"]), (ci, tr, iou)


def test_proposal_v1_v2_gen_proposal():
    for ci, c in enumerate(PROPOSAL_V3_CASES):
        c = dict(c)
        kw = c.pop("kw")
        cls, reg, info = rpn_case(**c)
        B, A2, H, W = cls.shape
        vr = np.array([[0, 64], [32, 1e5]], np.float32)[:B]
        anchors = oracle.gen_anchor(H, W, kw["feature_stride"], kw["scales"], kw["ratios"])
        for iou in (False, True):
            for tr in (False, True):
                r, sc = oracle.proposal_legacy(cls, reg, info, version=1, is_train=tr, iou_loss=iou, **kw)
                assert np.array_equal(r, G[f"p1_{ci}_{int(tr)}_{int(iou)}_out"]), (ci, tr, iou)
                assert np.array_equal(sc, G[f"p1_{ci}_{int(tr)}_{int(iou)}_score"])
            for filt in (False, True):
                r, sc = oracle.proposal_legacy(cls, reg, info, version=2, valid_ranges=vr, filter_scales=filt, iou_loss=iou, **kw)
                assert np.array_equal(r, G[f"p2_{ci}_{int(filt)}_{int(iou)}_out"]), (ci, filt, iou)
                assert np.array_equal(sc, G[f"p2_{ci}_{int(filt)}_{int(iou)}_score"])
            gp = oracle.gen_proposal(cls, reg, info, anchors, feature_stride=kw["feature_stride"], rpn_pre_nms_top_n=150,
                                     rpn_min_size=kw["rpn_min_size"], iou_loss=iou)
            n = G[f"gp_{ci}_{int(iou)}"].shape[1]
            assert np.array_equal(gp[:, :n], G[f"gp_{ci}_{int(iou)}"])


def test_contrib_nms_and_gen_proposal_retina():
    data = nms_case()
    for pre, post in NMS_CASES:
        r, sc = oracle.contrib_nms(data, rpn_pre_nms_top_n=pre, rpn_post_nms_top_n=post, threshold=0.6)
        n = G[f"nms_{pre}_{post}_out"].shape[1]
        assert np.array_equal(r[:, :n], G[f"nms_{pre}_{post}_out"]) and np.array_equal(sc[:, :n], G[f"nms_{pre}_{post}_score"])
    for K, thresh, pre, one_hot in RETINA_CASES:
        cls, reg, info, anchors = retina_case(K)
        rb, rs = oracle.gen_proposal_retina(cls, reg, info, anchors, num_anchors=9, rpn_pre_nms_top_n=pre, rpn_min_size=40,
                                            thresh=thresh, anchor_mean=(0.0, 0.1, 0.0, -0.1), anchor_std=(0.1, 0.1, 0.2, 0.2),
                                            output_one_hot=one_hot)
        assert np.array_equal(rb, G[f"gr_{K}_{pre}_box"]) and np.array_equal(rs, G[f"gr_{K}_{pre}_score"]), K


def test_proposal_mask_target_with_ratio():
    for M, few_fg in MASK_RATIO_CASES:
        rois, gt, polys = mask_ratio_case(M, few_fg)
        o = oracle_under_constant_rand(rois, gt, MASK_KW, 0, polys=polys, mask_size=M, output_ratio=True)
        for i in (0, 1, 4):
            assert np.array_equal(o[i], G[f"pm_{M}_out{i}"]), (M, i)
        assert np.array_equal(o[5], G[f"pm_{M}_mask"].astype(np.float32)) and np.array_equal(o[6], G[f"pm_{M}_ratio"])
