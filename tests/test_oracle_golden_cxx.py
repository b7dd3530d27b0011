"""The oracle against committed vectors produced by the reference's own operator_cxx sources (compiled unmodified,
oracle/build_ref_cxx.py; generator tests/golden/make_golden_cxx.py).  Runs anywhere: needs neither /root/reference
nor the compiled library.  Everything is compared bit for bit."""
import os

import numpy as np

import oracle
from test_oracle_ref_cxx import BASE, FOCAL_MODES, focal_case, oracle_under_constant_rand, sigmoid_ce_case

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
