#!/usr/bin/env python
"""Generates tests/golden/reference_cxx_ops.npz by RUNNING THE REFERENCE'S OWN operator_cxx SOURCES, compiled
unmodified by oracle/build_ref_cxx.py (oracle/_ref/libref_cxx.so).  Only input/output vectors are committed.
Run:  python tests/golden/make_golden_cxx.py     (needs /root/reference or the prebuilt library)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_cxx  # noqa: E402
from simpledet_b200 import synth  # noqa: E402
import test_oracle_ref_cxx as T  # noqa: E402


def main():
    g = {}
    rng = np.random.default_rng(0)
    # RoIAlign_v2: config-1-like (fewer channels) + edge rois, forward and GPU-functor backward
    data, rois, pooled, scale = synth.config1(0)
    data = np.ascontiguousarray(data[:, :6])
    kw = dict(pooled_size=pooled, spatial_scale=scale)
    out, ax, ay = ref_cxx.forward("_contrib_ROIAlign_v2", kw, [data, rois])
    og = rng.standard_normal(out.shape).astype(np.float32)
    gd, _ = ref_cxx.forward("_backward_ROIAlign_v2", kw, [og, rois, ax, ay], out_shapes=[data.shape, rois.shape], dev="gpu")
    g.update(ra_data=data, ra_rois=rois, ra_out=out, ra_ax=ax, ra_ay=ay, ra_ograd=og, ra_gdata=gd)
    edata = rng.standard_normal((1, 3, 25, 42)).astype(np.float32)
    for p in (7, 14):
        o, x, y = ref_cxx.forward("_contrib_ROIAlign_v2", dict(pooled_size=(p, p), spatial_scale=1 / 32), [edata, T.EDGE_ROIS])
        g.update({f"ra_edge{p}_out": o, f"ra_edge{p}_ax": x, f"ra_edge{p}_ay": y})
    g.update(ra_edge_data=edata, ra_edge_rois=T.EDGE_ROIS)
    # ROIPooling_v1
    pdata = rng.standard_normal((2, 3, 30, 40)).astype(np.float32)
    r = synth.random_rois(rng, 1, 40, 480, 640)[0]
    prois = np.concatenate([rng.integers(0, 2, (40, 1)).astype(np.float32), r], 1)
    po, pi = ref_cxx.forward("ROIPooling_v1", dict(pooled_size=(7, 7), spatial_scale=1 / 16), [pdata, prois])
    g.update(rp_data=pdata, rp_rois=prois, rp_out=po, rp_idx=pi)
    # DecodeBBox
    drois = synth.random_rois(rng, 2, 100)
    ddel = (rng.standard_normal((2, 100, 12)) * np.array([1, 1, 2.5, 2.5] * 3)).astype(np.float32)
    dinfo = np.array([[800, 1333, 1.0], [600, 901, 1.5]], np.float32)
    for ag in (True, False):
        for ty in ("xywh", "xyxy"):
            (o,) = ref_cxx.forward("_contrib_DecodeBBox", dict(class_agnostic=ag, bbox_decode_type=ty), [drois, ddel, dinfo])
            g[f"db_{int(ag)}_{ty}"] = o
    g.update(db_rois=drois, db_deltas=ddel, db_info=dinfo)
    # GenAnchor
    (a,) = ref_cxx.forward("_contrib_GenAnchor", dict(feature_stride=8, scales=(4.0, 5.04, 6.35), ratios=(0.5, 1.0, 2.0)),
                           [np.zeros((1, 18, 10, 17), np.float32)])
    g["ga_out"] = a
    # ProposalTarget with rand() == 0 (every shuffle rotates right by one); the matching priorities are rebuilt by the test
    trois, tgt = T._pt_inputs(np.random.default_rng(10), 2, 120, 10, 4, 40)
    ref_cxx.set_rand_const(0)
    outs = ref_cxx.forward("ProposalTarget", dict(T.BASE, batch_images=2, image_rois=32), [trois, tgt])
    g.update(pt_rois=trois, pt_gt=tgt, **{f"pt_out{i}": o for i, o in enumerate(outs)})
    # FocalLoss (three parameter sets) and BBoxNorm: forward + the operators' own Backward
    for m, kw in enumerate(T.FOCAL_MODES):
        data, label, ograd = T.focal_case(30 + m)
        (out,) = ref_cxx.forward("_contrib_FocalLoss", dict(kw, workspace=8), [data, label])
        gd, _ = ref_cxx.backward("_contrib_FocalLoss", dict(kw, workspace=8), [ograd], [data, label], [out])
        g.update({f"fl{m}_out": out, f"fl{m}_gdata": gd})
    brng = np.random.default_rng(41)
    bdata = brng.standard_normal((2, 12, 35)).astype(np.float32)
    blabel = brng.integers(-1, 3, (2, 3 * 35)).astype(np.float32)
    bgout = brng.standard_normal((2, 12, 35)).astype(np.float32)
    bg, _ = ref_cxx.backward("_contrib_BBoxNorm", {}, [bgout], [bdata, blabel], [bdata])
    g.update(bn_label=blabel, bn_gout=bgout, bn_gdata=bg)
    # SigmoidCrossEntropy (GPU path on the host)
    sdata, slabel = T.sigmoid_ce_case()
    souts = ref_cxx.forward("_contrib_SigmoidCrossEntropy", dict(grad_scale=0.37), [sdata, slabel], dev="gpu")
    (sg, _) = ref_cxx.backward("_contrib_SigmoidCrossEntropy", dict(grad_scale=0.37), [np.ones_like(souts[0])],
                               [sdata, slabel], souts, dev="gpu")
    g.update(sce_out=souts[0], sce_gdata=sg)
    # Proposal_v3: the GPU operator run serially on the host (inference and training padding, box-delta and IoU decode)
    for ci, c in enumerate(T.PROPOSAL_V3_CASES):
        c = dict(c)
        kw = c.pop("kw")
        cls, reg, info = T.rpn_case(**c)
        for tr in (False, True):
            for iou in (False, True):
                o, s_ = ref_cxx.forward("_contrib_Proposal_v3", dict(kw, is_train=tr, iou_loss=iou, output_score=True, workspace=64),
                                        [cls, reg, info], dev="gpu")
                n = min(kw["rpn_post_nms_top_n"], kw["rpn_pre_nms_top_n"], cls.shape[1] // 2 * cls.shape[2] * cls.shape[3]) if tr \
                    else kw["rpn_post_nms_top_n"]
                B = o.shape[0]                      # training with count < post: the written rows are the flat prefix
                g[f"p3_{ci}_{int(tr)}_{int(iou)}_out"] = o.reshape(-1, 4)[:B * n].reshape(B, n, 4)
                g[f"p3_{ci}_{int(tr)}_{int(iou)}_score"] = s_.reshape(-1)[:B * n].reshape(B, n, 1)
    # Proposal / Proposal_v2 / GenProposal on the same cases, _contrib_NMS, GenProposalRetina (GPU operators on the host)
    for ci, c in enumerate(T.PROPOSAL_V3_CASES):
        c = dict(c)
        kw = c.pop("kw")
        cls, reg, info = T.rpn_case(**c)
        B, A2, H, W = cls.shape
        count = A2 // 2 * H * W
        pre = min(kw["rpn_pre_nms_top_n"], count)
        vr = np.array([[0, 64], [32, 1e5]], np.float32)[:B]
        anchors = ref_cxx.forward("_contrib_GenAnchor", dict(feature_stride=kw["feature_stride"], scales=kw["scales"], ratios=kw["ratios"]),
                                  [np.zeros((1, A2, H, W), np.float32)])[0].reshape(-1, 4)
        for iou in (False, True):
            for tr in (False, True):
                o, s_ = ref_cxx.forward("_contrib_Proposal", dict(kw, is_train=tr, iou_loss=iou, output_score=True, workspace=64),
                                        [cls, reg, info], dev="gpu")
                n = min(kw["rpn_post_nms_top_n"], pre) if tr else kw["rpn_post_nms_top_n"]
                g[f"p1_{ci}_{int(tr)}_{int(iou)}_out"], g[f"p1_{ci}_{int(tr)}_{int(iou)}_score"] = T.written(o, n, 4), T.written(s_, n, 1)
            for filt in (False, True):
                o, s_ = ref_cxx.forward("_contrib_Proposal_v2", dict(kw, filter_scales=filt, iou_loss=iou, output_score=True, workspace=64),
                                        [cls, reg, info, vr], dev="gpu")
                n = min(kw["rpn_post_nms_top_n"], pre)
                g[f"p2_{ci}_{int(filt)}_{int(iou)}_out"], g[f"p2_{ci}_{int(filt)}_{int(iou)}_score"] = T.written(o, n, 4), T.written(s_, n, 1)
            k = dict(feature_stride=kw["feature_stride"], rpn_pre_nms_top_n=150, rpn_min_size=kw["rpn_min_size"], iou_loss=iou)
            (o,) = ref_cxx.forward("_contrib_GenProposal", dict(k, workspace=64), [cls, reg, info, anchors], dev="gpu")
            g[f"gp_{ci}_{int(iou)}"] = o[:, :min(150, count)]
    ndata = T.nms_case()
    for pre, post in T.NMS_CASES:
        kw = dict(rpn_pre_nms_top_n=pre, rpn_post_nms_top_n=post, threshold=0.6)
        o, s_ = ref_cxx.forward("_contrib_NMS", dict(kw, output_score=True, workspace=64), [ndata], dev="gpu")
        n = min(post, pre, ndata.shape[1])
        g[f"nms_{pre}_{post}_out"], g[f"nms_{pre}_{post}_score"] = T.written(o, n, 4), T.written(s_, n, 1)
    for K, thresh, pre, one_hot in T.RETINA_CASES:
        cls, reg, info, anchors = T.retina_case(K)
        kw = dict(num_anchors=9, rpn_pre_nms_top_n=pre, rpn_min_size=40, thresh=thresh, anchor_mean=(0.0, 0.1, 0.0, -0.1),
                  anchor_std=(0.1, 0.1, 0.2, 0.2), output_one_hot=one_hot)
        o, s_ = ref_cxx.forward("_contrib_GenProposalRetina", dict(kw, feature_stride=32, workspace=256),
                                [cls, reg, info, anchors], dev="gpu")
        # one-hot score rows are sparse; the npz is compressed
        g[f"gr_{K}_{pre}_box"], g[f"gr_{K}_{pre}_score"] = o, s_
    # ProposalMaskTarget with output_iou + output_ratio (rand() == 0), against the stand-in maskApi.h
    for M, few_fg in T.MASK_RATIO_CASES:
        rois, gtb, polys = T.mask_ratio_case(M, few_fg)
        ref_cxx.set_rand_const(0)
        outs = ref_cxx.forward("ProposalMaskTarget", dict(T.MASK_KW, num_args=3, batch_images=rois.shape[0], mask_size=M,
                                                          output_iou=True, output_ratio=True), [rois, gtb, polys])
        for i in (0, 1, 4):
            g[f"pm_{M}_out{i}"] = outs[i]
        g[f"pm_{M}_mask"] = outs[5].astype(np.int8)
        g[f"pm_{M}_ratio"] = outs[6]
    path = os.path.join(HERE, "reference_cxx_ops.npz")
    np.savez_compressed(path, **g)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
