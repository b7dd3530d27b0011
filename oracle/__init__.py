"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's hot-path operators (C in ``oracle/*.c``, numpy in
``oracle/np_ops.py``).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package; ``simpledet_b200`` never
does.  See the header of each source file for the reference file:line it follows and for its
parity-pin status.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/*.c -> oracle/_build/liboracle.so (gcc, -ffp-contract=off)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith(".c")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "_build/liboracle.so"], check=True,
                       capture_output=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def roi_align_v2_forward(data, rois, pooled_size, spatial_scale, with_argmax=True):
    """roi_align_v2-inl.h:61-153.  data (B,C,H,W), rois (B,N,4) -> out[, argmax_x, argmax_y]."""
    data, rois = _f32(data), _f32(rois)
    B, C, H, W = data.shape
    assert rois.shape[0] == B and rois.shape[2] == 4
    N = rois.shape[1]
    ph, pw = pooled_size
    out = np.empty((B, N, C, ph, pw), np.float32)
    ax = np.empty_like(out) if with_argmax else None
    ay = np.empty_like(out) if with_argmax else None
    lib().oracle_roi_align_v2_forward(_p(data), _p(rois), B, N, C, H, W, ph, pw,
                                      ctypes.c_float(spatial_scale), _p(out), _p(ax), _p(ay))
    return (out, ax, ay) if with_argmax else out


def roi_align_v2_backward(ograd, argmax_x, argmax_y, data_shape, accumulate_into=None):
    """roi_align_v2.cu:35-84 (GPU semantics), adds applied in index order."""
    ograd, argmax_x, argmax_y = _f32(ograd), _f32(argmax_x), _f32(argmax_y)
    B, N, C, ph, pw = ograd.shape
    _, _, H, W = data_shape
    if accumulate_into is None:
        grad = np.empty(data_shape, np.float32)
        acc = 0
    else:
        grad = _f32(accumulate_into).copy()
        acc = 1
    lib().oracle_roi_align_v2_backward(_p(ograd), _p(argmax_x), _p(argmax_y), B, N, C, H, W, ph, pw,
                                       acc, _p(grad))
    return grad


def roi_pool_v1_forward(data, rois, pooled_size, spatial_scale):
    """roi_pooling_v1.cu:49-113.  rois (R,5) -> out, max_idx (R,C,PH,PW)."""
    data, rois = _f32(data), _f32(rois)
    B, C, H, W = data.shape
    R = rois.shape[0]
    ph, pw = pooled_size
    out = np.empty((R, C, ph, pw), np.float32)
    idx = np.empty_like(out)
    lib().oracle_roi_pool_v1_forward(_p(data), _p(rois), R, C, H, W, ph, pw,
                                     ctypes.c_float(spatial_scale), _p(out), _p(idx))
    return out, idx


def roi_pool_v1_backward(ograd, max_idx, rois, data_shape, accumulate_into=None):
    """roi_pooling_v1.cu:116-152."""
    ograd, max_idx, rois = _f32(ograd), _f32(max_idx), _f32(rois)
    R, C, ph, pw = ograd.shape
    B, _, H, W = data_shape
    if accumulate_into is None:
        grad = np.empty(data_shape, np.float32)
        acc = 0
    else:
        grad = _f32(accumulate_into).copy()
        acc = 1
    lib().oracle_roi_pool_v1_backward(_p(ograd), _p(max_idx), _p(rois), R, B, C, H, W, ph, pw, acc,
                                      _p(grad))
    return grad


def fpn_assign_levels(rois, strides, roi_canonical_scale=224, roi_canonical_level=4):
    """assign_layer_fpn.py:17-40 -> per-roi index into `strides` (-1: no level matches)."""
    rois = _f32(rois).reshape(-1, 4)
    lv = np.empty(rois.shape[0], np.int32)
    k_min, k_max = float(np.log2(min(strides))), float(np.log2(max(strides)))
    lib().oracle_fpn_assign_levels(_p(rois), ctypes.c_long(rois.shape[0]),
                                   ctypes.c_float(roi_canonical_scale),
                                   ctypes.c_float(roi_canonical_level), ctypes.c_float(k_min),
                                   ctypes.c_float(k_max), _p(lv))
    idx = np.full_like(lv, -1)
    for i, s in enumerate(strides):
        idx[(lv >= 0) & ((2 ** np.maximum(lv, 0)) == s) & (lv >= 0)] = i
    return idx


def fpn_roi_align_v2_forward(feats, rois, strides, pooled_size, roi_canonical_scale=224,
                             roi_canonical_level=4):
    """The reference graph models/FPN/builder.py:573-605 literally: assign, zero the roi on the
    other levels, run ROIAlign_v2 on every level with all rois, add_n."""
    rois = _f32(rois)
    idx = fpn_assign_levels(rois, strides, roi_canonical_scale, roi_canonical_level).reshape(
        rois.shape[:2])
    total = None
    for i, (f, s) in enumerate(zip(feats, strides)):
        lvl_rois = np.where((idx == i)[..., None], rois, np.float32(0))
        o = roi_align_v2_forward(f, lvl_rois, pooled_size, 1.0 / s, with_argmax=False)
        total = o if total is None else total + o
    return total, idx


# ---------------------------------------------------------------------------------------------
# box decode / proposal / NMS (oracle/box_ops.c)
# ---------------------------------------------------------------------------------------------
def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def decode_bbox(rois, bbox_pred, im_info, bbox_mean=(0, 0, 0, 0), bbox_std=(0.1, 0.1, 0.2, 0.2),
                class_agnostic=True, bbox_decode_type="xywh"):
    """_contrib_DecodeBBox (decodebbox.cc:34-133)."""
    rois, bbox_pred, im_info = _f32(rois), _f32(bbox_pred), _f32(im_info)
    B, N, _ = rois.shape
    K4 = bbox_pred.shape[2]
    out = np.zeros((B, N, 4 if class_agnostic else K4), np.float32)
    m, s = _f32(bbox_mean), _f32(bbox_std)
    lib().oracle_decode_bbox(_p(rois), _p(bbox_pred), _p(im_info), B, N, K4, _p(m), _p(s),
                             int(bool(class_agnostic)), 0 if bbox_decode_type == "xywh" else 1, _p(out))
    return out


def generate_anchors_v3(feature_stride, ratios, scales):
    r, s = _f32(ratios), _f32(scales)
    out = np.empty((len(r) * len(s), 4), np.float32)
    lib().oracle_generate_anchors_v3(int(feature_stride), _p(r), len(r), _p(s), len(s), _p(out))
    return out


def proposal_v3(cls_prob, bbox_pred, im_info, feature_stride=16, scales=(4, 8, 16, 32),
                ratios=(0.5, 1, 2), rpn_pre_nms_top_n=6000, rpn_post_nms_top_n=300, threshold=0.7,
                rpn_min_size=16, iou_loss=False, is_train=False, debug=False):
    """_contrib_Proposal_v3, GPU (.cu) semantics.  -> (rois (B,post,4), scores (B,post,1))."""
    cls_prob, bbox_pred, im_info = _f32(cls_prob), _f32(bbox_pred), _f32(im_info)
    B, A2, H, W = cls_prob.shape
    A = A2 // 2
    assert A == len(scales) * len(ratios) and bbox_pred.shape == (B, 4 * A, H, W)
    count = A * H * W
    pre = min(rpn_pre_nms_top_n if rpn_pre_nms_top_n > 0 else count, count)
    post = rpn_post_nms_top_n if not is_train else min(rpn_post_nms_top_n, pre)
    out = np.empty((B, post, 4), np.float32)
    sc = np.empty((B, post, 1), np.float32)
    dets = np.empty((B, pre, 5), np.float32) if debug else None
    keep = np.empty((B, pre), np.int32) if debug else None
    nkeep = np.empty((B,), np.int32) if debug else None
    r, s = _f32(ratios), _f32(scales)
    lib().oracle_proposal_v3(_p(cls_prob), _p(bbox_pred), _p(im_info), B, A, H, W, int(feature_stride),
                             _p(s), len(s), _p(r), len(r), int(rpn_pre_nms_top_n),
                             int(rpn_post_nms_top_n), ctypes.c_float(threshold), int(rpn_min_size),
                             int(bool(iou_loss)), int(bool(is_train)), _p(out), _p(sc), _p(dets),
                             _p(keep), _p(nkeep))
    return (out, sc, dets, keep, nkeep) if debug else (out, sc)


def contrib_nms(proposals, rpn_pre_nms_top_n=6000, rpn_post_nms_top_n=300, threshold=0.7,
                already_sorted=False):
    """_contrib_NMS (nms.cu:274-364) -> (out (B,post,4), score (B,post,1)); rows >= min(post,pre)
    are left untouched by the reference and returned as NaN here."""
    proposals = _f32(proposals)
    B, count, _ = proposals.shape
    out = np.full((B, rpn_post_nms_top_n, 4), np.nan, np.float32)
    sc = np.full((B, rpn_post_nms_top_n, 1), np.nan, np.float32)
    lib().oracle_contrib_nms(_p(proposals), B, count, int(rpn_pre_nms_top_n), int(rpn_post_nms_top_n),
                             ctypes.c_float(threshold), int(bool(already_sorted)), _p(out), _p(sc))
    return out, sc


def stable_argsort_desc(score):
    score = _f32(score).ravel()
    order = np.empty(score.shape[0], np.int32)
    lib().oracle_stable_argsort_desc(_p(score), score.shape[0], _p(order))
    return order


def bbox_overlaps(boxes, query_boxes):
    """bbox_overlaps_cython (bbox.pyx:32-73)."""
    boxes, query_boxes = _f32(boxes), _f32(query_boxes)
    out = np.empty((boxes.shape[0], query_boxes.shape[0]), np.float32)
    lib().oracle_bbox_overlaps(_p(boxes), boxes.shape[0], _p(query_boxes), query_boxes.shape[0], _p(out))
    return out


def bbox_selfoverlaps(boxes, query_boxes):
    """bbox_selfoverlaps_cython (bbox_self.pyx:32-75): intersection / area(boxes[n])."""
    boxes, query_boxes = _f32(boxes), _f32(query_boxes)
    out = np.empty((boxes.shape[0], query_boxes.shape[0]), np.float32)
    lib().oracle_bbox_selfoverlaps(_p(boxes), boxes.shape[0], _p(query_boxes), query_boxes.shape[0], _p(out))
    return out


def greedy_nms(dets, thresh, order=None):
    """greedy_nms (cpu_nms.pyx:37-87) -> kept indices ascending (np.where(suppressed == 0)[0]).
    `order` defaults to the reference's own `scores.argsort()[::-1]`."""
    dets = _f32(dets)
    if order is None:
        order = dets[:, 4].argsort()[::-1]
    order = np.ascontiguousarray(order, dtype=np.int64)
    sup = np.empty(dets.shape[0], np.uint8)
    lib().oracle_greedy_nms(_p(dets), dets.shape[0], _p(order), ctypes.c_float(thresh), _p(sup))
    return np.where(sup == 0)[0]


def soft_nms(boxes_in, sigma=0.5, Nt=0.3, threshold=0.001, method=0):
    """soft_nms (cpu_nms.pyx:98-203) -> (boxes[:N], inds[:N])."""
    boxes = _f32(boxes_in).copy()
    inds = np.empty(boxes.shape[0], np.int64)
    lib().oracle_soft_nms.restype = ctypes.c_int
    n = lib().oracle_soft_nms(_p(boxes), _p(inds), boxes.shape[0], ctypes.c_float(sigma),
                              ctypes.c_float(Nt), ctypes.c_float(threshold), ctypes.c_uint(method))
    return boxes[:n], inds[:n]


def ref_cython():
    """The reference's own Cython modules built by oracle/build_ref.py (bbox, bbox_self, cpu_nms),
    or None when oracle/_ref is absent."""
    import importlib
    import sys

    d = os.path.join(_HERE, "_ref")
    if not os.path.isdir(d):
        return None
    if not hasattr(np, "float"):
        np.float = float  # the reference uses the alias NumPy removed (bbox_transform.py:20)
    if d not in sys.path:
        sys.path.insert(0, d)
    try:
        return {m: importlib.import_module(m) for m in ("bbox", "bbox_self", "cpu_nms")}
    except ImportError:
        return None


# ---------------------------------------------------------------------------------------------
# ProposalTarget (oracle/target_ops.c)
# ---------------------------------------------------------------------------------------------
def proposal_target(rois, gt_boxes, priorities, num_classes, image_rois, fg_fraction=0.25, fg_thresh=0.5,
                    bg_thresh_hi=0.5, bg_thresh_lo=0.0, proposal_without_gt=False, class_agnostic=False,
                    bbox_mean=(0, 0, 0, 0), bbox_std=(0.1, 0.1, 0.2, 0.2), bbox_weight=(1, 1, 1, 1),
                    return_match=False, valid_ranges=None, filter_scales=False):
    """ProposalTarget with injected shuffle priorities (B, D>=3, R+G) uint32.
    -> rois (B,IR,4), label (B,IR), bbox_target (B,IR,NC*4), bbox_weight (B,IR,NC*4),
       match_gt_iou (B,IR), kept (B,IR) int32."""
    rois, gt_boxes = _f32(rois), _f32(gt_boxes)
    B, R, _ = rois.shape
    G = gt_boxes.shape[1]
    pr = np.ascontiguousarray(priorities, dtype=np.uint32)
    assert pr.shape[0] == B and pr.shape[2] == R + G and pr.shape[1] >= 3
    no_cap = image_rois == -1  # ProposalTarget_v2: keep every fg roi, R rows per image
    if no_cap:
        image_rois = R
    vr = _f32(valid_ranges) if valid_ranges is not None else None
    IR, NC4 = image_rois, num_classes * 4
    o_rois = np.empty((B, IR, 4), np.float32)
    o_lab = np.empty((B, IR), np.float32)
    o_tgt = np.empty((B, IR, NC4), np.float32)
    o_wgt = np.empty((B, IR, NC4), np.float32)
    o_iou = np.empty((B, IR), np.float32)
    kept = np.empty((B, IR), np.int32)
    gt_index = np.empty((B, IR), np.int32)
    fg_count = np.empty((B,), np.int32)
    m, s, w = _f32(bbox_mean), _f32(bbox_std), _f32(bbox_weight)
    lib().oracle_proposal_target(_p(rois), _p(gt_boxes), B, R, G, int(num_classes), int(image_rois),
                                 ctypes.c_float(fg_fraction), ctypes.c_float(fg_thresh),
                                 ctypes.c_float(bg_thresh_hi), ctypes.c_float(bg_thresh_lo),
                                 int(bool(proposal_without_gt)), int(bool(class_agnostic)), _p(m), _p(s), _p(w),
                                 _p(pr), pr.shape[1], _p(o_rois), _p(o_lab), _p(o_tgt), _p(o_wgt), _p(o_iou),
                                 _p(kept), _p(gt_index), _p(fg_count), _p(vr), int(bool(filter_scales)),
                                 int(no_cap))
    if return_match:
        return o_rois, o_lab, o_tgt, o_wgt, o_iou, kept, gt_index, fg_count
    return o_rois, o_lab, o_tgt, o_wgt, o_iou, kept


def poly2mask(roi, poly, mask_size):
    """convertPoly2Mask (proposal_mask_target.cc:155-213) -> (M,M) float32 of 0/1."""
    roi, poly = _f32(roi), _f32(poly)
    out = np.empty(mask_size * mask_size, np.float32)
    lib().oracle_poly2mask(_p(roi), _p(poly), int(mask_size), _p(out))
    return out.reshape(mask_size, mask_size)


def poly2mask_ratio(roi, poly, mask_size):
    """convertPoly2MaskWithRatio (proposal_mask_target.cc:20-152) -> ((M,M) float32 mask, ratio as float64)."""
    roi, poly = _f32(roi), _f32(poly)
    out = np.empty(mask_size * mask_size, np.float32)
    f = lib().oracle_poly2mask_ratio
    f.restype = ctypes.c_double
    ratio = f(_p(roi), _p(poly), int(mask_size), _p(out))
    return out.reshape(mask_size, mask_size), ratio


def proposal_mask_target(rois, gt_boxes, gt_polys, priorities, num_classes, image_rois, mask_size,
                         fg_fraction=0.25, output_ratio=False, **kw):
    """ProposalMaskTarget (proposal_mask_target-inl.h:139-337, proposal_mask_target.cc:219-379) =
    ProposalTarget + masks of the first fg rows; mask_target (B, int(IR*fg_fraction), M, M) is
    pre-filled with -1 (-inl.h:242-243).  output_ratio adds mask_ratio (B, int(IR*fg_fraction)), pre-filled
    with 0 (-inl.h:244), and switches the mask's vertex transform to double (.cc:368-371)."""
    r = proposal_target(rois, gt_boxes, priorities, num_classes, image_rois, fg_fraction, return_match=True, **kw)
    o_rois, gt_index, fg_count = r[0], r[6], r[7]
    gt_polys = _f32(gt_polys)
    B = o_rois.shape[0]
    nm = int(image_rois * fg_fraction)
    mask = np.full((B, nm, mask_size, mask_size), -1, np.float32)
    ratio = np.zeros((B, nm), np.float32)
    for b in range(B):
        for i in range(min(int(fg_count[b]), nm)):
            if output_ratio:
                mask[b, i], ratio[b, i] = poly2mask_ratio(o_rois[b, i], gt_polys[b, gt_index[b, i]], mask_size)
            else:
                mask[b, i] = poly2mask(o_rois[b, i], gt_polys[b, gt_index[b, i]], mask_size)
    return r[:5] + ((mask, ratio) if output_ratio else (mask,))


# ---------------------------------------------------------------------------------------------
# losses (oracle/loss_ops.c)
# ---------------------------------------------------------------------------------------------
def sigmoid(x):
    x = _f32(x)
    y = np.empty_like(x)
    lib().oracle_sigmoid(_p(x), ctypes.c_long(x.size), _p(y))
    return y


def focal_loss_backward(out, label, alpha=0.25, gamma=2.0, grad_scale=1.0, normalization="null", ograd=None):
    out, label = _f32(out), _f32(label)
    B, N, K = out.shape
    g = np.empty_like(out)
    og = _f32(ograd) if ograd is not None else None
    lib().oracle_focal_loss_backward(_p(out), _p(label), _p(og), B, N, K, ctypes.c_float(alpha),
                                     ctypes.c_float(gamma), ctypes.c_float(grad_scale),
                                     {"null": 0, "batch": 1, "valid": 2}[normalization], _p(g))
    return g


def bbox_norm_backward(gout, label):
    gout, label = _f32(gout), _f32(label)
    g = np.empty_like(gout)
    lib().oracle_bbox_norm_backward(_p(gout), ctypes.c_long(gout.size), _p(label), ctypes.c_long(label.size), _p(g))
    return g


def sigmoid_ce_forward(data, label):
    data, label = _f32(data), _f32(label)
    R, D = data.shape
    out = np.empty((R,), np.float32)
    lib().oracle_sigmoid_ce_forward(_p(data), _p(label), R, ctypes.c_long(D), _p(out))
    return out


def sigmoid_ce_backward(data, label, scale=1.0):
    data, label = _f32(data), _f32(label)
    R, D = data.shape
    dx = np.empty_like(data)
    lib().oracle_sigmoid_ce_backward(_p(data), _p(label), R, ctypes.c_long(D), ctypes.c_float(scale), _p(dx))
    return dx


def proposal_legacy(cls_prob, bbox_pred, im_info, version=1, valid_ranges=None, feature_stride=16,
                    scales=(4, 8, 16, 32), ratios=(0.5, 1, 2), rpn_pre_nms_top_n=6000, rpn_post_nms_top_n=300,
                    threshold=0.7, rpn_min_size=16, iou_loss=False, is_train=False, filter_scales=False):
    """_contrib_Proposal (version=1) / _contrib_Proposal_v2 (version=2): legacy pipeline."""
    cls_prob, bbox_pred, im_info = _f32(cls_prob), _f32(bbox_pred), _f32(im_info)
    B, A2, H, W = cls_prob.shape
    A = A2 // 2
    count = A * H * W
    pre = min(rpn_pre_nms_top_n if rpn_pre_nms_top_n > 0 else count, count)
    post = min(rpn_post_nms_top_n, pre)
    if version == 1 and not is_train:
        post = rpn_post_nms_top_n
    out = np.empty((B, post, 4), np.float32)
    sc = np.empty((B, post, 1), np.float32)
    vr = _f32(valid_ranges) if valid_ranges is not None else None
    r, s_ = _f32(ratios), _f32(scales)
    lib().oracle_proposal_legacy(_p(cls_prob), _p(bbox_pred), _p(im_info), _p(vr), int(version), B, A, H, W,
                                 int(feature_stride), _p(s_), len(s_), _p(r), len(r), int(rpn_pre_nms_top_n),
                                 int(rpn_post_nms_top_n), ctypes.c_float(threshold), int(rpn_min_size),
                                 int(bool(iou_loss)), int(bool(is_train)), int(bool(filter_scales)), _p(out), _p(sc))
    return out, sc


def gen_anchor(H, W, feature_stride=16, scales=(4, 8, 16, 32), ratios=(0.5, 1, 2)):
    """_contrib_GenAnchor -> (H*W*A, 4) float32."""
    sc = np.ascontiguousarray(scales, np.float64)
    ra = np.ascontiguousarray(ratios, np.float64)
    out = np.empty((H * W * len(sc) * len(ra), 4), np.float32)
    lib().oracle_gen_anchor(int(H), int(W), int(feature_stride), _p(sc), len(sc), _p(ra), len(ra), _p(out))
    return out


def gen_proposal(cls_prob, bbox_pred, im_info, anchors, feature_stride=16, rpn_pre_nms_top_n=6000,
                 rpn_min_size=16, iou_loss=False):
    """_contrib_GenProposal -> (B, pre, 5)."""
    cls_prob, bbox_pred, im_info, anchors = _f32(cls_prob), _f32(bbox_pred), _f32(im_info), _f32(anchors)
    B, A2, H, W = cls_prob.shape
    out = np.empty((B, rpn_pre_nms_top_n, 5), np.float32)
    lib().oracle_gen_proposal(_p(cls_prob), _p(bbox_pred), _p(im_info), _p(anchors), B, A2 // 2, H, W,
                              int(feature_stride), int(rpn_pre_nms_top_n), int(rpn_min_size), int(bool(iou_loss)),
                              _p(out))
    return out


def gen_proposal_retina(cls_prob, bbox_pred, im_info, anchors, num_anchors, rpn_pre_nms_top_n=1000,
                        rpn_min_size=0, thresh=0.0, anchor_mean=(0, 0, 0, 0), anchor_std=(1, 1, 1, 1),
                        output_one_hot=True):
    """_contrib_GenProposalRetina -> (bbox (B,pre,4), score (B,pre,K+1|1))."""
    cls_prob, bbox_pred, im_info, anchors = _f32(cls_prob), _f32(bbox_pred), _f32(im_info), _f32(anchors)
    B, AK, H, W = cls_prob.shape
    K = AK // num_anchors
    oc = K + 1 if output_one_hot else 1
    out = np.empty((B, rpn_pre_nms_top_n, 4), np.float32)
    sc = np.empty((B, rpn_pre_nms_top_n, oc), np.float32)
    m, s_ = _f32(anchor_mean), _f32(anchor_std)
    lib().oracle_gen_proposal_retina(_p(cls_prob), _p(bbox_pred), _p(im_info), _p(anchors), B, AK, H, W,
                                     int(num_anchors), int(rpn_pre_nms_top_n), int(rpn_min_size),
                                     ctypes.c_float(thresh), _p(m), _p(s_), int(bool(output_one_hot)), _p(out), _p(sc))
    return out, sc


def set_threads(n=None):
    """OpenMP team size of the C restatements.  torchrun exports OMP_NUM_THREADS=1 to its children;
    the CPU baseline is meant to use every host core, so bench.py calls this explicitly."""
    import os
    n = int(n or os.cpu_count() or 1)
    lib()  # make sure libgomp is mapped
    for name in ("libgomp.so.1", "libgomp.so"):
        try:
            ctypes.CDLL(name).omp_set_num_threads(n)
            return n
        except OSError:
            continue
    return 0


# ---- test-time mask paste (oracle/paste_ops.c) ---------------------------------------------------------------------
def resize_linear_f32(src, dsize):
    """cv2.resize(src float32 (sh, sw), dsize=(dw, dh)) with the default INTER_LINEAR, as the installed wheel computes it."""
    src = _f32(src)
    dw, dh = int(dsize[0]), int(dsize[1])
    dst = np.empty((dh, dw), np.float32)
    lib().oracle_resize_linear_f32(_p(src), src.shape[0], src.shape[1], _p(dst), dh, dw)
    return dst


def segm_paste(box, mask, im_h, im_w):
    """models/maskrcnn/utils.py:39-59 for one detection -> (im_mask (im_h, im_w) uint8, ok); ok False where the
    reference's slice assignment cannot work (box entirely outside the image)."""
    box, mask = _f32(box), _f32(mask)
    M = mask.shape[-1]
    rb = np.empty(4, np.int32)
    lib().oracle_expand_box_int(_p(box), M, _p(rb))
    w, h = max(int(rb[2] - rb[0]) + 1, 1), max(int(rb[3] - rb[1]) + 1, 1)
    scratch = np.empty((M + 2) * (M + 2) + w * h, np.float32)
    im = np.zeros((im_h, im_w), np.uint8)
    rc = lib().oracle_segm_paste(_p(box), _p(mask), M, int(im_h), int(im_w), _p(im), _p(scratch))
    return im, rc == 0


def rle_encode(im):
    """pycocotools rleEncode of one (h, w) uint8 mask -> run lengths (uint32), the first run counting zeros."""
    im = np.ascontiguousarray(im, np.uint8)
    h, w = im.shape
    counts = np.empty(h * w + 1, np.uint32)
    fn = lib().oracle_rle_encode
    fn.restype = ctypes.c_long
    k = fn(_p(im), h, w, _p(counts))
    return counts[:k].copy()
