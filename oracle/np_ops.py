"""ORACLE — TEST INFRASTRUCTURE ONLY.  numpy restatements of the reference's pure-Python operators
(the originals import mxnet at module top, so the numpy bodies are restated here; each function
cites the reference file:line it follows).  Pinned by goldens generated from the reference's own Python
(tests/golden/make_golden*.py, tests/test_oracle_golden*.py) except where a function says "PARITY UNPINNED"."""
import numpy as np

import oracle


def get_top_proposal(bboxes, scores, top_n):
    """models/FPN/get_top_proposal.py:15-40.  mx.nd.argsort(is_ascend=False) is treated as a
    stable descending sort (ties keep the lower index first) — the same order as thrust's
    stable_sort_by_key(greater) the rest of the reference uses."""
    B = bboxes.shape[0]
    ob = np.empty((B, top_n, 4), np.float32)
    os_ = np.empty((B, top_n, 1), np.float32)
    for i in range(B):
        order = oracle.stable_argsort_desc(scores[i, :, 0])[:top_n]
        ob[i] = bboxes[i][order]
        os_[i] = scores[i][order]
    return ob, os_


def py_nms(dets, thresh):
    """operator_py/nms.py:41-75 `nms` (float32 in, keeps ovr <= thresh) -> kept rows.
    `scores.argsort()[::-1]` leaves the order of tied scores to numpy's introsort; callers that
    need a pinned order pass distinct scores."""
    dets = np.asarray(dets, np.float32)
    x1, y1, x2, y2, scores = dets[:, 0], dets[:, 1], dets[:, 2], dets[:, 3], dets[:, 4]
    areas = (x2 - x1 + 1) * (y2 - y1 + 1)
    order = scores.argsort()[::-1]
    keep = []
    while order.size > 0:
        i = order[0]
        keep.append(i)
        xx1 = np.maximum(x1[i], x1[order[1:]])
        yy1 = np.maximum(y1[i], y1[order[1:]])
        xx2 = np.minimum(x2[i], x2[order[1:]])
        yy2 = np.minimum(y2[i], y2[order[1:]])
        w = np.maximum(0.0, xx2 - xx1 + 1)
        h = np.maximum(0.0, yy2 - yy1 + 1)
        inter = w * h
        ovr = inter / (areas[i] + areas[order[1:]] - inter)
        inds = np.where(ovr <= thresh)[0]
        order = order[inds + 1]
    return dets[keep, :]


def _pair_iou(dets, i, js):
    """IoU of box i with boxes js, float32, the expression order of operator_py/nms.py:62-69."""
    x1, y1, x2, y2 = dets[:, 0], dets[:, 1], dets[:, 2], dets[:, 3]
    areas = (x2 - x1 + 1) * (y2 - y1 + 1)
    w = np.maximum(0.0, np.minimum(x2[i], x2[js]) - np.maximum(x1[i], x1[js]) + 1)
    h = np.maximum(0.0, np.minimum(y2[i], y2[js]) - np.maximum(y1[i], y1[js]) + 1)
    inter = w * h
    return inter / (areas[i] + areas[js] - inter)


def set_nms(dets, thresh):
    """operator_py/nms.py:77-107: greedy NMS in which boxes of the same set (column 5) never
    suppress each other -> kept rows (m', 6)."""
    dets = np.asarray(dets, np.float32)
    order = dets[:, 4].argsort()[::-1]
    keep = []
    while order.size:
        i, rest = order[0], order[1:]
        keep.append(i)
        ovr = _pair_iou(dets, i, rest)
        order = rest[(ovr <= thresh) | (dets[rest, 5] == dets[i, 5])]
    return dets[keep, :]


def py_weighted_nms(dets, thresh_lo, thresh_hi):
    """operator_py/nms.py:110-157: every surviving top box is replaced by the score-weighted mean of
    the remaining boxes with IoU > thresh_hi; boxes with IoU > thresh_lo leave the pool."""
    dets = np.asarray(dets, np.float32)
    scores = dets[:, 4]
    order = scores.argsort()[::-1]
    out = []
    while order.size:
        i = order[0]
        ovr = _pair_iou(dets, i, order)
        voters = order[ovr > thresh_hi]
        if len(voters) == 0:
            break
        sw = np.sum(scores[voters])
        out.append([np.sum(scores[voters] * dets[voters, k]) / sw for k in range(4)] + [scores[i]])
        order = order[ovr <= thresh_lo]
    return np.array(out)


def do_nms(cls_score, bbox_xyxy, nms_thresh, min_det_score):
    """detection_test.py:233-260 for one image: {cid: kept dets (m,5)}."""
    out = {}
    for cid in range(cls_score.shape[1]):
        score = cls_score[:, cid]
        cls_box = bbox_xyxy[:, cid * 4:(cid + 1) * 4] if bbox_xyxy.shape[1] != 4 else bbox_xyxy
        valid = np.where(score > min_det_score)[0]
        det = np.concatenate((cls_box[valid], score[valid].reshape(-1, 1)), axis=1).astype(np.float32)
        out[cid] = py_nms(det, nms_thresh) if det.shape[0] else det
    return out


def final_detections(cls_score, bbox_xyxy, nms_thresh, min_det_score, max_det):
    """detection_test.py:233-291 for one image: do_nms per class, then the COCO rows
    sorted(result, key=score)[-max_det:] -> (rows, 6) [x, y, w, h, score, class index]."""
    per_class = do_nms(cls_score, bbox_xyxy, nms_thresh, min_det_score)
    result = []
    for cid, det in per_class.items():
        for k in range(det.shape[0]):
            result.append((float(det[k, 0]), float(det[k, 1]), float(det[k, 2] - det[k, 0] + 1),
                           float(det[k, 3] - det[k, 1] + 1), float(det[k, 4]), cid))
    result = sorted(result, key=lambda r: r[4])[-max_det:]
    return np.array(result, np.float32).reshape(-1, 6)


# ---- operator_py/bbox_transform.py (numpy float64 utilities) — PINNED by tests/golden ----------
BBOX_XFORM_CLIP = np.log(1000. / 16.)  # bbox_transform.py:5


def clip_boxes(boxes, im_shape):
    """bbox_transform.py:34-49 (in place on a copy here)."""
    boxes = boxes.copy()
    boxes[:, 0::4] = np.maximum(np.minimum(boxes[:, 0::4], im_shape[1] - 1), 0)
    boxes[:, 1::4] = np.maximum(np.minimum(boxes[:, 1::4], im_shape[0] - 1), 0)
    boxes[:, 2::4] = np.maximum(np.minimum(boxes[:, 2::4], im_shape[1] - 1), 0)
    boxes[:, 3::4] = np.maximum(np.minimum(boxes[:, 3::4], im_shape[0] - 1), 0)
    return boxes


def nonlinear_transform(ex_rois, gt_rois):
    """bbox_transform.py:52-78."""
    ew = ex_rois[:, 2] - ex_rois[:, 0] + 1.0
    eh = ex_rois[:, 3] - ex_rois[:, 1] + 1.0
    ecx = ex_rois[:, 0] + 0.5 * (ew - 1.0)
    ecy = ex_rois[:, 1] + 0.5 * (eh - 1.0)
    gw = gt_rois[:, 2] - gt_rois[:, 0] + 1.0
    gh = gt_rois[:, 3] - gt_rois[:, 1] + 1.0
    gcx = gt_rois[:, 0] + 0.5 * (gw - 1.0)
    gcy = gt_rois[:, 1] + 0.5 * (gh - 1.0)
    return np.vstack(((gcx - ecx) / (ew + 1e-14), (gcy - ecy) / (eh + 1e-14), np.log(gw / ew),
                      np.log(gh / eh))).transpose()


def nonlinear_pred(boxes, box_deltas):
    """bbox_transform.py:81-120 (float64; dw, dh clipped at log(1000/16))."""
    if boxes.shape[0] == 0:
        return np.zeros((0, box_deltas.shape[1]))
    boxes = boxes.astype(np.float64, copy=False)
    w = boxes[:, 2] - boxes[:, 0] + 1.0
    h = boxes[:, 3] - boxes[:, 1] + 1.0
    cx = boxes[:, 0] + 0.5 * (w - 1.0)
    cy = boxes[:, 1] + 0.5 * (h - 1.0)
    dx, dy = box_deltas[:, 0::4], box_deltas[:, 1::4]
    dw = np.minimum(box_deltas[:, 2::4], BBOX_XFORM_CLIP)
    dh = np.minimum(box_deltas[:, 3::4], BBOX_XFORM_CLIP)
    pcx = dx * w[:, None] + cx[:, None]
    pcy = dy * h[:, None] + cy[:, None]
    pw = np.exp(dw) * w[:, None]
    ph = np.exp(dh) * h[:, None]
    out = np.zeros(box_deltas.shape)
    out[:, 0::4] = pcx - 0.5 * (pw - 1.0)
    out[:, 1::4] = pcy - 0.5 * (ph - 1.0)
    out[:, 2::4] = pcx + 0.5 * (pw - 1.0)
    out[:, 3::4] = pcy + 0.5 * (ph - 1.0)
    return out


def iou_pred(boxes, box_deltas):
    """bbox_transform.py:129-161."""
    if boxes.shape[0] == 0:
        return np.zeros((0, box_deltas.shape[1]))
    boxes = boxes.astype(np.float64, copy=False)
    out = np.zeros(box_deltas.shape)
    for k in range(4):
        out[:, k::4] = box_deltas[:, k::4] + boxes[:, k][:, None]
    return out


def flip_boxes(boxes, im_width):
    """bbox_transform.py:164-169."""
    f = boxes.copy()
    f[:, 0::4] = im_width - boxes[:, 2::4] - 1
    f[:, 2::4] = im_width - boxes[:, 0::4] - 1
    return f


def box_voting(top_dets, all_dets, thresh=0.5, scoring_method="ID", beta=1.0, overlaps_fn=None):
    """operator_py/bbox_transform.py:172-221: every top box becomes the score-weighted mean of the boxes
    of `all_dets` overlapping it by >= thresh (float32 Cython IoU); the score is re-derived by
    `scoring_method`."""
    from . import bbox_overlaps as c_overlaps
    overlaps_fn = overlaps_fn or c_overlaps
    out = top_dets.copy()
    boxes, scores = all_dets[:, :4], all_dets[:, 4]
    ov = overlaps_fn(top_dets[:, :4], boxes)
    for k in range(out.shape[0]):
        sel = np.where(ov[k] >= thresh)[0]
        ws = scores[sel]
        out[k, :4] = np.average(boxes[sel, :], axis=0, weights=ws)
        if scoring_method == "ID":
            pass
        elif scoring_method == "TEMP_AVG":
            P = np.vstack((ws, 1.0 - ws))
            X = np.log(P / np.max(P, axis=0))
            E = np.exp(X / beta)
            out[k, 4] = (E / np.sum(E, axis=0))[0].mean()
        elif scoring_method == "AVG":
            out[k, 4] = ws.mean()
        elif scoring_method == "IOU_AVG":
            out[k, 4] = np.average(ws, weights=ov[k, sel])
        elif scoring_method == "GENERALIZED_AVG":
            out[k, 4] = np.mean(ws ** beta) ** (1.0 / beta)
        elif scoring_method == "QUASI_SUM":
            out[k, 4] = ws.sum() / float(len(ws)) ** beta
        else:
            raise NotImplementedError("Unknown scoring method {}".format(scoring_method))
    return out


# ---- DCNv1 sampling: restated from the published formulation (upstream MXNet
# src/operator/contrib/nn/deformable_im2col.h/.cuh; not in the reference tree) — PARITY UNPINNED (cross-checked against
# torchvision.ops.deform_conv2d, an independent implementation: tests/test_oracle_dcn_torchvision.py) ----
def deformable_im2col(data, offset, kernel, stride=(1, 1), dilate=(1, 1), pad=(0, 0), num_deformable_group=1):
    """data (B,C,H,W), offset (B, dg*2*KH*KW, Ho, Wo) -> col (B, C*KH*KW, Ho*Wo), float32 arithmetic."""
    f = np.float32
    data = np.asarray(data, f)
    offset = np.asarray(offset, f)
    B, C, H, W = data.shape
    kh, kw = kernel
    Ho = (H + 2 * pad[0] - (dilate[0] * (kh - 1) + 1)) // stride[0] + 1
    Wo = (W + 2 * pad[1] - (dilate[1] * (kw - 1) + 1)) // stride[1] + 1
    cpg = C // num_deformable_group
    col = np.zeros((B, C * kh * kw, Ho * Wo), f)
    hc, wc = np.meshgrid(np.arange(Ho), np.arange(Wo), indexing="ij")
    for b in range(B):
        for c in range(C):
            g = c // cpg
            im = data[b, c]
            for i in range(kh):
                for j in range(kw):
                    t = i * kw + j
                    oh = offset[b, g * 2 * kh * kw + 2 * t]
                    ow = offset[b, g * 2 * kh * kw + 2 * t + 1]
                    h = (hc * stride[0] - pad[0] + i * dilate[0]).astype(f) + oh
                    w = (wc * stride[1] - pad[1] + j * dilate[1]).astype(f) + ow
                    inside = (h >= 0) & (w >= 0) & (h < H) & (w < W)
                    hl = np.floor(h).astype(np.int64)
                    wl = np.floor(w).astype(np.int64)
                    hcl = hl >= H - 1
                    wcl = wl >= W - 1
                    hl = np.where(hcl, H - 1, hl)
                    wl = np.where(wcl, W - 1, wl)
                    hh_ = np.where(hcl, H - 1, hl + 1)
                    wh_ = np.where(wcl, W - 1, wl + 1)
                    h2 = np.where(hcl, hl.astype(f), h)
                    w2 = np.where(wcl, wl.astype(f), w)
                    hl_c, wl_c = np.clip(hl, 0, H - 1), np.clip(wl, 0, W - 1)
                    hh_c, wh_c = np.clip(hh_, 0, H - 1), np.clip(wh_, 0, W - 1)
                    lh = (h2 - hl.astype(f)).astype(f)
                    lw = (w2 - wl.astype(f)).astype(f)
                    hhw, hww = f(1) - lh, f(1) - lw
                    v = (hhw * hww * im[hl_c, wl_c] + hhw * lw * im[hl_c, wh_c] + lh * hww * im[hh_c, wl_c]
                         + lh * lw * im[hh_c, wh_c]).astype(f)
                    col[b, c * kh * kw + t] = np.where(inside, v, f(0)).reshape(-1)
    return col


# ---------------------------------------------------------------------------------------------
# AnchorTarget2D (core/detection_input.py:353-565) and PyramidAnchorTarget2D (models/FPN/input.py:55-148)
# ---------------------------------------------------------------------------------------------
def anchor_base(stride, scales, aspects):
    """base_anchor property (detection_input.py:377-403): (len(aspects)*len(scales), 4) float64,
    aspect-major; np.round = round-half-even."""
    side = float(stride)                      # w = h = stride of the (0,0,stride-1,stride-1) cell
    ctr = 0.5 * (side - 1)
    asp = np.asarray(aspects, np.float64)
    wr = np.round(np.sqrt(side * side / asp))
    hr = np.round(wr * asp)
    sc = np.asarray(scales, np.float64)
    ws, hs = np.outer(wr, sc).ravel(), np.outer(hr, sc).ravel()
    return np.stack([ctr - 0.5 * (ws - 1), ctr - 0.5 * (hs - 1), ctr + 0.5 * (ws - 1), ctr + 0.5 * (hs - 1)], 1)


def anchor_grid(fh, fw, stride, base):
    """v_all_anchor / h_all_anchor (:405-441): float32 shifts + float64 base -> (fh*fw*A, 4) float64,
    cell-major (y, x), anchor-minor."""
    sx = np.arange(fw, dtype=np.float32) * stride
    sy = np.arange(fh, dtype=np.float32) * stride
    gx, gy = np.meshgrid(sx, sy)
    cell = np.stack([gx.ravel(), gy.ravel(), gx.ravel(), gy.ravel()], 1)
    return (cell[:, None, :] + base[None, :, :]).reshape(-1, 4)


def anchor_target(im_info, gt_bbox, strides, shorts, longs, scales, aspects, allowed_border, neg_thr, pos_thr,
                  min_pos_thr, image_anchor, pos_fraction, priorities=None, overlaps_fn=None):
    """One image.  strides/shorts/longs: sequences over pyramid levels (length 1 = AnchorTarget2D).
    `priorities` (one float per anchor over ALL levels) replaces np.random.choice in _sample_anchor
    (:477-494): of the surplus fg (then bg) anchors, the ones with the SMALLEST priority are disabled,
    ties by larger index first; priorities = arange reproduces the reference's DEBUG mode (the first
    surplus indices are disabled).  Returns (cls_label (A*sum HW,), reg_target (4A, sum HW),
    reg_weight (4A, sum HW)) in the pyramid layout of models/FPN/input.py:117-140 — for one level
    that is AnchorTarget2D's (A*fh*fw,), (4A, fh, fw) memory order."""
    from . import bbox_overlaps as c_overlaps
    overlaps_fn = overlaps_fn or c_overlaps
    h, w = float(im_info[0]), float(im_info[1])
    vertical = h >= w                                                   # :449, :547-551
    dims = [(lo, sh) if vertical else (sh, lo) for sh, lo in zip(shorts, longs)]
    per_level = [anchor_grid(fh, fw, s, anchor_base(s, scales, aspects)) for (fh, fw), s in zip(dims, strides)]
    anchors = np.concatenate(per_level)
    n_all = anchors.shape[0]
    A = len(scales) * len(aspects)
    gt = np.asarray(gt_bbox, np.float32)
    gt = gt[gt[:, 0] != -1][:, :4]                                      # :531-535
    inside = np.flatnonzero((anchors[:, 0] >= -allowed_border) & (anchors[:, 1] >= -allowed_border) &
                            (anchors[:, 2] < w + allowed_border) & (anchors[:, 3] < h + allowed_border))
    va = anchors[inside]
    label = np.full(len(va), -1, np.float32)
    if len(gt):                                                         # _assign_label_to_anchor :455-475
        ov = overlaps_fn(va.astype(np.float32), gt)
        best = ov.max(1)
        which = ov.argmax(1)
        per_gt = ov.max(0)
        hit = ((ov == per_gt[None, :]) & (ov >= min_pos_thr)).any(1)
        label[best < neg_thr] = 0
        label[hit] = 1
        label[best >= pos_thr] = 1
    else:
        label[:] = 0
        which = np.zeros(len(va), np.int64)
    pr = np.arange(n_all, dtype=np.float64) if priorities is None else np.asarray(priorities, np.float64)
    pr = pr[inside]

    def cap(value, quota):                                              # _sample_anchor :477-494
        idx = np.flatnonzero(label == value)
        if len(idx) > quota:
            order = np.lexsort((idx, -pr[idx]))                         # priority desc, index asc
            label[idx[order[quota:]]] = -1
    cap(1, int(pos_fraction * image_anchor))
    cap(0, image_anchor - int(np.sum(label == 1)))
    tgt = np.zeros((len(va), 4), np.float32)
    wgt = np.zeros((len(va), 4), np.float32)
    fg = np.flatnonzero(label == 1)
    if len(fg):                                                         # _cal_anchor_target :496-506
        tgt[fg] = nonlinear_transform(va[fg], gt[which[fg], :4])
        wgt[fg] = 1.0
    all_label = np.full(n_all, -1, np.float32)
    all_tgt = np.zeros((n_all, 4), np.float32)
    all_wgt = np.zeros((n_all, 4), np.float32)
    all_label[inside], all_tgt[inside], all_wgt[inside] = label, tgt, wgt
    labs, tgts, wgts, o = [], [], [], 0
    for (fh, fw) in dims:                                               # per-level (h,w,A) -> (A, h*w)
        n = fh * fw * A
        labs.append(all_label[o:o + n].reshape(fh * fw, A).T)
        tgts.append(all_tgt[o:o + n].reshape(fh * fw, A * 4).T)
        wgts.append(all_wgt[o:o + n].reshape(fh * fw, A * 4).T)
        o += n
    return np.concatenate(labs, 1).reshape(-1), np.concatenate(tgts, 1), np.concatenate(wgts, 1)


# ---- pycocotools RLE strings + segm_results (models/maskrcnn/utils.py:26-67) --------------------------------------
def rle_to_string(counts):
    """cocoapi common/maskApi.c rleToString (published algorithm; pycocotools absent here: parity unpinned): each
    count, from the 4th on as the difference to the count two places back, in 5-bit groups, least significant first,
    bit 5 = continuation, bit 4 of the last group = sign, + 48."""
    out = bytearray()
    cnts = [int(c) for c in counts]
    for i, x in enumerate(cnts):
        if i > 2:
            x -= cnts[i - 2]
        more = True
        while more:
            c = x & 0x1F
            x >>= 5                      # arithmetic shift, like the C `long`
            more = (x != -1) if (c & 0x10) else (x != 0)
            if more:
                c |= 0x20
            out.append(c + 48)
    return bytes(out)


def rle_from_string(s):
    """maskApi.c rleFrString: the inverse of rle_to_string."""
    cnts, p, s = [], 0, bytes(s)
    while p < len(s):
        x, k, more = 0, 0, True
        while more:
            c = s[p] - 48
            x |= (c & 0x1F) << (5 * k)
            more = bool(c & 0x20)
            p += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if len(cnts) > 2:
            x += cnts[-2]
        cnts.append(x)
    return np.asarray(cnts, np.int64)


def segm_results(bbox_xyxy, cls, masks, im_h, im_w):
    """models/maskrcnn/utils.py:26-67 -> list of {'size': [im_h, im_w], 'counts': bytes} (what mask_util.encode
    returns per mask); a box with no pixel inside the image (the reference raises there) gives the empty mask."""
    out = []
    for box, m, c in zip(np.asarray(bbox_xyxy, np.float32), masks, cls):
        im, _ = oracle.segm_paste(box, m[int(c)], im_h, im_w)
        out.append({"size": [int(im_h), int(im_w)], "counts": rle_to_string(oracle.rle_encode(im))})
    return out
