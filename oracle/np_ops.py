"""ORACLE — TEST INFRASTRUCTURE ONLY.  numpy restatements of the reference's pure-Python operators
(the originals import mxnet at module top, so the numpy bodies are restated here; each function
cites the reference file:line it follows).  "Parity unpinned" unless noted."""
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


# ---- DCNv1 sampling: restated from the published formulation (upstream MXNet
# src/operator/contrib/nn/deformable_im2col.h/.cuh; not in the reference tree) — PARITY UNPINNED ----
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
