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
