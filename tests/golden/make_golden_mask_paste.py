#!/usr/bin/env python
"""Generates tests/golden/reference_mask_paste.npz by running the reference's OWN `segm_results`
(models/maskrcnn/utils.py:26-67, unmodified) on top of the cv2 installed here and a stand-in for
`pycocotools.mask.encode` (pycocotools is absent: the oracle's restatement of cocoapi's rleEncode + rleToString).
So expand_boxes, the int32 truncation, the zero ring, cv2.resize, the threshold and the paste slices are the
reference's and cv2's own; only the run-length string codec is a restatement.
Run:  python tests/golden/make_golden_mask_paste.py     (needs /root/reference and cv2)"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import np_ops  # noqa: E402

REF = "/root/reference"


def blobs(rng, n, k, m):
    """Smooth mask probabilities: sigmoid of a few random low-frequency bumps, so the thresholded shape is a blob
    with holes and islands, plus a checkerboard and an all-ones / all-zeros mask as stress cases."""
    yy, xx = np.mgrid[0:m, 0:m].astype(np.float32) / m
    out = np.zeros((n, k, m, m), np.float32)
    for i in range(n):
        for c in range(k):
            z = np.full((m, m), -1.0, np.float32)
            for _ in range(int(rng.integers(2, 6))):
                cx, cy, s, a = rng.uniform(0, 1), rng.uniform(0, 1), rng.uniform(0.08, 0.4), rng.uniform(1.5, 5)
                z += (a * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * s * s))).astype(np.float32)
            out[i, c] = 1 / (1 + np.exp(-3 * z))
    out[0, :] = (np.indices((m, m)).sum(0) % 2).astype(np.float32)     # checkerboard: many flips per column
    out[1, :] = 1.0
    out[2, :] = 0.0
    out[3, :] = 0.5                                                    # exactly at the threshold: `>` keeps it out
    return out


def cases():
    rng = np.random.default_rng(5)
    res = []
    for name, im_h, im_w, n, k, m in (("a", 240, 320, 28, 4, 28), ("b", 97, 131, 16, 3, 14)):
        xy = rng.uniform(-20, [im_w - 30, im_h - 30], (n, 2))
        wh = rng.uniform(25, [im_w * 0.7, im_h * 0.7], (n, 2))   # x2, y2 >= 5: inside the image (else the reference raises)
        box = np.concatenate([xy, xy + wh], 1).astype(np.float32)
        box[4] = [-15.5, -9.25, im_w + 12.0, im_h + 7.5]       # covers the whole image: full-height columns
        box[5] = [10.2, -3.0, 60.7, im_h + 4.0]                # full height, part of the width
        box[6] = [30.0, 40.0, 30.0, 40.0]                      # a point: 1..3 pixels after expansion
        box[7] = [0.0, 0.0, im_w - 1.0, im_h - 1.0]            # exactly the image
        box[8] = [im_w - 8.0, im_h - 6.0, im_w + 40.0, im_h + 30.0]   # sticks out at the bottom right
        box[9] = [50.4, 20.6, 50.9, 90.3]                      # sub-pixel wide
        box[1] = [5.0, 5.0, im_w * 0.6, im_h * 0.8]            # the all-ones mask on a large box
        cls = rng.integers(0, k, n).astype(np.int32)
        res.append((name, im_h, im_w, box, cls, blobs(rng, n, k, m)))
    return res


def main():
    import cv2  # noqa: F401  (the reference imports it)

    pm = types.ModuleType("pycocotools")
    mk = types.ModuleType("pycocotools.mask")

    def encode(arr):  # arr (h, w, n) uint8 Fortran order -> list of RLE dicts, as pycocotools
        h, w, n = arr.shape
        return [{"size": [h, w], "counts": np_ops.rle_to_string(oracle.rle_encode(np.ascontiguousarray(arr[:, :, i])))}
                for i in range(n)]

    mk.encode = encode
    pm.mask = mk
    sys.modules["pycocotools"], sys.modules["pycocotools.mask"] = pm, mk
    sys.path.insert(0, REF)
    from models.maskrcnn.utils import segm_results

    d = {"names": np.array([c[0] for c in cases()])}
    for name, im_h, im_w, box, cls, masks in cases():
        segms = segm_results(box, cls, masks, im_h, im_w)
        d[f"{name}_hw"] = np.array([im_h, im_w])
        d[f"{name}_box"], d[f"{name}_cls"], d[f"{name}_masks"] = box, cls, masks
        d[f"{name}_counts"] = np.array([s["counts"] for s in segms])
        assert all(s["size"] == [im_h, im_w] for s in segms)
        print(name, "detections", len(segms), "longest RLE string", max(len(s["counts"]) for s in segms))
    np.savez_compressed(os.path.join(HERE, "reference_mask_paste.npz"), **d)


if __name__ == "__main__":
    main()
