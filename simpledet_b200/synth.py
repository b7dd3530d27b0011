"""Seeded synthetic workloads (numpy, CPU) shared by tests/ and bench.py.

Shapes follow SURVEY.md §8(d): 800x1333 input padded, FPN levels P2..P5 at strides 4..32.
"""
from __future__ import annotations

import numpy as np

FPN_STRIDES = (4, 8, 16, 32)
IMG_H, IMG_W = 800, 1333


def fpn_shapes(img_h: int = IMG_H, img_w: int = IMG_W, strides=FPN_STRIDES):
    """Feature-map sizes of a ResNet-FPN on a (img_h, img_w) input: ceil(size / stride)
    (config/faster_r50v1_fpn_1x.py:218-220: 200x334, 100x167, 50x84, 25x42)."""
    return [(-(-img_h // s), -(-img_w // s)) for s in strides]


def random_rois(rng: np.random.Generator, batch: int, n: int, img_h: int = IMG_H,
                img_w: int = IMG_W, min_side: float = 16.0, max_side: float = 512.0):
    """(batch, n, 4) float32 boxes with an FPN-like scale mix: sqrt(area) log-uniform in
    [min_side, max_side], aspect ratio log-uniform in [0.5, 2], centre uniform, clipped to the
    image like Proposal_v3's output (proposal_v3.cu:139-149)."""
    side = np.exp(rng.uniform(np.log(min_side), np.log(max_side), (batch, n)))
    ar = np.exp(rng.uniform(np.log(0.5), np.log(2.0), (batch, n)))
    w, h = side * np.sqrt(ar), side / np.sqrt(ar)
    cx, cy = rng.uniform(0, img_w, (batch, n)), rng.uniform(0, img_h, (batch, n))
    x1 = np.clip(cx - w / 2, 0, img_w - 1)
    y1 = np.clip(cy - h / 2, 0, img_h - 1)
    x2 = np.clip(cx + w / 2, 0, img_w - 1)
    y2 = np.clip(cy + h / 2, 0, img_h - 1)
    return np.stack([x1, y1, x2, y2], -1).astype(np.float32)


def config1(seed: int = 0):
    """BASELINE.json configs[0] / SURVEY §8(d) config 1: data N(0,1) (1,256,50,50); 128 rois
    x1,y1~U(0,700), w,h~U(8,400), clipped to [0,799]; spatial_scale 1/16; pooled 7x7."""
    rng = np.random.default_rng(seed)
    data = rng.standard_normal((1, 256, 50, 50)).astype(np.float32)
    xy = rng.uniform(0, 700, (1, 128, 2))
    wh = rng.uniform(8, 400, (1, 128, 2))
    rois = np.concatenate([xy, np.clip(xy + wh, 0, 799)], -1).astype(np.float32)
    return data, rois, (7, 7), 1.0 / 16


def mask_scene(rng: np.random.Generator, B: int, R: int, G: int, PL: int):
    """BASELINE config 4 inputs: rois (B,R,4) jittered around the gt boxes plus padding rows,
    gt (B,G,5) with class -1 padding, gt_polys (B,G,PL) encoded [cls, nseg, len_1..len_n, xy...] padded with -1
    (models/maskrcnn/input.py:166-175): 1-3 segments of 3-30 vertices per instance."""
    gt = np.full((B, G, 5), -1, np.float32)
    polys = np.full((B, G, PL), -1, np.float32)
    rois = np.zeros((B, R, 4), np.float32)
    for b in range(B):
        k = int(rng.integers(2, G))
        for j in range(k):
            x1, y1 = rng.uniform(0, 500, 2)
            w, h = rng.uniform(40, 300, 2)
            gt[b, j] = [x1, y1, x1 + w, y1 + h, rng.integers(1, 81)]
            nseg = int(rng.integers(1, 4))
            row, coords = [float(gt[b, j, 4]), float(nseg)], []
            for _ in range(nseg):
                nv = int(rng.integers(3, 30))
                ang = np.sort(rng.uniform(0, 2 * np.pi, nv))
                rad = rng.uniform(0.15, 0.5, nv)
                cx, cy = x1 + w * rng.uniform(0.3, 0.7), y1 + h * rng.uniform(0.3, 0.7)
                xs, ys = cx + w * rad * np.cos(ang), cy + h * rad * np.sin(ang)
                row.append(float(2 * nv))
                coords += np.stack([xs, ys], 1).reshape(-1).tolist()
            row += coords
            polys[b, j, :len(row)] = row
        m = R - 20
        near = gt[b, rng.integers(0, k, m), :4] + rng.normal(0, 12, (m, 4))
        near[:, 3] = np.maximum(near[:, 3], 1)
        rois[b, :m] = near
    return rois, gt, polys
