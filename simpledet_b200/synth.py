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
