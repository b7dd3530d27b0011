"""Host-side estimate for the round-2 RoIAlign design question: how much staging traffic does sharing
windows between overlapping rois remove?  Rois of one (image, level) are sorted on a coarse grid and
merged greedily while the union window (rows x 16-byte padded row pitch) stays under the shared-memory
plane cap; reports staged cells with and without sharing.  No GPU involved."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simpledet_b200 import synth  # noqa: E402


def windows(rois, shapes, strides):
    w = rois[:, 2] - rois[:, 0] + 1
    h = rois[:, 3] - rois[:, 1] + 1
    lvl = np.clip(np.floor(4 + np.log2(np.sqrt(w * h) / 224 + 1e-6)), 2, 5).astype(int) - 2
    out = []
    for r, l in zip(rois, lvl):
        s = strides[l]
        H, W = shapes[l]
        x1, y1, x2, y2 = r / s
        out.append((l, int(np.clip(np.floor(y1), 0, H - 1)), int(np.clip(np.ceil(y2), 0, H - 1)),
                    int(np.clip(np.floor(x1), 0, W - 1)), int(np.clip(np.ceil(x2), 0, W - 1))))
    return out


def plane(y0, y1, x0, x1):
    return (y1 - y0 + 1) * ((x1 - x0 + 1 + 6) & ~3)


def simulate(wins, cap):
    solo = sum(plane(*w[1:]) for w in wins)
    shared, groups = 0, 0
    for l in sorted({w[0] for w in wins}):
        ws = sorted((w[1:] for w in wins if w[0] == l), key=lambda t: (t[0] // 16, t[2]))
        cur = None
        for y0, y1, x0, x1 in ws:
            if cur is not None:
                u = (min(cur[0], y0), max(cur[1], y1), min(cur[2], x0), max(cur[3], x1))
                if plane(*u) <= cap:
                    cur = u
                    continue
                shared += plane(*cur)
                groups += 1
            cur = (y0, y1, x0, x1)
        if cur is not None:
            shared += plane(*cur)
            groups += 1
    return solo, shared, groups


def clustered_rois(rng, n, n_obj=25, h=800, w=1333):
    """Proposal-like rois: jittered copies of a few object boxes plus 20 % background noise."""
    cx, cy = rng.uniform(0, w, n_obj), rng.uniform(0, h, n_obj)
    bw, bh = rng.uniform(30, 500, n_obj), rng.uniform(30, 400, n_obj)
    k = rng.integers(0, n_obj, n)
    j = rng.normal(0, 0.12, (n, 4))
    x1 = cx[k] - bw[k] / 2 * (1 + j[:, 0])
    y1 = cy[k] - bh[k] / 2 * (1 + j[:, 1])
    x2 = cx[k] + bw[k] / 2 * (1 + j[:, 2])
    y2 = cy[k] + bh[k] / 2 * (1 + j[:, 3])
    r = np.stack([x1, y1, x2, y2], 1)
    noise = synth.random_rois(rng, 1, n)[0]
    m = rng.random(n) < 0.2
    r[m] = noise[m]
    return np.clip(r, 0, [w - 1, h - 1, w - 1, h - 1]).astype(np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1000)
    a = ap.parse_args()
    rng = np.random.default_rng(0)
    shapes, strides = synth.fpn_shapes(), synth.FPN_STRIDES
    for name, rois in (("uniform random (bench.py)", synth.random_rois(rng, 1, a.n)[0]),
                       ("clustered, proposal-like", clustered_rois(rng, a.n))):
        wins = windows(rois, shapes, strides)
        for cap in (768, 1536, 3072):
            solo, shared, groups = simulate(wins, cap)
            print(json.dumps({"rois": name, "n": a.n, "plane_cap_cells": cap, "staged_cells_per_channel_solo": solo,
                              "shared": shared, "groups": groups, "reduction": round(solo / shared, 2),
                              "rois_per_group": round(a.n / groups, 2)}))


if __name__ == "__main__":
    main()
