"""ProposalMaskTarget(output_ratio=True) at BASELINE config 4's size (2 images x 2000 rois, 100 gt with 1-3 polygon
segments, image_rois 512 -> 128 mask rows per image, 28x28): time of the plain operator vs the ratio variant
(CUDA events, median of 20).  python benchmarks/mask_ratio_time.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simpledet_b200 import ops, synth  # noqa: E402

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
B, R, G, PL, IR, M = 2, 2000, 100, 2500, 512, 28
rois, gt, polys = synth.mask_scene(rng, B, R, G, PL)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
rois, gt, polys = t(rois), t(gt), t(polys)
res = {}
for ratio in (False, True):
    fn = lambda: ops.ProposalMaskTarget(rois, gt, polys, 81, B, IR, M, 0.5, 0.5, 0.0, False, output_iou=True,
                                        output_ratio=ratio, seed=1)
    for _ in range(3):
        out = fn()
    ts = []
    for _ in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    res["ratio" if ratio else "plain"] = round(float(np.median(ts)), 1)
    if ratio:
        r = out[6]
        res["ratio_rows"] = int((r > 0).sum())
        res["ratio_min_max"] = [float(r[r > 0].min()), float(r.max())]
        res["nan"] = int(torch.isnan(r).sum())
print(json.dumps({"workload": "ProposalMaskTarget config 4 (2 x 2000 rois, 128 mask rows / image, 28x28)", "us": res}))
