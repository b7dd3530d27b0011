"""Prints a digest of the fused FPN RoIAlign output for the sweep shapes; run once with and once without
SDET_RA_SHARE=1 (the switch is read once per process) — the digests must be identical."""
import hashlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simpledet_b200 import ops, synth  # noqa: E402

dev = torch.device("cuda:0")
for name, (B, N, pooled) in {"bench": (2, 1000, 7), "target": (1, 512, 14), "train": (2, 512, 7)}.items():
    rng = np.random.default_rng(7)
    feats = [torch.from_numpy(rng.standard_normal((B, 256, h, w), dtype=np.float32)).to(dev) for h, w in synth.fpn_shapes()]
    rois = torch.from_numpy(synth.random_rois(rng, B, N)).to(dev)
    out = ops.fpn_roi_align_raw(feats, rois, synth.FPN_STRIDES, pooled, with_argmax=False)[0]
    torch.cuda.synchronize()
    print(name, hashlib.sha1(out.cpu().numpy().tobytes()).hexdigest(), float(out.abs().sum()))
