"""Timeline of one RoIAlign forward CTA (profiling build only: SDET_NVCC_EXTRA=-DSDET_RA_ABLATE).
    python benchmarks/ra_trace.py --shape bench --cta 0
Arms the trace for CTA `--cta` of the main kernel, runs one launch and prints, per channel tile, when the
producer warp found the buffer free / finished issuing its copies and when consumer warp 0 saw the data /
finished computing (clock64 ticks since the CTA started, and the same in microseconds at 1.965 GHz)."""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simpledet_b200 import _lib, ops, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="bench")
    ap.add_argument("--cta", type=int, default=0)
    a = ap.parse_args()
    f = _lib.lib().sdet_debug_ra_trace  # AttributeError unless the library was built with -DSDET_RA_ABLATE
    f.argtypes, f.restype = [ctypes.c_int, ctypes.c_void_p], ctypes.c_int
    B, N, pooled = {"target": (1, 512, 14), "bench": (2, 1000, 7), "train": (2, 512, 7)}[a.shape]
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    feats = [torch.randn((B, 256, h, w), device=dev) for h, w in synth.fpn_shapes()]
    rois = torch.from_numpy(synth.random_rois(rng, B, N)).to(dev)
    for _ in range(3):
        ops.fpn_roi_align_raw(feats, rois, synth.FPN_STRIDES, pooled, with_argmax=False)
    f(a.cta, None)
    ops.fpn_roi_align_raw(feats, rois, synth.FPN_STRIDES, pooled, with_argmax=False)
    buf = (ctypes.c_longlong * (2 * 520))()
    f(-1, buf)
    cons, prod = buf[:520], buf[520:]
    ntiles = int(cons[1])
    us = lambda t: t / 1965.0  # noqa: E731
    print(f"preamble done at {cons[0]} ticks ({us(cons[0]):.2f} us); {ntiles} tiles")
    print(" tile | producer: buffer free, copies issued | consumer 0: data arrived, compute done   (us)")
    for t in range(min(ntiles, 256)):
        print(f"{t:5d} | {us(prod[2 + 2 * t]):8.2f} {us(prod[3 + 2 * t]):8.2f} | {us(cons[2 + 2 * t]):8.2f} {us(cons[3 + 2 * t]):8.2f}")


if __name__ == "__main__":
    main()
