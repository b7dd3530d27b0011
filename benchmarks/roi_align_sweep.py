"""Kernel-level timing of the RoIAlign kernels on one GPU (tuning tool, not the bench contract).

  python benchmarks/roi_align_sweep.py [--shape target|infer|train] [--iters 20]

Prints one JSON line per configuration: device time (CUDA events, L2 flushed between
iterations), algorithmic bytes and achieved GB/s."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simpledet_b200 import ops, synth  # noqa: E402


def algorithmic_bytes(rois, levels, shapes, C, pooled, strides, with_argmax):
    """bytes = sz(out) [x3 with argmax] + sum_l min(sz(feat_l), sum of window bytes on l) + sz(rois)."""
    B, N = rois.shape[:2]
    out = B * N * C * pooled * pooled * 4
    total = out * (3 if with_argmax else 1) + rois.size * 4
    for l, ((h, w), s) in enumerate(zip(shapes, strides)):
        m = levels == l
        if not m.any():
            continue
        r = rois[m] / s
        x1 = np.clip(np.floor(r[:, 0]), 0, w - 1)
        x2 = np.clip(np.ceil(r[:, 2]), 0, w - 1)
        y1 = np.clip(np.floor(r[:, 1]), 0, h - 1)
        y2 = np.clip(np.ceil(r[:, 3]), 0, h - 1)
        win = float(((x2 - x1 + 1) * (y2 - y1 + 1)).sum()) * C * 4
        total += min(win, B * C * h * w * 4)
    return int(total)


def time_op(fn, iters, flush, do_flush=True):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(iters)]
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    for s, e in ev:
        if do_flush:
            flush.fill_(1.0)  # 512 MB write: evicts the 126 MB L2
        s.record()
        fn()
        e.record()
    torch.cuda.synchronize()
    ts = sorted(s.elapsed_time(e) for s, e in ev)
    return ts[len(ts) // 2] * 1e3, ts[0] * 1e3  # median, min (us)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="target")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--argmax", type=int, default=0)
    ap.add_argument("--plan", type=int, default=1)
    ap.add_argument("--noflush", type=int, default=0)
    ap.add_argument("--sorted", type=int, default=0, help="1: rois pre-sorted by (level, y, x) on the host (locality probe)")
    ap.add_argument("--path", type=int, default=0, help="0: automatic, 1: per-roi kernel, 2: band-stationary, 3: channels-last "
                    "(NCHW in, re-layout inside the call), 4: channels-last with NHWC features given (no re-layout)")
    ap.add_argument("--same-roi", type=int, default=0, help="1: every roi is a copy of the first (cache-hit probe: compute-bound rate)")
    ap.add_argument("--backward", type=int, default=0, help="time the backward pass (incl. zero-fill of the grads)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    B, N, pooled = {"target": (1, 512, 14), "infer": (1, 1000, 7), "train": (2, 512, 7),
                    "mask": (2, 128, 14), "bench": (2, 1000, 7)}[a.shape]
    C = 256
    shapes = synth.fpn_shapes()
    feats = [torch.randn((B, C, h, w), device=dev) for h, w in shapes]
    rois_np = synth.random_rois(rng, B, N)
    if a.sorted:
        for b in range(B):
            r = rois_np[b]
            wh = np.sqrt((r[:, 2] - r[:, 0] + 1) * (r[:, 3] - r[:, 1] + 1))
            lvl = np.clip(np.floor(4 + np.log2(wh / 224 + 1e-6)), 2, 5)
            if a.sorted == 1:
                rois_np[b] = r[np.lexsort((r[:, 0], r[:, 1] // 32, lvl))]
            else:  # 2: largest window first (longest-processing-time order: a scheduling probe)
                s_ = 2.0 ** (lvl + 2)
                cells = (np.ceil(r[:, 2] / s_) - np.floor(r[:, 0] / s_) + 2) * (np.ceil(r[:, 3] / s_) - np.floor(r[:, 1] / s_) + 2)
                rois_np[b] = r[np.argsort(-cells, kind="stable")]
    if a.same_roi:
        rois_np[:] = np.array([400, 300, 624, 524], np.float32)  # 224 px: level P4, 14 x 14 cells
    rois = torch.from_numpy(rois_np).to(dev)
    flush = torch.empty(128 * 1024 * 1024, device=dev)
    out, _, _, lv = ops.fpn_roi_align_raw(feats, rois, synth.FPN_STRIDES, pooled, with_argmax=False)
    lv = lv.cpu().numpy()
    nbytes = algorithmic_bytes(rois_np, lv, shapes, C, pooled, synth.FPN_STRIDES, bool(a.argmax))

    feats_cl = [f.permute(0, 2, 3, 1).contiguous() for f in feats] if a.path == 4 else None

    def fn():
        if a.path == 4:
            ops.fpn_roi_align_nhwc(feats_cl, rois, synth.FPN_STRIDES, pooled)
        else:
            ops.fpn_roi_align_raw(feats, rois, synth.FPN_STRIDES, pooled, with_argmax=bool(a.argmax),
                                  use_plan=bool(a.plan), path=a.path)

    if a.backward:
        import ctypes
        from simpledet_b200 import _lib
        out, ax, ay, lvt = ops.fpn_roi_align_raw(feats, rois, synth.FPN_STRIDES, pooled, with_argmax=True)
        ograd = torch.randn_like(out)
        grads = [torch.empty_like(f) for f in feats]
        L = len(grads)
        ptrs = (ctypes.c_void_p * L)(*[g.data_ptr() for g in grads])
        Hs = (ctypes.c_int * L)(*[h for h, w in shapes])
        Ws = (ctypes.c_int * L)(*[w for h, w in shapes])
        lib = _lib.lib()
        # algorithmic: read ograd + both argmax planes, zero-fill every level, 4 read-modify-writes per element
        nbytes = 3 * out.numel() * 4 + sum(g.numel() * 4 for g in grads)

        def fn():  # noqa: F811
            _lib.check(lib.sdet_fpn_roi_align_v2_backward(
                ograd.data_ptr(), ax.data_ptr(), ay.data_ptr(), lvt.data_ptr(), ptrs, Hs, Ws, L, B, N, C, pooled,
                pooled, 0, torch.cuda.current_stream().cuda_stream))

    med, mn = time_op(fn, a.iters, flush, not a.noflush)
    print(json.dumps({"shape": a.shape, "path": a.path, "B": B, "N": N, "pooled": pooled, "argmax": a.argmax, "sorted": a.sorted, "backward": a.backward, "plan": a.plan, "noflush": a.noflush, "us_median": round(med, 2), "us_min": round(mn, 2), "alg_bytes": nbytes,
                      "GBps": round(nbytes / med / 1e3, 1),
                      "levels": np.bincount(lv.ravel() + 1, minlength=5).tolist()}))


if __name__ == "__main__":
    main()
