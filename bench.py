#!/usr/bin/env python
"""bench.py — the driver's benchmark contract for the simpledet_b200 hot path.

  python bench.py --gpus N --steps K --warmup W            (our CUDA path)
  python bench.py --impl reference --gpus N --steps K ...  (the reference's CPU path = oracle)

One "step" = one pass of the detection-specific hot path of `faster_r50v1_fpn_1x` inference
(BASELINE.json configs[1], detection_infer_speed.py's graph + detection_test.py's NMS) over a
batch of synthetic 800x1333 images per GPU:

   5 x _contrib_Proposal_v3 (strides 4..64) -> get_top_proposal(1000) -> fused FPN RoIAlign_v2 7x7
   (1000 rois x 256 ch) -> _contrib_DecodeBBox (81 classes) -> per-class NMS (80 classes)

The backbone / RoI-head GEMMs are not on this path (tensor-core library work); their outputs
(FPN features, RPN maps, head logits/deltas) are the synthetic inputs.  Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STRIDES_RPN = (4, 8, 16, 32, 64)
STRIDES_ROI = (4, 8, 16, 32)
IMG_H, IMG_W = 800, 1333
C_FEAT, N_ROI, K_CLS, POOLED = 256, 1000, 81, 7
WORKLOAD = ("faster_r50v1_fpn_1x inference hot path, synthetic 800x1333: 5x Proposal_v3 -> get_top_proposal(1000)"
            " -> FPN RoIAlign_v2 7x7 (1000 rois x 256 ch) -> DecodeBBox(81) -> per-class NMS(80)")


def level_shapes(strides):
    return [(-(-IMG_H // s), -(-IMG_W // s)) for s in strides]


def make_inputs_np(rng, B):
    """One image set on the host (numpy, float32)."""
    d = {}
    for s, (h, w) in zip(STRIDES_RPN, level_shapes(STRIDES_RPN)):
        logit = rng.standard_normal((B, 3, h, w)).astype(np.float32) * 2 - 3
        fg = 1 / (1 + np.exp(-logit))
        d[f"cls_prob{s}"] = np.concatenate([1 - fg, fg], 1).astype(np.float32)
        d[f"bbox_pred{s}"] = (rng.standard_normal((B, 12, h, w)) * 0.3).astype(np.float32)
    for s, (h, w) in zip(STRIDES_ROI, level_shapes(STRIDES_ROI)):
        d[f"feat{s}"] = rng.standard_normal((B, C_FEAT, h, w)).astype(np.float32)
    d["im_info"] = np.tile(np.array([[IMG_H, IMG_W, 1.0]], np.float32), (B, 1))
    z = rng.standard_normal((B, N_ROI, K_CLS)).astype(np.float32) * 2
    z[..., 0] += 3  # mostly background, a few confident classes
    e = np.exp(z - z.max(-1, keepdims=True))
    d["cls_score"] = (e / e.sum(-1, keepdims=True)).astype(np.float32)
    d["head_bbox_pred"] = (rng.standard_normal((B, N_ROI, 4 * K_CLS)) * 0.5).astype(np.float32)
    return d


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def hot_path_step(ops, d, ev=None):
    """The hot path on device tensors `d`.  `ev` = (start, end) CUDA events recorded around the
    dominant kernel (RoIAlign) on the current stream."""
    boxes, scores = ops.Proposal_v3_fpn([d[f"cls_prob{s}"] for s in STRIDES_RPN],
                                        [d[f"bbox_pred{s}"] for s in STRIDES_RPN], d["im_info"], STRIDES_RPN,
                                        rpn_pre_nms_top_n=1000, rpn_post_nms_top_n=1000, threshold=0.7,
                                        rpn_min_size=0, scales=(8,), ratios=(0.5, 1.0, 2.0))
    rois, _ = ops.get_top_proposal(boxes, scores, N_ROI)
    feats = [d[f"feat{s}"] for s in STRIDES_ROI]
    if ev:
        ev[0].record()
    roi_feat = ops.fpn_roi_align_raw(feats, rois, STRIDES_ROI, POOLED, 224, 4, with_argmax=False)[0]
    if ev:
        ev[1].record()
    # (RoI head: 2 fc + cls/reg fc on tensor cores — library GEMMs, not on this path)
    bbox = ops.DecodeBBox(rois, d["head_bbox_pred"], d["im_info"], (0, 0, 0, 0), (0.1, 0.1, 0.2, 0.2),
                          class_agnostic=False)
    dets, counts, keep, nkeep, _ = ops.multiclass_nms(d["cls_score"], bbox, 0.5, 0.05, first_class=1)
    return rois, roi_feat, dets, counts, keep, nkeep


def roialign_algorithmic_bytes(rois_np, B):
    """sz(out) + sum_l min(sz(feat_l), sum of window bytes on l) + sz(rois)."""
    import oracle

    lv = oracle.fpn_assign_levels(rois_np, STRIDES_ROI).reshape(rois_np.shape[:2])
    total = B * N_ROI * C_FEAT * POOLED * POOLED * 4 + rois_np.size * 4
    for l, ((h, w), s) in enumerate(zip(level_shapes(STRIDES_ROI), STRIDES_ROI)):
        m = lv == l
        if not m.any():
            continue
        r = rois_np[m] / s
        x1 = np.clip(np.floor(r[:, 0]), 0, w - 1)
        x2 = np.clip(np.ceil(r[:, 2]), 0, w - 1)
        y1 = np.clip(np.floor(r[:, 1]), 0, h - 1)
        y2 = np.clip(np.ceil(r[:, 3]), 0, h - 1)
        total += min(float(((x2 - x1 + 1) * (y2 - y1 + 1)).sum()) * C_FEAT * 4, B * C_FEAT * h * w * 4)
    return int(total)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md): one
    `nvidia-smi -lms 20` child streams samples while the region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.idx, self.proc = [], gpu_index, None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.proc.stdout.readline()  # first sample = the child is up; region starts after it
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is None:
            return
        time.sleep(0.03)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 6 and f[0].replace(".", "").isdigit():
                self.rows.append(f)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def cpu_hot_path(d, n_images):
    """The reference's CPU path (oracle restatements) for `n_images` images of set `d`."""
    import oracle
    from oracle import np_ops

    for b in range(n_images):
        sl = slice(b, b + 1)
        boxes, scores = [], []
        for s in STRIDES_RPN:
            r, sc = oracle.proposal_v3(d[f"cls_prob{s}"][sl], d[f"bbox_pred{s}"][sl], d["im_info"][sl],
                                       feature_stride=s, scales=(8,), ratios=(0.5, 1, 2), rpn_pre_nms_top_n=1000,
                                       rpn_post_nms_top_n=1000, threshold=0.7, rpn_min_size=0)
            boxes.append(r)
            scores.append(sc)
        rois, _ = np_ops.get_top_proposal(np.concatenate(boxes, 1), np.concatenate(scores, 1), N_ROI)
        oracle.fpn_roi_align_v2_forward([d[f"feat{s}"][sl] for s in STRIDES_ROI], rois, STRIDES_ROI,
                                        (POOLED, POOLED))
        bbox = oracle.decode_bbox(rois, d["head_bbox_pred"][sl], d["im_info"][sl], (0, 0, 0, 0),
                                  (0.1, 0.1, 0.2, 0.2), False)
        np_ops.do_nms(d["cls_score"][b][:, 1:], bbox[0][:, 4:], 0.5, 0.05)


def time_cpu(d, budget_s=12.0):
    import oracle

    oracle.build()
    oracle.set_threads()
    t0 = time.perf_counter()
    cpu_hot_path(d, 1)
    one = time.perf_counter() - t0
    n = max(1, min(d["im_info"].shape[0] * 4, int(budget_s / max(one, 1e-3))))
    t0 = time.perf_counter()
    done = 0
    while done < n:
        m = min(d["im_info"].shape[0], n - done)
        cpu_hot_path(d, m)
        done += m
    dt = time.perf_counter() - t0
    return done / dt, done


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port: the
    operator_cxx sources need MXNet headers and cannot be compiled here; the Cython NMS/IoU that
    does compile is pinned against the port in tests).  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import oracle

    oracle.build()
    rng = np.random.default_rng(0)
    d = make_inputs_np(rng, 1)
    cores = oracle.set_threads() or 1  # (torchrun exports OMP_NUM_THREADS=1: undo it for the CPU arm)
    for _ in range(min(args.warmup, 1)):
        cpu_hot_path(d, 1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_hot_path(d, 1)  # one step = a bounded sample: 1 image of the workload
    dt = time.perf_counter() - t0
    v = args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": "images/sec", "value": round(v, 3), "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(1e3 * dt / args.steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "images_per_step": 1},
        "cpu_baseline": {"value": round(v, 3), "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": "1 image per step; RoIAlign/decode/proposal C restatement with OpenMP over "
                                   "all host cores, NMS numpy (single thread) as in the reference"},
        "e2e": {"value": round(v, 3), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_ours(args):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as g
    from simpledet_b200 import _lib, ops, shard

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        g.build()
    if world > 1:
        dist.barrier()
    _lib.lib()

    B, K, W = args.images_per_gpu, args.steps, max(args.warmup, 3)
    R = max(2, -(-300 // (B * 99)))  # rotating input sets: footprint R*B*99 MB >> 126 MB L2
    rng = np.random.default_rng(1234 + rank)
    host_sets = [make_inputs_np(rng, B) for _ in range(R)]
    dev_sets = [{k: torch.from_numpy(v).to(dev) for k, v in hs.items()} for hs in host_sets]
    pinned = [{k: torch.from_numpy(v).pin_memory() for k, v in hs.items()} for hs in host_sets]
    h2d_bytes = sum(v.numel() * 4 for v in pinned[0].values())

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput (`value`) ----
    for i in range(W):
        out = hot_path_step(ops, dev_sets[i % R])
    rois_np = out[0].cpu().numpy()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    n0 = _lib.launch_count()
    with ClockSampler(local) as clk:
        e0.record()
        for i in range(K):
            hot_path_step(ops, dev_sets[i % R], kev[i])
        e1.record()
        sync_all()
    launches = _lib.launch_count() - n0
    ms = shard.max_over_ranks(e0.elapsed_time(e1), dev)
    ra_us = float(np.mean([a.elapsed_time(b) for a, b in kev])) * 1e3

    # ---- end to end through the public API with HOST buffers (`e2e`) ----
    # Every step copies its inputs from pinned host memory and reads its detections back into pinned
    # host memory; all of it is inside the timed region.  Steps are independent images, so the copy of
    # step i+1 runs on a second stream while step i computes (two device input slots, event-ordered).
    main = torch.cuda.current_stream()
    copy_s, back_s = torch.cuda.Stream(), torch.cuda.Stream()
    slots = [{k: torch.empty_like(v, device=dev) for k, v in pinned[0].items()} for _ in range(2)]
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    host_out = [None, None]
    done_ev = [torch.cuda.Event() for _ in range(2)]

    def e2e_upload(i):
        j = i % 2
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(consumed[j])  # the step that last read this slot has finished
            for k, v in pinned[i % R].items():
                slots[j][k].copy_(v, non_blocking=True)
            copied[j].record(copy_s)

    def e2e_compute(i):
        j = i % 2
        main.wait_event(copied[j])
        rois, roi_feat, dets, counts, keep, nkeep = hot_path_step(ops, slots[j])
        consumed[j].record(main)
        res = (dets, counts, keep, nkeep)
        if host_out[j] is None:
            host_out[j] = [torch.empty(x.shape, dtype=x.dtype).pin_memory() for x in res]
        done_ev[j].record(main)
        with torch.cuda.stream(back_s):
            back_s.wait_event(done_ev[j])
            for h, x in zip(host_out[j], res):
                x.record_stream(back_s)
                h.copy_(x, non_blocking=True)
        return host_out[j]

    def e2e_run(n):
        for ev in consumed:
            ev.record(main)
        e2e_upload(0)
        for i in range(n):
            if i + 1 < n:
                e2e_upload(i + 1)
            res_ = e2e_compute(i)
        main.wait_stream(back_s)
        main.wait_stream(copy_s)
        return res_

    res = e2e_run(3)
    d2h_bytes = sum(x.numel() * x.element_size() for x in res)
    sync_all()
    with ClockSampler(local) as clk2:
        e0.record()
        e2e_run(K)
        e1.record()
        sync_all()
    clk.rows += clk2.rows  # clocks are reported over both timed regions
    e2e_ms = shard.max_over_ranks(e0.elapsed_time(e1), dev)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = world * B * K / (ms / 1e3)
    e2e_value = world * B * K / (e2e_ms / 1e3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg = roialign_algorithmic_bytes(rois_np, B)
    achieved = alg / (ra_us * 1e-6) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get("fpn_roi_align_infer")
    except Exception:
        pass
    out = {
        "metric": "images/sec", "value": round(value, 2), "unit": "images/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": round(ms / K, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "images_per_gpu_per_step": B, "global_images_per_step": B * world,
                   "parallelism": f"dp{world} (sharded by image, no data-path collective)",
                   "l2": f"{R} rotating input sets ({R * B * 99} MB) larger than L2, no flush",
                   "im_info": [IMG_H, IMG_W, 1.0], "weights": "random synthetic activations (seeded)"},
        "clocks": clk.summary(),
        "e2e": {"value": round(e2e_value, 2), "unit": "images/s", "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": d2h_bytes, "ms_per_step": round(e2e_ms / K, 4)},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "roi_align_v2_fwd_kernel (fused FPN RoIAlign 7x7, %d rois x 256 ch)" % (B * N_ROI),
                     "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": traffic,
                     "algorithmic_bytes_per_launch": alg, "us_per_launch": round(ra_us, 2),
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6.65 TB/s (of fallback)",
                     "share_of_step": round(ra_us / (1e3 * ms / K), 4)},
    }
    if world == 1 and not args.no_cpu_baseline:
        v, n = time_cpu(host_sets[0])
        out["cpu_baseline"] = {"value": round(v, 3), "unit": "images/s", "cores": os.cpu_count() or 1,
                               "kind": "port",
                               "sample": f"{n} image(s) of the same workload; C restatement + OpenMP over all host "
                                         "cores for RoIAlign, single thread elsewhere (as the reference)"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images-per-gpu", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
