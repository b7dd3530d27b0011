"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's hot-path operators (C in ``oracle/*.c``, numpy in
``oracle/np_ops.py``).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package; ``simpledet_b200`` never
does.  See the header of each source file for the reference file:line it follows and for its
parity-pin status.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/*.c -> oracle/_build/liboracle.so (gcc, -ffp-contract=off)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith(".c")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "_build/liboracle.so"], check=True,
                       capture_output=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def roi_align_v2_forward(data, rois, pooled_size, spatial_scale, with_argmax=True):
    """roi_align_v2-inl.h:61-153.  data (B,C,H,W), rois (B,N,4) -> out[, argmax_x, argmax_y]."""
    data, rois = _f32(data), _f32(rois)
    B, C, H, W = data.shape
    assert rois.shape[0] == B and rois.shape[2] == 4
    N = rois.shape[1]
    ph, pw = pooled_size
    out = np.empty((B, N, C, ph, pw), np.float32)
    ax = np.empty_like(out) if with_argmax else None
    ay = np.empty_like(out) if with_argmax else None
    lib().oracle_roi_align_v2_forward(_p(data), _p(rois), B, N, C, H, W, ph, pw,
                                      ctypes.c_float(spatial_scale), _p(out), _p(ax), _p(ay))
    return (out, ax, ay) if with_argmax else out


def roi_align_v2_backward(ograd, argmax_x, argmax_y, data_shape, accumulate_into=None):
    """roi_align_v2.cu:35-84 (GPU semantics), adds applied in index order."""
    ograd, argmax_x, argmax_y = _f32(ograd), _f32(argmax_x), _f32(argmax_y)
    B, N, C, ph, pw = ograd.shape
    _, _, H, W = data_shape
    if accumulate_into is None:
        grad = np.empty(data_shape, np.float32)
        acc = 0
    else:
        grad = _f32(accumulate_into).copy()
        acc = 1
    lib().oracle_roi_align_v2_backward(_p(ograd), _p(argmax_x), _p(argmax_y), B, N, C, H, W, ph, pw,
                                       acc, _p(grad))
    return grad


def roi_pool_v1_forward(data, rois, pooled_size, spatial_scale):
    """roi_pooling_v1.cu:49-113.  rois (R,5) -> out, max_idx (R,C,PH,PW)."""
    data, rois = _f32(data), _f32(rois)
    B, C, H, W = data.shape
    R = rois.shape[0]
    ph, pw = pooled_size
    out = np.empty((R, C, ph, pw), np.float32)
    idx = np.empty_like(out)
    lib().oracle_roi_pool_v1_forward(_p(data), _p(rois), R, C, H, W, ph, pw,
                                     ctypes.c_float(spatial_scale), _p(out), _p(idx))
    return out, idx


def roi_pool_v1_backward(ograd, max_idx, rois, data_shape, accumulate_into=None):
    """roi_pooling_v1.cu:116-152."""
    ograd, max_idx, rois = _f32(ograd), _f32(max_idx), _f32(rois)
    R, C, ph, pw = ograd.shape
    B, _, H, W = data_shape
    if accumulate_into is None:
        grad = np.empty(data_shape, np.float32)
        acc = 0
    else:
        grad = _f32(accumulate_into).copy()
        acc = 1
    lib().oracle_roi_pool_v1_backward(_p(ograd), _p(max_idx), _p(rois), R, B, C, H, W, ph, pw, acc,
                                      _p(grad))
    return grad


def fpn_assign_levels(rois, strides, roi_canonical_scale=224, roi_canonical_level=4):
    """assign_layer_fpn.py:17-40 -> per-roi index into `strides` (-1: no level matches)."""
    rois = _f32(rois).reshape(-1, 4)
    lv = np.empty(rois.shape[0], np.int32)
    k_min, k_max = float(np.log2(min(strides))), float(np.log2(max(strides)))
    lib().oracle_fpn_assign_levels(_p(rois), ctypes.c_long(rois.shape[0]),
                                   ctypes.c_float(roi_canonical_scale),
                                   ctypes.c_float(roi_canonical_level), ctypes.c_float(k_min),
                                   ctypes.c_float(k_max), _p(lv))
    idx = np.full_like(lv, -1)
    for i, s in enumerate(strides):
        idx[(lv >= 0) & ((2 ** np.maximum(lv, 0)) == s) & (lv >= 0)] = i
    return idx


def fpn_roi_align_v2_forward(feats, rois, strides, pooled_size, roi_canonical_scale=224,
                             roi_canonical_level=4):
    """The reference graph models/FPN/builder.py:573-605 literally: assign, zero the roi on the
    other levels, run ROIAlign_v2 on every level with all rois, add_n."""
    rois = _f32(rois)
    idx = fpn_assign_levels(rois, strides, roi_canonical_scale, roi_canonical_level).reshape(
        rois.shape[:2])
    total = None
    for i, (f, s) in enumerate(zip(feats, strides)):
        lvl_rois = np.where((idx == i)[..., None], rois, np.float32(0))
        o = roi_align_v2_forward(f, lvl_rois, pooled_size, 1.0 / s, with_argmax=False)
        total = o if total is None else total + o
    return total, idx
