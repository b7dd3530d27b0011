"""Host-side mirror of the reference's operator interface for the detection hot path.

Every public callable keeps the reference operator's name, argument names, defaults, output
arity and shape rules and forwards to the C ABI (include/simpledet_b200.h).
PyTorch is used for device memory, streams and autograd plumbing only — all arithmetic happens in
libsimpledet_b200.so.  There is no CPU path: CPU tensors raise.

`OPS` maps the reference's registry strings (``_contrib_ROIAlign_v2``, ``ROIPooling_v1`` …) to
these callables, which is what a symbol/builder façade binds to.
"""
from __future__ import annotations

import ctypes
from typing import Sequence

import torch

from . import _lib
from ._lib import check

__all__ = ["ROIAlign_v2", "roi_align_v2_raw", "ROIPooling_v1", "roi_pooling_v1_raw",
           "fpn_roi_align", "fpn_roi_align_raw", "OPS"]


def _dev(t: torch.Tensor | None, name: str, dtype=torch.float32) -> torch.Tensor | None:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"{name}: simpledet_b200 ops are CUDA-only (no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype} (the reference path is fp32, "
                        "symbol/builder.py:882-894)")
    return t.contiguous()


def _p(t: torch.Tensor | None):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _pair(v) -> tuple[int, int]:
    if isinstance(v, int):
        return v, v
    v = tuple(int(x) for x in v)
    if len(v) != 2:
        raise ValueError("pooled_size must have 2 dims (set_expect_ndim(2))")
    return v


# --------------------------------------------------------------------------------------------
# _contrib_ROIAlign_v2  (operator_cxx/contrib/roi_align_v2.cc:170-228)
# --------------------------------------------------------------------------------------------
def roi_align_v2_raw(data, rois, pooled_size, spatial_scale, with_argmax=True):
    """All three outputs of the reference op: (out, argmax_x, argmax_y), each (B,N,C,PH,PW).

    Shape rules of ROIAlign_v2 FInferShape (roi_align_v2.cc:187-210): data 4-D, rois 3-D
    (B,N,4) with rois.shape[0] == data.shape[0].
    """
    data, rois = _dev(data, "data"), _dev(rois, "rois")
    if data.dim() != 4:
        raise ValueError("data should be a 4D tensor")
    if rois.dim() != 3 or rois.shape[2] != 4:
        raise ValueError("bbox should be a 3D tensor of shape [batch, rois, 4]")
    if rois.shape[0] != data.shape[0]:
        raise ValueError("rois.shape[0] must equal data.shape[0]")
    ph, pw = _pair(pooled_size)
    B, C, H, W = data.shape
    N = rois.shape[1]
    out = torch.empty((B, N, C, ph, pw), device=data.device, dtype=torch.float32)
    ax = torch.empty_like(out) if with_argmax else None
    ay = torch.empty_like(out) if with_argmax else None
    check(_lib.lib().sdet_roi_align_v2_forward(_p(data), _p(rois), _p(out), _p(ax), _p(ay), B, N, C,
                                               H, W, ph, pw, float(spatial_scale), _stream()))
    return out, ax, ay


class _ROIAlignV2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, data, rois, ph, pw, spatial_scale):
        need = data.requires_grad
        out, ax, ay = roi_align_v2_raw(data, rois, (ph, pw), spatial_scale, with_argmax=need)
        if need:
            ctx.save_for_backward(ax, ay)
            ctx.dshape = tuple(data.shape)
            ctx.rshape = tuple(rois.shape)
        return out

    @staticmethod
    def backward(ctx, ograd):
        ax, ay = ctx.saved_tensors
        B, C, H, W = ctx.dshape
        N = ctx.rshape[1]
        ograd = _dev(ograd, "ograd")
        grad = torch.empty(ctx.dshape, device=ograd.device, dtype=torch.float32)
        ph, pw = ograd.shape[3], ograd.shape[4]
        check(_lib.lib().sdet_roi_align_v2_backward(_p(ograd), _p(ax), _p(ay), _p(grad), None, B, N,
                                                    C, H, W, ph, pw, 0, _stream()))
        # grad_rois is identically zero in the reference (roi_align_v2.cu:139-141)
        return grad, None, None, None, None


def ROIAlign_v2(data, rois, pooled_size, spatial_scale):
    """mx.sym.contrib.ROIAlign_v2(data, rois, pooled_size=(h,w), spatial_scale=s) — one visible
    output (B,N,C,PH,PW); differentiable w.r.t. data."""
    ph, pw = _pair(pooled_size)
    return _ROIAlignV2Fn.apply(data, rois, ph, pw, float(spatial_scale))


# --------------------------------------------------------------------------------------------
# Fused FPN RoIAlign = fpn_roi_assign + 4 x ROIAlign_v2 + add_n (models/FPN/builder.py:573-605)
# --------------------------------------------------------------------------------------------
def _level_arrays(feats: Sequence[torch.Tensor], strides: Sequence[int]):
    L = len(feats)
    if L != len(strides):
        raise ValueError("one stride per feature level")
    feats = [_dev(f, f"feat[{i}]") for i, f in enumerate(feats)]
    B, C = feats[0].shape[:2]
    for f in feats:
        if f.dim() != 4 or f.shape[0] != B or f.shape[1] != C:
            raise ValueError("all levels must be (B,C,H_l,W_l) with the same B and C")
    ptrs = (ctypes.c_void_p * L)(*[f.data_ptr() for f in feats])
    Hs = (ctypes.c_int * L)(*[f.shape[2] for f in feats])
    Ws = (ctypes.c_int * L)(*[f.shape[3] for f in feats])
    Ss = (ctypes.c_int * L)(*[int(s) for s in strides])
    return feats, ptrs, Hs, Ws, Ss, B, C


def fpn_roi_align_raw(feats, rois, strides, out_size, roi_canonical_scale=224,
                      roi_canonical_level=4, with_argmax=True):
    """-> (out, argmax_x, argmax_y, levels).  out (B,N,C,PH,PW); levels (B,N) int32 index into
    `strides` (-1: the roi matched no level and pools to zeros)."""
    feats, ptrs, Hs, Ws, Ss, B, C = _level_arrays(feats, strides)
    rois = _dev(rois, "rois")
    if rois.dim() != 3 or rois.shape[2] != 4 or rois.shape[0] != B:
        raise ValueError("rois must be (B,N,4)")
    ph, pw = _pair(out_size)
    N = rois.shape[1]
    out = torch.empty((B, N, C, ph, pw), device=rois.device, dtype=torch.float32)
    ax = torch.empty_like(out) if with_argmax else None
    ay = torch.empty_like(out) if with_argmax else None
    levels = torch.empty((B, N), device=rois.device, dtype=torch.int32)
    check(_lib.lib().sdet_fpn_roi_align_v2_forward(
        ptrs, Hs, Ws, Ss, len(feats), _p(rois), _p(out), _p(ax), _p(ay), _p(levels), B, N, C, ph, pw,
        int(roi_canonical_scale), int(roi_canonical_level), _stream()))
    return out, ax, ay, levels


class _FpnRoiAlignFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rois, strides, ph, pw, scale0, lvl0, *feats):
        need = any(f.requires_grad for f in feats)
        out, ax, ay, levels = fpn_roi_align_raw(feats, rois, strides, (ph, pw), scale0, lvl0,
                                                with_argmax=need)
        if need:
            ctx.save_for_backward(ax, ay, levels)
            ctx.shapes = [tuple(f.shape) for f in feats]
            ctx.N = rois.shape[1]
        return out

    @staticmethod
    def backward(ctx, ograd):
        ax, ay, levels = ctx.saved_tensors
        ograd = _dev(ograd, "ograd")
        grads = [torch.empty(s, device=ograd.device, dtype=torch.float32) for s in ctx.shapes]
        L = len(grads)
        B, C = ctx.shapes[0][:2]
        ptrs = (ctypes.c_void_p * L)(*[g.data_ptr() for g in grads])
        Hs = (ctypes.c_int * L)(*[s[2] for s in ctx.shapes])
        Ws = (ctypes.c_int * L)(*[s[3] for s in ctx.shapes])
        check(_lib.lib().sdet_fpn_roi_align_v2_backward(
            _p(ograd), _p(ax), _p(ay), _p(levels), ptrs, Hs, Ws, L, B, ctx.N, C, ograd.shape[3],
            ograd.shape[4], 0, _stream()))
        return (None, None, None, None, None, None, *grads)


def fpn_roi_align(feats, rois, strides=(4, 8, 16, 32), out_size=7, roi_canonical_scale=224,
                  roi_canonical_level=4):
    """FPNRoiAlign.get_roi_feature (models/FPN/builder.py:567-610) in one kernel.  Returns the
    5-D (B,N,C,PH,PW) tensor the reference reshapes with (-3,-2)."""
    ph, pw = _pair(out_size)
    return _FpnRoiAlignFn.apply(rois, tuple(int(s) for s in strides), ph, pw,
                                int(roi_canonical_scale), int(roi_canonical_level), *feats)


# --------------------------------------------------------------------------------------------
# ROIPooling_v1  (operator_cxx/roi_pooling_v1.cc:243; shape rules roi_pooling_v1-inl.h:170-195)
# --------------------------------------------------------------------------------------------
def roi_pooling_v1_raw(data, rois, pooled_size, spatial_scale):
    """-> (out, maxidx), each (R,C,PH,PW); rois (R,5) = [batch_index, x1, y1, x2, y2]."""
    data, rois = _dev(data, "data"), _dev(rois, "rois")
    if data.dim() != 4:
        raise ValueError("data should be a 4D tensor")
    if rois.dim() != 2 or rois.shape[1] != 5:
        raise ValueError("bbox should be a 2D tensor of shape [batch, 5]")
    ph, pw = _pair(pooled_size)
    B, C, H, W = data.shape
    R = rois.shape[0]
    out = torch.empty((R, C, ph, pw), device=data.device, dtype=torch.float32)
    idx = torch.empty_like(out)
    check(_lib.lib().sdet_roi_pooling_v1_forward(_p(data), _p(rois), _p(out), _p(idx), B, R, C, H, W,
                                                 ph, pw, float(spatial_scale), _stream()))
    return out, idx


class _ROIPoolingV1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, data, rois, ph, pw, spatial_scale):
        out, idx = roi_pooling_v1_raw(data, rois, (ph, pw), spatial_scale)
        ctx.save_for_backward(idx, rois)
        ctx.dshape = tuple(data.shape)
        return out

    @staticmethod
    def backward(ctx, ograd):
        idx, rois = ctx.saved_tensors
        B, C, H, W = ctx.dshape
        ograd = _dev(ograd, "ograd")
        R, _, ph, pw = ograd.shape
        grad = torch.empty(ctx.dshape, device=ograd.device, dtype=torch.float32)
        check(_lib.lib().sdet_roi_pooling_v1_backward(_p(ograd), _p(idx), _p(rois), _p(grad), None, B,
                                                      R, C, H, W, ph, pw, 0, _stream()))
        return grad, None, None, None, None


def ROIPooling_v1(data, rois, pooled_size, spatial_scale):
    """mx.sym.ROIPooling_v1(data, rois, pooled_size, spatial_scale) — one visible output."""
    ph, pw = _pair(pooled_size)
    return _ROIPoolingV1Fn.apply(data, rois, ph, pw, float(spatial_scale))


# Registry keyed by the reference's operator names (what symbol/builder.py binds by string).
OPS = {
    "_contrib_ROIAlign_v2": ROIAlign_v2,
    "ROIPooling_v1": ROIPooling_v1,
    "fpn_roi_align": fpn_roi_align,  # fusion of assign_layer_fpn + ROIAlign_v2 x L + add_n
}
