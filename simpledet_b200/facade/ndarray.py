"""`mx.nd` / `mx.io` stand-ins: an NDArray that wraps a torch tensor, DataBatch / DataDesc, CustomOp bases."""
from __future__ import annotations

from collections import namedtuple

import numpy as np
import torch


class NDArray:
    def __init__(self, t, ctx=None):
        self.t = t if isinstance(t, torch.Tensor) else torch.as_tensor(np.asarray(t, dtype=np.float32))
        self._ctx = ctx

    shape = property(lambda self: tuple(self.t.shape))
    dtype = property(lambda self: np.dtype(str(self.t.dtype).replace("torch.", "")).type)
    stype = "default"
    size = property(lambda self: self.t.numel())
    ndim = property(lambda self: self.t.dim())

    @property
    def context(self):
        from . import Context

        return self._ctx or (Context("gpu", self.t.device.index or 0) if self.t.is_cuda else Context("cpu", 0))

    ctx = context

    def asnumpy(self):
        return self.t.detach().cpu().numpy()

    def asscalar(self):
        return self.asnumpy().reshape(-1)[0]

    def wait_to_read(self):
        if self.t.is_cuda:
            torch.cuda.synchronize(self.t.device)

    def reshape(self, *shape, **kw):
        shape = shape[0] if len(shape) == 1 and isinstance(shape[0], (tuple, list)) else shape
        return NDArray(self.t.reshape(tuple(shape)), self._ctx)

    def astype(self, dtype):
        return NDArray(self.t.to(getattr(torch, np.dtype(dtype).name)), self._ctx)

    def copyto(self, other):
        if isinstance(other, NDArray):
            other.t.copy_(self.t)
            return other
        return as_in_context(self, other)

    def as_in_context(self, ctx):
        return as_in_context(self, ctx)

    def copy(self):
        return NDArray(self.t.clone(), self._ctx)

    def __getitem__(self, k):
        return NDArray(self.t[k], self._ctx)

    def __setitem__(self, k, v):
        self.t[k] = v.t if isinstance(v, NDArray) else v

    def __len__(self):
        return self.t.shape[0]

    def __repr__(self):
        return f"<NDArray {self.shape} @{self.context}>"


def _dev(ctx):
    if ctx is None or ctx.device_type == "cpu":
        return torch.device("cpu")
    return torch.device("cuda", ctx.device_id)


def as_in_context(a, ctx):
    return NDArray(a.t.to(_dev(ctx)), ctx)


def array(src, ctx=None, dtype=None):
    a = np.asarray(src.asnumpy() if isinstance(src, NDArray) else src, dtype=np.float32 if dtype is None else dtype)
    return NDArray(torch.from_numpy(np.ascontiguousarray(a)).to(_dev(ctx)), ctx)


def _shape(shape, kw):
    shape = kw.get("shape", shape)
    return tuple(shape) if isinstance(shape, (tuple, list)) else (int(shape),)


def zeros(shape=None, ctx=None, dtype=None, stype=None, **kw):
    return NDArray(torch.zeros(_shape(shape, kw), device=_dev(ctx)), ctx)


def ones(shape=None, ctx=None, dtype=None, **kw):
    return NDArray(torch.ones(_shape(shape, kw), device=_dev(ctx)), ctx)


def empty(shape=None, ctx=None, dtype=None, **kw):
    return zeros(shape, ctx, dtype, **kw)


def exports():
    return dict(NDArray=NDArray, array=array, zeros=zeros, ones=ones, empty=empty, waitall=lambda: torch.cuda.synchronize()
                if torch.cuda.is_available() else None)


class DataDesc(namedtuple("DataDesc", ["name", "shape"])):
    def __new__(cls, name, shape, dtype=np.float32, layout="NCHW"):
        ret = super().__new__(cls, name, tuple(shape))
        ret.dtype, ret.layout = dtype, layout
        return ret


class DataBatch:
    def __init__(self, data, label=None, pad=None, index=None, bucket_key=None, provide_data=None, provide_label=None):
        self.data, self.label, self.pad, self.index = data, label, pad, index
        self.bucket_key, self.provide_data, self.provide_label = bucket_key, provide_data, provide_label


class InitDesc(str):
    def __new__(cls, name, attrs=None, global_init=None):
        ret = super().__new__(cls, name)
        ret.attrs, ret.global_init = attrs or {}, global_init
        return ret


class CustomOp:
    def assign(self, dst, req, src):
        if req in ("null",):
            return
        dst[:] = src


class CustomOpProp:
    def __init__(self, need_top_grad=False):
        self.need_top_grad_ = need_top_grad


_CUSTOM = {}


def register_custom(name):
    def deco(cls):
        _CUSTOM[name] = cls
        return cls

    return deco
