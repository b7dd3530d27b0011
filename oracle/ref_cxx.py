"""TEST INFRASTRUCTURE: ctypes access to oracle/_ref/libref_cxx.so — the reference's own operator_cxx
sources compiled unmodified against oracle/shim (oracle/build_ref_cxx.py).  Only tests/ and the golden
generator may import this; it needs either /root/reference (to build) or the prebuilt .so."""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref_cxx.so")
_lib = None

K_NULL, K_WRITE, K_INPLACE, K_ADD = 0, 1, 2, 3


def available() -> bool:
    if os.path.exists(LIB_PATH):
        return True
    if os.path.isdir("/root/reference/operator_cxx"):
        import subprocess
        import sys

        return subprocess.run([sys.executable, os.path.join(_HERE, "build_ref_cxx.py")]).returncode == 0 and \
            os.path.exists(LIB_PATH)
    return False


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not available():
            raise ImportError("oracle/_ref/libref_cxx.so missing and /root/reference not present")
        L = ctypes.CDLL(LIB_PATH)
        L.ref_last_error.restype = ctypes.c_char_p
        L.ref_rand_calls.restype = ctypes.c_longlong
        _lib = L
    return _lib


def _kw(kwargs: dict) -> bytes:
    def fmt(v):
        if isinstance(v, bool):
            return "True" if v else "False"
        if isinstance(v, float):
            return repr(float(np.float32(v))) if False else "%.17g" % v
        if isinstance(v, (tuple, list)):
            return "(" + ",".join(fmt(x) for x in v) + ")"
        return str(v)

    return "|".join(f"{k}={fmt(v)}" for k, v in kwargs.items()).encode()


def _shape_rows(arrs):
    nd = (ctypes.c_int * max(1, len(arrs)))(*[a.ndim for a in arrs])
    dims = np.zeros((max(1, len(arrs)), 8), np.int64)
    for i, a in enumerate(arrs):
        dims[i, :a.ndim] = a.shape
    return nd, dims


def _check(rc):
    if rc != 0:
        raise RuntimeError("reference op failed: " + lib().ref_last_error().decode())


def infer_shape(op: str, kwargs: dict, in_shapes):
    """-> (list of output shapes, number of visible outputs), by the operator's own InferShape / FInferShape."""
    L = lib()
    n = len(in_shapes)
    nd = (ctypes.c_int * n)(*[len(s) for s in in_shapes])
    dims = np.zeros((n, 8), np.int64)
    for i, s in enumerate(in_shapes):
        dims[i, :len(s)] = s
    n_out, n_vis = ctypes.c_int(0), ctypes.c_int(0)
    ond = (ctypes.c_int * 8)()
    odims = np.zeros((8, 8), np.int64)
    _check(L.ref_infer_shape(op.encode(), _kw(kwargs), n, nd, dims.ctypes.data_as(ctypes.c_void_p),
                             ctypes.byref(n_out), ond, odims.ctypes.data_as(ctypes.c_void_p), ctypes.byref(n_vis)))
    return [tuple(int(x) for x in odims[i, :ond[i]]) for i in range(n_out.value)], n_vis.value


def forward(op: str, kwargs: dict, inputs, out_shapes=None, dev: str = "cpu", reqs=None, outputs=None):
    """Run the reference operator.  inputs: float32 arrays.  out_shapes default to the operator's own shape
    inference; `outputs` may supply pre-filled arrays (kAddTo).  Returns every output, hidden ones included."""
    L = lib()
    ins = [np.ascontiguousarray(a, np.float32) for a in inputs]
    if outputs is None:
        if out_shapes is None:
            out_shapes, _ = infer_shape(op, kwargs, [a.shape for a in ins])
        outputs = [np.zeros(s, np.float32) for s in out_shapes]
    reqs = reqs or [K_WRITE] * len(outputs)
    ind, idims = _shape_rows(ins)
    ond, odims = _shape_rows(outputs)
    iptr = (ctypes.c_void_p * max(1, len(ins)))(*[a.ctypes.data for a in ins])
    optr = (ctypes.c_void_p * max(1, len(outputs)))(*[a.ctypes.data for a in outputs])
    _check(L.ref_forward(op.encode(), _kw(kwargs), dev.encode(), len(ins), iptr, ind,
                         idims.ctypes.data_as(ctypes.c_void_p), len(outputs), optr, ond,
                         odims.ctypes.data_as(ctypes.c_void_p), (ctypes.c_int * len(outputs))(*reqs)))
    return outputs


def backward(op: str, kwargs: dict, out_grads, inputs, outputs, reqs=None, dev: str = "cpu"):
    """Backward of a legacy OperatorProperty operator (FocalLoss, BBoxNorm): -> list of in_grad arrays, one per
    input, written by the operator's own Backward."""
    L = lib()
    ogs = [np.ascontiguousarray(a, np.float32) for a in out_grads]
    ins = [np.ascontiguousarray(a, np.float32) for a in inputs]
    outs = [np.ascontiguousarray(a, np.float32) for a in outputs]
    igs = [np.zeros(a.shape, np.float32) for a in ins]
    reqs = reqs or [K_WRITE] * len(ins)

    def ptrs(arrs):
        return (ctypes.c_void_p * max(1, len(arrs)))(*[a.ctypes.data for a in arrs])

    gnd, gdims = _shape_rows(ogs)
    ind, idims = _shape_rows(ins)
    ond, odims = _shape_rows(outs)
    _check(L.ref_backward(op.encode(), _kw(kwargs), dev.encode(), len(ogs), ptrs(ogs), gnd, gdims.ctypes.data_as(ctypes.c_void_p),
                          len(ins), ptrs(ins), ind, idims.ctypes.data_as(ctypes.c_void_p), len(outs), ptrs(outs), ond,
                          odims.ctypes.data_as(ctypes.c_void_p), ptrs(igs), (ctypes.c_int * len(ins))(*reqs)))
    return igs


def set_rand_const(v: int):
    """Every rand() the reference's std::random_shuffle draws returns v.  v = 0: each shuffle rotates its list
    right by one (libstdc++ Fisher-Yates, j = rand() % (i+1)); v = 27719 = lcm(1..12) - 1: identity for lists
    of up to 12 entries."""
    lib().ref_set_rand_const(int(v))


def set_rand_seq(vals):
    a = (ctypes.c_int * len(vals))(*[int(v) for v in vals])
    lib().ref_set_rand_seq(a, len(vals))


def rand_calls() -> int:
    return int(lib().ref_rand_calls())
