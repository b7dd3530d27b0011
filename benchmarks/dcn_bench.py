"""Device timing of the DCNv1 sampling kernels at BASELINE config 5's shape: data (2,256,50,84),
3x3, num_deformable_group=4, offsets N(0,2).  Algorithmic bytes = sz(data) + sz(offset) + sz(col).

  python benchmarks/dcn_bench.py [--iters 20]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simpledet_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--shape", default="2,256,50,84")
    ap.add_argument("--dg", type=int, default=4)
    a = ap.parse_args()
    B, C, H, W = (int(x) for x in a.shape.split(","))
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    data = torch.randn((B, C, H, W), device=dev, generator=g)
    offset = 2.0 * torch.randn((B, a.dg * 18, H, W), device=dev, generator=g)
    col = torch.empty((B, C * 9, H * W), device=dev)
    gcol = torch.randn_like(col)
    gdata, goff = torch.empty_like(data), torch.empty_like(offset)
    flush = torch.empty(128 * 1024 * 1024, device=dev)
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    geom = (B, C, H, W, 3, 3, 1, 1, 1, 1, 1, 1, a.dg)

    def fwd():
        _lib.check(L.sdet_deformable_im2col(data.data_ptr(), offset.data_ptr(), col.data_ptr(), *geom, st))

    data_cl = data.permute(0, 2, 3, 1).contiguous()
    col_t = torch.empty((B, H * W, 9 * C), device=dev)

    def fwd_cl():
        _lib.check(L.sdet_deformable_im2col_nhwc(data_cl.data_ptr(), offset.data_ptr(), col_t.data_ptr(), *geom, st))

    def bwd():
        _lib.check(L.sdet_deformable_col2im(gcol.data_ptr(), data.data_ptr(), offset.data_ptr(), gdata.data_ptr(),
                                            goff.data_ptr(), *geom, st))

    def time_op(fn):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.iters)]
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        for s, e in ev:
            flush.fill_(1.0)  # evict the 126 MB L2
            s.record()
            fn()
            e.record()
        torch.cuda.synchronize()
        ts = sorted(s.elapsed_time(e) for s, e in ev)
        return ts[len(ts) // 2] * 1e3

    nbytes = 4 * (data.numel() + offset.numel() + col.numel())
    t_f = time_op(fwd)
    t_cl = time_op(fwd_cl)
    t_b = time_op(bwd)
    nb_b = 4 * (gcol.numel() + 2 * data.numel() + 2 * offset.numel())  # read gcol, data, offset; write both grads
    print(json.dumps({"shape": [B, C, H, W], "dg": a.dg, "im2col_us": round(t_f, 2), "im2col_alg_bytes": nbytes,
                      "im2col_GBps": round(nbytes / t_f / 1e3, 1), "im2col_nhwc_us": round(t_cl, 2),
                      "im2col_nhwc_GBps": round(nbytes / t_cl / 1e3, 1), "col2im_us": round(t_b, 2),
                      "col2im_alg_bytes": nb_b, "col2im_GBps": round(nb_b / t_b / 1e3, 1)}))


if __name__ == "__main__":
    main()
