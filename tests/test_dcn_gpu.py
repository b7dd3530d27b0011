"""DCNv1 sampling: CUDA gather vs the numpy restatement of the published formulation (parity
unpinned: the operator's source is not in the reference tree), the full op vs a torch reference
built on the same columns, and the scatter kernels vs torch autograd of a float64 re-implementation."""
import numpy as np
import pytest
import torch

from oracle import np_ops
from simpledet_b200 import _lib, ops
from simpledet_b200._lib import check

pytestmark = pytest.mark.gpu


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _im2col(data, offset, geo):
    kh, kw, ph, pw, sh, sw, dh, dw, dg = geo
    B, C, H, W = data.shape
    Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1
    Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1
    col = torch.empty((B, C * kh * kw, Ho * Wo), device=data.device)
    check(_lib.lib().sdet_deformable_im2col(data.data_ptr(), offset.data_ptr(), col.data_ptr(), B, C, H, W, kh, kw,
                                            ph, pw, sh, sw, dh, dw, dg, None))
    torch.cuda.synchronize()
    return col


@pytest.mark.parametrize("stride,dilate,pad,dg", [(1, 1, 1, 4), (2, 1, 1, 1), (1, 2, 2, 2)])
def test_im2col_matches_restatement(cuda, stride, dilate, pad, dg):
    rng = np.random.default_rng(stride * 10 + dilate)
    B, C, H, W = 2, 8, 13, 17
    data = rng.standard_normal((B, C, H, W)).astype(np.float32)
    Ho = (H + 2 * pad - (dilate * 2 + 1)) // stride + 1
    Wo = (W + 2 * pad - (dilate * 2 + 1)) // stride + 1
    offset = (rng.standard_normal((B, dg * 18, Ho, Wo)) * 2).astype(np.float32)
    offset[0, 0] = 0.0          # integer positions
    offset[0, 1] = 50.0         # far outside -> zeros
    ref = np_ops.deformable_im2col(data, offset, (3, 3), (stride, stride), (dilate, dilate), (pad, pad), dg)
    col = _im2col(_t(data, cuda), _t(offset, cuda), (3, 3, pad, pad, stride, stride, dilate, dilate, dg))
    np.testing.assert_allclose(col.cpu().numpy(), ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("stride,dilate,pad,dg,C", [(1, 1, 1, 4, 8), (2, 1, 1, 1, 6), (1, 2, 2, 2, 132), (1, 1, 1, 4, 256)])
def test_im2col_channels_last_is_the_same_columns(cuda, stride, dilate, pad, dg, C):
    """sdet_deformable_im2col_nhwc: col_t[b, p, t, c] == col[b, c*9 + t, p] bit for bit (same arithmetic, other layout),
    including pixels past the last full block of 32, clamped borders and samples outside the image."""
    rng = np.random.default_rng(stride * 100 + dilate * 10 + dg)
    B, H, W = 2, 13, 19
    data = rng.standard_normal((B, C, H, W)).astype(np.float32)
    Ho = (H + 2 * pad - (dilate * 2 + 1)) // stride + 1
    Wo = (W + 2 * pad - (dilate * 2 + 1)) // stride + 1
    offset = (rng.standard_normal((B, dg * 18, Ho, Wo)) * 3).astype(np.float32)
    offset[0, 0] = 0.0
    offset[0, 1] = 50.0
    d, o = _t(data, cuda), _t(offset, cuda)
    geo = (3, 3, pad, pad, stride, stride, dilate, dilate, dg)
    col = _im2col(d, o, geo)
    x = d.permute(0, 2, 3, 1).contiguous()
    col_t = torch.empty((B, Ho * Wo, 9, C), device=cuda)
    check(_lib.lib().sdet_deformable_im2col_nhwc(x.data_ptr(), o.data_ptr(), col_t.data_ptr(), B, C, H, W, 3, 3, pad, pad,
                                                 stride, stride, dilate, dilate, dg, None))
    torch.cuda.synchronize()
    assert torch.equal(col_t.permute(0, 3, 2, 1).reshape(B, C * 9, Ho * Wo), col)


def test_channels_last_input_gives_the_same_output(cuda):
    torch.manual_seed(3)
    x = torch.randn(2, 16, 20, 24, device=cuda)
    w = torch.randn(32, 16, 3, 3, device=cuda) * 0.1
    off = torch.randn(2, 4 * 18, 20, 24, device=cuda)
    kw = dict(kernel=(3, 3), pad=(1, 1), num_filter=32, num_deformable_group=4, no_bias=True)
    y0 = ops.DeformableConvolution(x, off, w, **kw)
    y1 = ops.DeformableConvolution(x.contiguous(memory_format=torch.channels_last), off, w, **kw)
    assert y0.shape == (2, 32, 20, 24) and torch.equal(y0, y1)


def test_zero_offset_equals_convolution(cuda):
    torch.manual_seed(0)
    x = torch.randn(2, 16, 20, 24, device=cuda)
    w = torch.randn(32, 16, 3, 3, device=cuda) * 0.1
    off = torch.zeros(2, 4 * 18, 20, 24, device=cuda)
    y = ops.DeformableConvolution(x, off, w, kernel=(3, 3), pad=(1, 1), num_filter=32, num_deformable_group=4,
                                  no_bias=True)
    torch.testing.assert_close(y, torch.nn.functional.conv2d(x, w, padding=1), rtol=1e-4, atol=1e-4)


def _torch_ref(x, off, w, pad, dg):
    """float64 autograd re-implementation of the same sampling rule (3x3, stride 1, dilation 1)."""
    B, C, H, W = x.shape
    Ho, Wo = off.shape[2:]
    cpg = C // dg
    hs = torch.arange(Ho, device=x.device, dtype=x.dtype).view(1, Ho, 1)
    ws = torch.arange(Wo, device=x.device, dtype=x.dtype).view(1, 1, Wo)
    cols = []
    for c in range(C):
        g = c // cpg
        for t in range(9):
            i, j = divmod(t, 3)
            h = hs - pad + i + off[:, g * 18 + 2 * t]
            w_ = ws - pad + j + off[:, g * 18 + 2 * t + 1]
            inside = (h >= 0) & (w_ >= 0) & (h < H) & (w_ < W)
            hl = torch.floor(h).long()
            wl = torch.floor(w_).long()
            hcl, wcl = hl >= H - 1, wl >= W - 1
            hl = torch.where(hcl, torch.full_like(hl, H - 1), hl)
            wl = torch.where(wcl, torch.full_like(wl, W - 1), wl)
            hh = torch.where(hcl, hl, hl + 1)
            wh = torch.where(wcl, wl, wl + 1)
            h2 = torch.where(hcl, hl.to(x.dtype), h)
            w2 = torch.where(wcl, wl.to(x.dtype), w_)
            lh, lw = h2 - hl.to(x.dtype), w2 - wl.to(x.dtype)
            im = x[:, c]
            idx = lambda a, b_: im.reshape(B, -1).gather(1, (a.clamp(0, H - 1) * W + b_.clamp(0, W - 1)).reshape(B, -1)).reshape(B, Ho, Wo)
            v = (1 - lh) * (1 - lw) * idx(hl, wl) + (1 - lh) * lw * idx(hl, wh) + lh * (1 - lw) * idx(hh, wl) + lh * lw * idx(hh, wh)
            cols.append(torch.where(inside, v, torch.zeros_like(v)))
    col = torch.stack(cols, 1).reshape(B, C * 9, Ho * Wo)
    return torch.einsum("fk,bkp->bfp", w.reshape(w.shape[0], -1), col).reshape(B, -1, Ho, Wo)


def test_backward_matches_autograd(cuda):
    torch.manual_seed(1)
    B, C, H, W, F, dg = 1, 4, 9, 10, 6, 2
    x = torch.randn(B, C, H, W, device=cuda)
    off = torch.randn(B, dg * 18, H, W, device=cuda) * 1.5
    w = torch.randn(F, C, 3, 3, device=cuda) * 0.2
    g = torch.randn(B, F, H, W, device=cuda)
    xs, os_, ws = [t.clone().requires_grad_(True) for t in (x, off, w)]
    y = ops.DeformableConvolution(xs, os_, ws, kernel=(3, 3), pad=(1, 1), num_deformable_group=dg, no_bias=True)
    y.backward(g)
    xd, od, wd = [t.double().clone().requires_grad_(True) for t in (x, off, w)]
    yr = _torch_ref(xd, od, wd, 1, dg)
    torch.testing.assert_close(y.double(), yr, rtol=1e-4, atol=1e-4)
    yr.backward(g.double())
    torch.testing.assert_close(xs.grad.double(), xd.grad, rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(ws.grad.double(), wd.grad, rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(os_.grad.double(), od.grad, rtol=1e-3, atol=1e-3)


def _dcn2_torch(data, offset, mask, weight, kh, kw, stride, pad, dil, dg):
    """Differentiable float64 restatement of DCNv2 (zero-padded bilinear taps x mask, then the dense
    contraction) used as the gradient reference."""
    B, C, H, W = data.shape
    Ho = (H + 2 * pad - (dil * (kh - 1) + 1)) // stride + 1
    Wo = (W + 2 * pad - (dil * (kw - 1) + 1)) // stride + 1
    hc = torch.arange(Ho, dtype=data.dtype).view(Ho, 1).expand(Ho, Wo)
    wc = torch.arange(Wo, dtype=data.dtype).view(1, Wo).expand(Ho, Wo)
    cpg = C // dg
    cols = []
    padded = torch.nn.functional.pad(data, (1, 1, 1, 1))  # zero ring: corners outside the map read 0
    for t in range(kh * kw):
        i, j = divmod(t, kw)
        taps = []
        for g in range(dg):
            h = hc * stride - pad + i * dil + offset[:, g * 2 * kh * kw + 2 * t]
            w = wc * stride - pad + j * dil + offset[:, g * 2 * kh * kw + 2 * t + 1]
            inside = ((h > -1) & (w > -1) & (h < H) & (w < W)).to(data.dtype)
            hl, wl = torch.floor(h).detach(), torch.floor(w).detach()
            lh, lw = h - hl, w - wl
            hl_i = hl.long().clamp(-1, H - 1) + 1
            wl_i = wl.long().clamp(-1, W - 1) + 1
            hh_i, wh_i = (hl_i + 1).clamp(max=H + 1), (wl_i + 1).clamp(max=W + 1)
            src = padded[:, g * cpg:(g + 1) * cpg]
            bi = torch.arange(B).view(B, 1, 1, 1)
            ci = torch.arange(cpg).view(1, cpg, 1, 1)
            def at(hi, wi):
                return src[bi, ci, hi.unsqueeze(1), wi.unsqueeze(1)]
            v = ((1 - lh) * (1 - lw)).unsqueeze(1) * at(hl_i, wl_i) + ((1 - lh) * lw).unsqueeze(1) * at(hl_i, wh_i) \
                + (lh * (1 - lw)).unsqueeze(1) * at(hh_i, wl_i) + (lh * lw).unsqueeze(1) * at(hh_i, wh_i)
            taps.append(v * (inside * mask[:, g * kh * kw + t]).unsqueeze(1))
        cols.append(torch.cat(taps, 1))
    col = torch.stack(cols, 2).reshape(B, C * kh * kw, Ho * Wo)  # (B, C, T, P) -> channel-major taps
    out = torch.einsum("fk,bkp->bfp", weight.reshape(weight.shape[0], -1), col)
    return out.reshape(B, weight.shape[0], Ho, Wo), col


@pytest.mark.parametrize("stride,dilate,pad,dg", [(1, 1, 1, 4), (2, 2, 2, 1)])
def test_modulated_dcn_forward_backward(stride, dilate, pad, dg):
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(11 + stride)
    B, C, H, W, F, k = 2, 8, 13, 17, 6, 3
    Ho = (H + 2 * pad - (dilate * (k - 1) + 1)) // stride + 1
    Wo = (W + 2 * pad - (dilate * (k - 1) + 1)) // stride + 1
    data = rng.standard_normal((B, C, H, W))
    offset = rng.standard_normal((B, dg * 2 * k * k, Ho, Wo)) * 2.5 + 0.013   # reaches outside the map
    mask = rng.uniform(0, 1, (B, dg * k * k, Ho, Wo))
    weight = rng.standard_normal((F, C, k, k)) * 0.2
    ref_in = [torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (data, offset, mask, weight)]
    ref_out, ref_col = _dcn2_torch(*ref_in, k, k, stride, pad, dilate, dg)
    gout = torch.tensor(rng.standard_normal(tuple(ref_out.shape)))
    ref_out.backward(gout)
    ins = [torch.tensor(a, dtype=torch.float32, device=dev, requires_grad=True) for a in (data, offset, mask, weight)]
    out = ops.ModulatedDeformableConvolution(*ins[:3], ins[3], None, kernel=(k, k), stride=(stride, stride),
                                             dilate=(dilate, dilate), pad=(pad, pad), num_deformable_group=dg,
                                             no_bias=True)
    out.backward(gout.to(dev, torch.float32))
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref_out.detach().numpy(), rtol=2e-4, atol=2e-4)
    for got, ref, name in zip(ins, ref_in, ("data", "offset", "mask", "weight")):
        np.testing.assert_allclose(got.grad.cpu().numpy(), ref.grad.numpy(), rtol=2e-3, atol=2e-3, err_msg=name)
