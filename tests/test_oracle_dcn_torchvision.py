"""The deformable-convolution restatements against an INDEPENDENT implementation of the same published operator:
torchvision.ops.deform_conv2d (CPU kernel; its lineage is the original MXNet DCN code via mmdetection).  The MXNet
source the reference calls is not in the reference tree, so this is the strongest pin this image
allows: a third-party implementation, not the reference itself - DESIGN.md keeps DCN under "parity unpinned".

* DCNv2 sampling rule (taps valid for h in (-1, H), zero-padded bilinear corners, x mask): torchvision implements
  exactly this; compared everywhere, forward and all four gradients.
* DCNv1 in MXNet 1.x differs at the border only (taps valid for h in [0, H), corners clamped to H-1 instead of
  zero-padded): compared on samples that lie inside [0, H-1] x [0, W-1], where the two rules coincide, and a second
  assertion shows that the border rule really is the only difference."""
import numpy as np
import pytest
import torch

tv = pytest.importorskip("torchvision.ops")

from oracle import np_ops  # noqa: E402
from test_dcn_gpu import _dcn2_torch  # noqa: E402  (the float64 restatement the GPU tests differentiate)


def _geometry(H, W, k, stride, pad, dil):
    Ho = (H + 2 * pad - (dil * (k - 1) + 1)) // stride + 1
    Wo = (W + 2 * pad - (dil * (k - 1) + 1)) // stride + 1
    return Ho, Wo


@pytest.mark.parametrize("stride,dilate,pad,dg", [(1, 1, 1, 4), (2, 2, 2, 1), (1, 2, 0, 2)])
def test_dcnv2_restatement_equals_torchvision(stride, dilate, pad, dg):
    g = torch.Generator().manual_seed(7 + dg)
    B, C, H, W, F, k = 2, 8, 11, 13, 6, 3
    Ho, Wo = _geometry(H, W, k, stride, pad, dilate)
    x = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
    off = torch.randn(B, dg * 2 * k * k, Ho, Wo, generator=g, dtype=torch.float64) * 2.0   # many taps leave the map
    msk = torch.rand(B, dg * k * k, Ho, Wo, generator=g, dtype=torch.float64)
    w = torch.randn(F, C, k, k, generator=g, dtype=torch.float64) * 0.3
    go = torch.randn(B, F, Ho, Wo, generator=g, dtype=torch.float64)
    a = [t.clone().requires_grad_(True) for t in (x, off, msk, w)]
    b = [t.clone().requires_grad_(True) for t in (x, off, msk, w)]
    ya, _ = _dcn2_torch(a[0], a[1], a[2], a[3], k, k, stride, pad, dilate, dg)
    yb = tv.deform_conv2d(b[0], b[1], b[3], None, stride=stride, padding=pad, dilation=dilate, mask=b[2])
    torch.testing.assert_close(ya, yb, rtol=1e-10, atol=1e-10)
    ya.backward(go)
    yb.backward(go)
    for name, p, q in zip(("data", "offset", "mask", "weight"), a, b):
        torch.testing.assert_close(p.grad, q.grad, rtol=1e-8, atol=1e-9, msg=lambda m, n=name: f"grad {n}: {m}")


@pytest.mark.parametrize("stride,dilate,dg", [(1, 1, 1), (1, 1, 4), (2, 1, 2), (1, 2, 2)])
def test_dcnv1_restatement_equals_torchvision_inside_the_map(stride, dilate, dg):
    rng = np.random.default_rng(3 + dg)
    B, C, H, W, F, k, pad = 2, 8, 12, 14, 5, 3, 0
    Ho, Wo = _geometry(H, W, k, stride, pad, dilate)
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    off = (rng.standard_normal((B, dg * 2 * k * k, Ho, Wo)) * 1.5).astype(np.float32)
    # keep every sample inside [0, H-1] x [0, W-1]: there MXNet's rule (clamp) and torchvision's (zero ring) agree
    hc, wc = np.meshgrid(np.arange(Ho), np.arange(Wo), indexing="ij")
    for t in range(k * k):
        i, j = divmod(t, k)
        bh = (hc * stride - pad + i * dilate).astype(np.float32)
        bw = (wc * stride - pad + j * dilate).astype(np.float32)
        for gidx in range(dg):
            oh = off[:, gidx * 2 * k * k + 2 * t]
            ow = off[:, gidx * 2 * k * k + 2 * t + 1]
            oh[...] = np.clip(bh + oh, 0, H - 1) - bh
            ow[...] = np.clip(bw + ow, 0, W - 1) - bw
    w = (rng.standard_normal((F, C, k, k)) * 0.3).astype(np.float32)
    col = np_ops.deformable_im2col(x, off, (k, k), (stride, stride), (dilate, dilate), (pad, pad), dg)
    ours = np.einsum("fk,bkp->bfp", w.reshape(F, -1).astype(np.float64), col.astype(np.float64)).reshape(B, F, Ho, Wo)
    ref = tv.deform_conv2d(torch.from_numpy(x).double(), torch.from_numpy(off).double(), torch.from_numpy(w).double(), None,
                           stride=stride, padding=pad, dilation=dilate).numpy()
    np.testing.assert_allclose(ours, ref, rtol=2e-5, atol=2e-5)       # float32 sampling vs float64


def test_dcnv1_border_rule_is_the_only_difference():
    """A tap at h = H - 0.5: MXNet 1.x DCNv1 clamps the low corner to H-1 and drops the fraction (value of row H-1),
    torchvision blends row H-1 with a zero ring (half of it).  A tap at h = -0.5 is outside for MXNet (0) and half of
    row 0 for torchvision.  The restatement follows MXNet's rule (what the reference executes)."""
    H = W = 4
    x = np.arange(16, dtype=np.float32).reshape(1, 1, H, W) + 1
    off = np.zeros((1, 2, H, W), np.float32)       # 1x1 kernel, one tap
    off[0, 0, 3, :] = 0.5                           # bottom row samples at h = 3.5
    off[0, 0, 0, :] = -0.5                          # top row samples at h = -0.5
    col = np_ops.deformable_im2col(x, off, (1, 1)).reshape(H, W)
    ref = tv.deform_conv2d(torch.from_numpy(x), torch.from_numpy(off), torch.ones(1, 1, 1, 1)).numpy().reshape(H, W)
    assert np.array_equal(col[3], x[0, 0, 3]) and np.allclose(ref[3], 0.5 * x[0, 0, 3])
    assert np.array_equal(col[0], np.zeros(W, np.float32)) and np.allclose(ref[0], 0.5 * x[0, 0, 0])
    assert np.array_equal(col[1:3], ref[1:3])
