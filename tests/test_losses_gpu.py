"""FocalLoss / BBoxNorm / SigmoidCrossEntropy: fused CUDA kernels vs the oracle (floats through
powf/logf/expf: rtol 1e-4 as north_star states, plus a small atol where the expression cancels)."""
import numpy as np
import pytest
import torch

import oracle
from simpledet_b200 import ops

pytestmark = pytest.mark.gpu


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


@pytest.mark.parametrize("norm", ["valid", "batch", "null"])
@pytest.mark.parametrize("use_og", [False, True])
def test_focal_loss(cuda, norm, use_og):
    rng = np.random.default_rng(2)
    B, N, K = 2, 5000, 80
    data = (rng.standard_normal((B, N, K)) * 2 - 3).astype(np.float32)
    label = np.zeros((B, N), np.float32)
    label[:, rng.choice(N, 60, replace=False)] = rng.integers(1, K + 1, (B, 60))
    label[:, rng.choice(N, 100, replace=False)] = -1
    og = rng.uniform(0.5, 1.5, data.shape).astype(np.float32)
    d = _t(data, cuda).requires_grad_(True)
    out = ops.FocalLoss(d, _t(label, cuda), alpha=0.25, gamma=2.0, normalization=norm, grad_scale=1.5,
                        out_grad=use_og)
    ro = oracle.sigmoid(data)
    np.testing.assert_allclose(out.detach().cpu().numpy(), ro, rtol=1e-5, atol=1e-7)
    out.backward(_t(og, cuda))
    # the oracle backward consumes the op's own forward output, like the reference
    rg = oracle.focal_loss_backward(out.detach().cpu().numpy(), label, 0.25, 2.0, 1.5, norm, og if use_og else None)
    np.testing.assert_allclose(d.grad.cpu().numpy(), rg, rtol=1e-4, atol=1e-7)
    assert (d.grad.cpu().numpy()[label == -1] == 0).all()


def test_bbox_norm(cuda):
    rng = np.random.default_rng(3)
    data = rng.standard_normal((2, 36, 2000)).astype(np.float32)
    label = rng.integers(-1, 3, (2, 9 * 2000)).astype(np.float32)
    g = rng.standard_normal(data.shape).astype(np.float32)
    d = _t(data, cuda).requires_grad_(True)
    out = ops.BBoxNorm(d, _t(label, cuda))
    assert torch.equal(out.detach(), d.detach())
    out.backward(_t(g, cuda))
    assert np.array_equal(d.grad.cpu().numpy(), oracle.bbox_norm_backward(g, label))


def test_sigmoid_cross_entropy(cuda):
    rng = np.random.default_rng(4)
    R, D = 3, 2 * 128 * 28 * 28 // 3
    data = (rng.standard_normal((R, D)) * 3).astype(np.float32)
    label = rng.integers(-1, 2, (R, D)).astype(np.float32)
    label[2] = -1  # a fully ignored row: count = 1e-5
    d = _t(data, cuda).requires_grad_(True)
    out = ops.SigmoidCrossEntropy(d, _t(label, cuda), grad_scale=0.7)
    np.testing.assert_allclose(out.detach().cpu().numpy(), oracle.sigmoid_ce_forward(data, label), rtol=2e-4)
    out.sum().backward()
    np.testing.assert_allclose(d.grad.cpu().numpy(), oracle.sigmoid_ce_backward(data, label, 0.7), rtol=1e-4,
                               atol=1e-9)
