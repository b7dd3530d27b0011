"""ROIPooling_v1 CUDA vs oracle (integer bins, max + flat argmax): bit-exact forward incl. the
reference docstring vector; backward within 1e-4 (atomics)."""
import numpy as np
import pytest
import torch

import oracle
from simpledet_b200 import ops

pytestmark = pytest.mark.gpu


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def test_docstring_example(cuda):
    x = np.arange(48, dtype=np.float32).reshape(1, 1, 8, 6)
    rois = np.array([[0, 0, 0, 4, 4]], np.float32)
    out, idx = ops.roi_pooling_v1_raw(_t(x, cuda), _t(rois, cuda), (2, 2), 1.0)
    assert out.cpu().reshape(2, 2).tolist() == [[14.0, 16.0], [26.0, 28.0]]
    out, idx = ops.roi_pooling_v1_raw(_t(x, cuda), _t(rois, cuda), (2, 2), 0.7)
    assert out.cpu().reshape(2, 2).tolist() == [[7.0, 9.0], [19.0, 21.0]]


@pytest.mark.parametrize("pooled", [(7, 7), (14, 14), (2, 3)])
def test_random_bit_exact_and_backward(cuda, pooled):
    rng = np.random.default_rng(pooled[0])
    data = rng.standard_normal((2, 19, 38, 50)).astype(np.float32)
    R = 70
    xy = rng.uniform(-40, 700, (R, 2))
    wh = rng.uniform(1, 400, (R, 2))
    rois = np.concatenate([rng.integers(0, 2, (R, 1)), xy, xy + wh], 1).astype(np.float32)
    rois[0, 1:] = 0  # 1x1 roi at the origin
    rois[1, 1:] = [900, 900, 950, 950]  # outside -> empty bins
    rois[2, 1:] = [300, 300, 100, 100]  # malformed -> forced 1x1
    out, idx = ops.roi_pooling_v1_raw(_t(data, cuda), _t(rois, cuda), pooled, 1 / 16)
    ro, ri = oracle.roi_pool_v1_forward(data, rois, pooled, 1 / 16)
    assert np.array_equal(out.cpu().numpy(), ro) and np.array_equal(idx.cpu().numpy(), ri)
    g = rng.standard_normal(ro.shape).astype(np.float32)
    d = _t(data, cuda).requires_grad_(True)
    o = ops.ROIPooling_v1(d, _t(rois, cuda), pooled, 1 / 16)
    o.backward(_t(g, cuda))
    rg = oracle.roi_pool_v1_backward(g, ri, rois, data.shape)
    np.testing.assert_allclose(d.grad.cpu().numpy(), rg, rtol=1e-4, atol=1e-4)
