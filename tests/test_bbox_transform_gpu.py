"""Device box utilities vs the goldens produced by the reference's own bbox_transform.py / Cython."""
import os

import numpy as np
import pytest
import torch

import oracle
import oracle.np_ops
from simpledet_b200 import ops

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_python_ops.npz"))


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def test_overlaps_golden_and_large(cuda):
    b, q = G["overlaps_boxes"], G["overlaps_query"]
    assert np.array_equal(ops.bbox_overlaps(_t(b, cuda), _t(q, cuda)).cpu().numpy(), G["overlaps_out"])
    assert np.array_equal(ops.bbox_overlaps(_t(b, cuda), _t(q, cuda), "ioa").cpu().numpy(), G["selfoverlaps_out"])
    rng = np.random.default_rng(3)
    xy = rng.uniform(0, 900, (20000, 2))
    big = np.concatenate([xy, xy + rng.uniform(0, 300, (20000, 2))], 1).astype(np.float32)
    big[:50, 2:] = big[:50, :2]  # one-pixel boxes
    qq = big[rng.permutation(20000)[:100]]
    assert np.array_equal(ops.bbox_overlaps(_t(big, cuda), _t(qq, cuda)).cpu().numpy(), oracle.bbox_overlaps(big, qq))
    assert np.array_equal(ops.bbox_overlaps(_t(big, cuda), _t(qq, cuda), "ioa").cpu().numpy(),
                          oracle.bbox_selfoverlaps(big, qq))
    assert ops.bbox_overlaps(_t(big[:0], cuda), _t(qq, cuda)).shape == (0, 100)


def test_encode_decode_golden(cuda):
    ex, gt, d = G["xf_ex"], G["xf_gt"], G["xf_deltas"]
    enc = ops.nonlinear_transform(_t(ex, cuda), _t(gt, cuda)).cpu().numpy()
    np.testing.assert_allclose(enc, G["nonlinear_transform"], rtol=1e-14, atol=1e-15)
    ex32 = ex.astype(np.float32)
    dec = ops.nonlinear_pred(_t(ex32, cuda), _t(d, cuda)).cpu().numpy()
    np.testing.assert_allclose(dec, G["nonlinear_pred"], rtol=1e-14, atol=1e-12)
    np.testing.assert_array_equal(ops.iou_pred(_t(ex32, cuda), _t(d, cuda)).cpu().numpy(), G["iou_pred"])
    clipped = ops.nonlinear_pred(_t(ex32, cuda), _t(d, cuda), im_shape=(400, 500)).cpu().numpy()
    np.testing.assert_allclose(clipped, G["clip_boxes"], rtol=1e-14, atol=1e-12)
    assert clipped.min() >= 0 and clipped[:, 0::4].max() <= 499 and clipped[:, 1::4].max() <= 399
    # round trip: decode(encode(ex -> gt)) == gt
    rt = ops.nonlinear_pred(_t(ex32, cuda), ops.nonlinear_transform(_t(ex32.astype(np.float64), cuda), _t(gt, cuda)))
    np.testing.assert_allclose(rt.cpu().numpy(), gt, rtol=1e-9, atol=1e-9)


def test_flip_and_box_voting_golden(cuda):
    ex = G["xf_ex"]
    assert np.array_equal(ops.flip_boxes(_t(ex, cuda), 640).cpu().numpy(), G["flip_boxes"])
    f32 = ex.astype(np.float32)
    assert np.array_equal(ops.flip_boxes(_t(f32, cuda), 640).cpu().numpy(), oracle.np_ops.flip_boxes(f32, 640))
    top, alld = G["vote_top"], G["nms_dets"]
    for meth, beta in (("ID", 1.0), ("AVG", 1.0), ("IOU_AVG", 1.0), ("GENERALIZED_AVG", 2.0), ("QUASI_SUM", 0.5),
                       ("TEMP_AVG", 0.7)):
        got = ops.box_voting(_t(top, cuda), _t(alld, cuda), 0.5, meth, beta).cpu().numpy()
        # voter sets are exact (bit-exact IoU); the float32 sums run in warp order, numpy's in pairwise order
        np.testing.assert_allclose(got, G[f"vote_{meth}"], rtol=3e-6, atol=1e-4, err_msg=meth)
