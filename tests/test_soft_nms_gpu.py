"""Batched Soft-NMS kernel vs the oracle (itself pinned against the compiled reference Cython) and
vs the committed golden vectors: boxes, rescored scores, ORDER and indices must be bit-identical."""
import os

import numpy as np
import pytest
import torch

import oracle
from simpledet_b200 import ops

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_python_ops.npz"))


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


@pytest.mark.parametrize("method", [0, 1, 2])
def test_golden(cuda, method):
    bx, idx = ops.soft_nms(_t(G["nms_dets"], cuda), 0.5, 0.3, 0.05, method)
    assert np.array_equal(bx.cpu().numpy(), G[f"soft_boxes_{method}"])
    assert np.array_equal(idx.cpu().numpy(), G[f"soft_inds_{method}"])


def test_wrapper_matches_reference_wrapper(cuda):
    out = ops.cython_soft_nms_wrapper(0.5)(_t(G["nms_dets"], cuda))
    assert np.array_equal(out.cpu().numpy(), G["soft_wrapper_linear"])


@pytest.mark.parametrize("method", [0, 1, 2])
def test_batched_random_with_ties_and_ragged_counts(cuda, method):
    rng = np.random.default_rng(10 + method)
    P, m = 12, 700
    dets = np.zeros((P, m, 5), np.float32)
    counts = rng.integers(0, m + 1, P).astype(np.int32)
    counts[0], counts[1], counts[2] = m, 1, 0
    for p in range(P):
        xy = rng.uniform(0, 250, (m, 2))
        wh = rng.uniform(5, 120, (m, 2))
        sc = rng.uniform(0, 1, (m, 1))
        sc[::3] = np.round(sc[::3], 1)  # score ties: first position must win
        dets[p] = np.concatenate([xy, xy + wh, sc], 1)
    ob, oi, oc = ops.soft_nms_batched(_t(dets, cuda), 0.5, 0.3, 0.01, method, counts=_t(counts, cuda))
    ob, oi, oc = ob.cpu().numpy(), oi.cpu().numpy(), oc.cpu().numpy()
    for p in range(P):
        rb, ri = oracle.soft_nms(dets[p, : counts[p]], 0.5, 0.3, 0.01, method)
        assert oc[p] == len(ri), p
        assert np.array_equal(ob[p, : oc[p]], rb) and np.array_equal(oi[p, : oc[p]], ri), p
        assert (oi[p, oc[p]:] == -1).all()


def test_set_nms_and_weighted_nms_match_reference_goldens(cuda):
    """operator_py/nms.py set_nms / py_weighted_nms: goldens produced by the reference's own Python."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_python_ops.npz"))
    got = ops.set_nms(torch.from_numpy(g["set_nms_dets"]).to(cuda), 0.4).cpu().numpy()
    assert np.array_equal(got, g["set_nms_0.4"])
    d = torch.from_numpy(g["nms_dets"]).to(cuda)
    for lo, hi in ((0.3, 0.6), (0.5, 0.5)):
        want = g[f"weighted_nms_{lo}_{hi}"]
        got = ops.py_weighted_nms(d, lo, hi).cpu().numpy()
        assert got.shape == want.shape
        assert np.array_equal(got[:, 4], want[:, 4])                    # same top boxes, same order
        np.testing.assert_allclose(got[:, :4], want[:, :4], rtol=2e-6, atol=1e-4)  # float32 sums, other order
    assert ops.py_weighted_nms(d, 0.3, 1.0).shape[0] == 0               # nothing can vote: the reference breaks
