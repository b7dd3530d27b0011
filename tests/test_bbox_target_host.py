"""CustomOp 'bbox_target' (operator_py/bbox_target.py): the host logic of `simpledet_b200.ops.bbox_target` - padding
filters, gt append, first-maximum match, the two numpy.random.choice draws in the reference's order, label / target /
weight layout - driven on CPU tensors with the oracle's IoU in place of the CUDA kernel, against vectors produced by
running the reference's operator unmodified (tests/golden/make_golden_bbox_target.py).  The product entry point
itself refuses CPU tensors (checked here); its GPU twin is tests/test_zz_late_gpu.py."""
import os

import numpy as np
import pytest
import torch

import oracle
from simpledet_b200 import ops

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_bbox_target.npz"))


def kwargs_of(name):
    kw = {}
    for k in ("num_class", "add_gt_to_proposal", "image_rois", "fg_fraction", "fg_thresh", "bg_thresh_hi", "bg_thresh_lo"):
        kw[k] = G[f"{name}_kw_{k}"].item()
    kw["bbox_target_std"] = tuple(float(v) for v in G[f"{name}_kw_bbox_target_std"])
    return kw


def cpu_overlaps(a, b):
    return torch.from_numpy(oracle.bbox_overlaps(a.numpy(), b.numpy()))


def check_case(name, out):
    rois, label, target, weight = (o.cpu().numpy() for o in out)
    assert np.array_equal(rois, G[f"{name}_rois"])
    assert np.array_equal(label, G[f"{name}_label"])
    assert np.array_equal(weight, G[f"{name}_weight"])
    # float32 log: libm (numpy) vs the device's / torch's logf, a few ulp on values of order 1-10
    np.testing.assert_allclose(target, G[f"{name}_target"], rtol=2e-6, atol=2e-6)
    assert np.array_equal(target != 0, G[f"{name}_target"] != 0)


@pytest.mark.parametrize("name", [str(n) for n in G["names"]])
def test_host_logic_against_the_reference_operator(name):
    kw = kwargs_of(name)
    rng = np.random.RandomState(int(G[f"{name}_seed"]))          # == np.random.seed(seed) + the global functions
    out = ops._bbox_target_impl(torch.from_numpy(G[f"{name}_prop"]), torch.from_numpy(G[f"{name}_gt"]), rng=rng,
                                overlaps=cpu_overlaps, **kw)
    check_case(name, out)


def test_ragged_and_empty_images_raise():
    name = "c81"
    kw = kwargs_of(name)
    prop, gt = torch.from_numpy(G[f"{name}_prop"]).clone(), torch.from_numpy(G[f"{name}_gt"]).clone()
    few = prop.clone()
    few[:, 20:] = 0                                                 # 20 proposals + gt < image_rois
    with pytest.raises(ValueError, match="ragged"):
        ops._bbox_target_impl(few, gt, rng=np.random.RandomState(0), overlaps=cpu_overlaps, **kw)
    nogt = gt.clone()
    nogt[1, :, 4] = -1
    with pytest.raises(ValueError, match="without ground-truth"):
        ops._bbox_target_impl(prop, nogt, rng=np.random.RandomState(0), overlaps=cpu_overlaps, **kw)


def test_product_entry_point_refuses_cpu_tensors():
    name = "c81"
    with pytest.raises(Exception):
        ops.bbox_target(torch.from_numpy(G[f"{name}_prop"]), torch.from_numpy(G[f"{name}_gt"]), **kwargs_of(name))
    assert ops.OPS["bbox_target"] is ops.bbox_target


def test_product_wrapper_with_the_device_checks_lifted(monkeypatch):
    """ops.bbox_target itself (string attributes as the CustomOp hands them over, the global numpy RNG) with only the
    CUDA tensor check and the IoU kernel replaced."""
    monkeypatch.setattr(ops, "_dev", lambda t, name, dtype=torch.float32: t.contiguous())
    monkeypatch.setattr(ops, "bbox_overlaps", cpu_overlaps)
    for name in (str(n) for n in G["names"]):
        kw = kwargs_of(name)
        kw["add_gt_to_proposal"] = str(bool(kw["add_gt_to_proposal"]))      # CustomOpProp passes strings
        kw["bbox_target_std"] = str(kw["bbox_target_std"])
        np.random.seed(int(G[f"{name}_seed"]))
        check_case(name, ops.OPS["bbox_target"](torch.from_numpy(G[f"{name}_prop"]), torch.from_numpy(G[f"{name}_gt"]), **kw))
