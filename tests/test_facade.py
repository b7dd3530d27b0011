"""The `mx` / `mxnext` façade on CPU: the reference's OWN config and builders import and build
their graph on the stand-in modules; graph queries, MXNet reshape codes and shape inference; the committed graph
fixture is what those builders produce."""
import importlib
import os
import sys

import pytest

from simpledet_b200.facade import executor as E
from simpledet_b200.facade import symbol as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "golden", "faster_r50v1_fpn_1x_test_symbol.json")
SHAPES = dict(data=(1, 3, 800, 1333), im_info=(1, 3), im_id=(1,), rec_id=(1,))


def test_mx_reshape_codes():
    assert E.mx_reshape((2, 3, 4, 5), (0, -1)) == (2, 60)
    assert E.mx_reshape((2, 3, 4, 5), (-3, -2)) == (6, 4, 5)              # symbol/builder.py:896
    assert E.mx_reshape((1, 6, 50, 84), (0, 2, -1, 0)) == (1, 2, 150, 84)  # models/FPN/builder.py:262
    assert E.mx_reshape((2, 12, 5), (0, -4, 3, -1, 0)) == (2, 3, 4, 5)
    assert E.mx_reshape((2, 3, 4), (-2,)) == (2, 3, 4)


def test_fixture_graph_shapes_and_ops():
    sym = S.fromjson(open(FIX).read())
    args, outs, aux = sym.infer_shape(**SHAPES)
    assert sym.list_outputs() == ["rec_id", "im_id", "im_info", "bbox_cls_score_reshape_output", "decode_bbox_output"]
    assert outs == [(1,), (1,), (1, 3), (1, 1000, 81), (1, 1000, 324)]
    assert all(s is not None for s in args + aux)
    ops = {n.op for n in sym._topo() if n.op}
    # the detection operators are in the graph under the reference's registration strings
    assert {"_contrib_Proposal_v3", "_contrib_ROIAlign_v2", "_contrib_DecodeBBox", "Custom"} <= ops
    custom = {n.attrs["op_type"] for n in sym._topo() if n.op == "Custom"}
    assert custom == {"get_top_proposal", "assign_layer_fpn"}
    nparam = sum(__import__("math").prod(s) for n, s in zip(sym.list_arguments(), args) if n not in SHAPES)
    assert 41e6 < nparam < 43e6   # ResNet-50 FPN Faster R-CNN: ~41.8 M parameters
    # the executor recognises assign_layer_fpn -> 4 x ROIAlign_v2 -> add_n as ONE fused FPN RoIAlign
    ex = E.Executor(sym, device="cpu")
    assert len(ex._fusions) == 1


@pytest.mark.skipif(not os.path.isdir("/root/reference/config"), reason="reference checkout not present")
def test_reference_builders_run_unchanged_on_the_facade():
    from simpledet_b200 import facade

    facade.install("/root/reference")
    cfg = importlib.import_module("config.faster_r50v1_fpn_1x")
    sym = cfg.get_config(is_train=False)[6].test_symbol
    want = S.fromjson(open(FIX).read())
    assert [n.name for n in sym._topo()] == [n.name for n in want._topo()]
    assert [n.op for n in sym._topo()] == [n.op for n in want._topo()]
    assert sym.infer_shape(**SHAPES)[1] == want.infer_shape(**SHAPES)[1]
    # rpn-only symbol of the same config
    rpn = cfg.get_config(is_train=False)[6].rpn_test_symbol
    assert rpn.infer_shape(**SHAPES)[1][-2:] == [(1, 1000, 4), (1, 1000, 1)]
    for m in [k for k in sys.modules if k.split(".")[0] in ("config", "symbol", "models", "core", "utils", "operator_py")]:
        sys.modules.pop(m, None)   # leave no half-imported reference packages behind for other tests


def test_frozen_batchnorm_is_folded_into_the_convolution():
    """Convolution -> BatchNorm on moving statistics (fixbn) runs as ONE convolution with folded weights; same
    values up to fp32 rounding, and a BatchNorm whose input is shared is left alone."""
    import torch

    from simpledet_b200.facade import mxnext_impl as X

    data = S.Variable("data")
    c1 = X.conv(data, "c1", 8, kernel=3, stride=1)
    r = X.relu(X.fixbn(c1, "bn1"))
    c2 = X.conv(r, "c2", 4, kernel=1)
    shared = X.add_n(X.fixbn(c2, "bn2"), c2, name="sum")  # c2 has two consumers: bn2 must not be folded
    shapes = dict(data=(1, 3, 16, 20))
    x = torch.randn(shapes["data"], generator=torch.Generator().manual_seed(2))
    res = {}
    for fold in (False, True):
        ex = E.Executor(shared, "cpu", channels_last=False, fuse_fpn_roi_align=False, fold_bn=fold)
        ex.init_params(shapes, rng_std=0.3)
        g = torch.Generator().manual_seed(1)
        for k, v in ex.params.items():
            if k.endswith("moving_var"):
                ex.params[k] = torch.rand(v.shape, generator=g) + 0.5
            if k.endswith(("moving_mean", "beta")):
                ex.params[k] = torch.randn(v.shape, generator=g)
        ex._folded = None
        assert len(ex._bn_of_conv) == (1 if fold else 0)
        res[fold] = ex.forward(data=x)[0]
    torch.testing.assert_close(res[True], res[False], rtol=1e-5, atol=1e-5)
