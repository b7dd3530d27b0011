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


# ---- other detectors of the reference on the same stand-ins ("the four configs import and run unchanged")
_PROBE = r"""
import importlib, json, math, sys
sys.path.insert(0, {root!r})
from simpledet_b200 import facade
from simpledet_b200.facade import executor as E
facade.install("/root/reference")
cfg = importlib.import_module("config." + {name!r})
shapes = dict(data=(1, 3, 800, 1333), im_info=(1, 3), im_id=(1,), rec_id=(1,))
test = cfg.get_config(is_train=False)[6].test_symbol
try:
    args, outs, aux = test.infer_shape(**shapes)
    npar = sum(math.prod(s) for n, s in zip(test.list_arguments(), args) if n not in shapes)
    fusions = len(E.Executor(test, device="cpu")._fusions)
except NotImplementedError:          # graph built, an operator of it has no shape rule in the executor
    outs = npar = fusions = None
train = cfg.get_config(is_train=True)[6].train_symbol
print(json.dumps(dict(outs=outs, npar=npar, ops=sorted({{n.op for n in test._topo() if n.op}}), fusions=fusions,
                      train_ops=sorted({{n.op for n in train._topo() if n.op}}), train_outs=train.list_outputs())))
"""

# name -> (visible outputs after rec_id / im_id / im_info, parameter count in millions, operators that must be present)
DETECTORS = {
    "retina_r50v1_fpn_1x": ([(1, 5000, 81), (1, 5000, 4)], (37, 39), {"_contrib_GenAnchor", "_contrib_GenProposalRetina"},
                            {"_contrib_FocalLoss", "_contrib_BBoxNorm"}),
    "mask_r50v1_fpn_1x": ([(1, 100, 1), (1, 100, 4), (1, 100, 1), (100, 81, 28, 28), (1,)], (43, 46),
                          {"_contrib_Proposal_v3", "_contrib_ROIAlign_v2", "Deconvolution", "Custom"},
                          {"ProposalMaskTarget", "_contrib_SigmoidCrossEntropy"}),
    # the test graph carries five fc1 weights: three stages + the 1st / 2nd heads re-applied to 3rd-stage rois
    "cascade_r50v1_fpn_1x": ([(1, 1000, 81), (1, 1000, 4)], (95, 99), {"_contrib_DecodeBBox", "_contrib_ROIAlign_v2"},
                             {"ProposalTarget"}),
    "faster_r50v1c4_c5_512roi_1x": ([(1, 1000, 81), (1, 1000, 4)], (27, 30), {"_contrib_Proposal", "_contrib_ROIAlign_v2"},
                                    {"ProposalTarget"}),
    "dcn.faster_dcn_r50v1bc4_c5_512roi_1x": ([(1, 300, 81), (1, 300, 4)], (28, 32), {"_contrib_DeformableConvolution"},
                                             {"ProposalTarget", "_contrib_DeformableConvolution"}),
    # Mask Scoring R-CNN: the graphs build (its training graph is the ProposalMaskTarget(output_ratio=True) caller);
    # the test graph's index gymnastics (arange / stack / gather_nd) are not wired into the executor
    "ms_r50v1_fpn_1x": (None, None, {"_contrib_Proposal_v3", "gather_nd"}, {"ProposalMaskTarget"}),
    # TridentNet (three weight-sharing branches stacked into the batch axis): the callers of Proposal_v2 and
    # ProposalTarget_v2 (valid_ranges); ResNet-v1 helper and pre-activation ResNet-v2 stand-ins
    "tridentnet_r50v1c4_c5_1x": (None, None, {"_contrib_Proposal", "stack"}, {"_contrib_Proposal_v2", "ProposalTarget_v2"}),
    "tridentnet_r50v2c4_c5_1x": (None, None, {"_contrib_Proposal", "stack"}, {"_contrib_Proposal_v2", "ProposalTarget_v2"}),
    "faster_r50v2c4_c5_256roi_1x": ([(1, 1000, 81), (1, 1000, 4)], (25, 31), {"_contrib_Proposal"}, {"ProposalTarget"}),
    # samples of the rest of the model zoo (tools/facade_sweep.py builds all 116 configs: 113 inference / 110 training
    # graphs): SyncBatchNorm + feature-pyramid grids, GroupNorm heads with symbol comparisons in the loss
    "FPG.faster_r50v1b_fpg6@128_syncbn_1x": (None, None, {"_contrib_SyncBatchNorm", "_contrib_ROIAlign_v2"}, {"ProposalTarget"}),
    "fcos_r50v1_fpn_1x": (None, None, {"_contrib_GroupNorm"}, {"_contrib_GroupNorm"}),
    # FreeAnchor: mxnext.tvm.decode_bbox (the TVM twin of _contrib_DecodeBBox) in the loss and in the test-time proposal
    "FreeAnchor.free_anchor_r50v1_fpn_1x": (None, None, {"_contrib_DecodeBBox"}, {"_contrib_DecodeBBox"}),
}


@pytest.mark.skipif(not os.path.isdir("/root/reference/config"), reason="reference checkout not present")
@pytest.mark.parametrize("name", sorted(DETECTORS))
def test_other_reference_detectors_build_on_the_facade(name):
    """Each config is built in its own interpreter: the reference caches the RPN sub-graph in a class attribute
    (symbol/builder.py FasterRcnn._rpn_output), so two detectors built in one process share nodes - there as here."""
    import json
    import subprocess

    r = subprocess.run([sys.executable, "-c", _PROBE.format(root=ROOT, name=name)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    got = json.loads(r.stdout.strip().splitlines()[-1])
    outs, npar, need, need_train = DETECTORS[name]
    if outs is not None:
        assert [tuple(s) for s in got["outs"][3:]] == outs
        assert npar[0] * 1e6 < got["npar"] < npar[1] * 1e6, got["npar"]
        assert got["fusions"] is not None
    assert need <= set(got["ops"]) and need_train <= set(got["train_ops"])
    assert got["train_outs"]


# ---- executor glue for the other detectors, without a GPU: every detection operator replaced by a stub that checks the
# ---- arguments it is handed (names, shapes, attribute parsing) and returns tensors of the operator's output shapes
def _stubs(torch, log):
    def proposal(cls_prob, bbox_pred, im_info, rpn_pre_nms_top_n, rpn_post_nms_top_n, threshold, rpn_min_size, scales,
                 ratios, feature_stride, output_score, iou_loss):
        B, A2, H, W = cls_prob.shape
        assert A2 == 2 * len(scales) * len(ratios) and tuple(bbox_pred.shape) == (B, 2 * A2, H, W)
        assert tuple(im_info.shape) == (B, 3) and output_score and 0 < threshold < 1
        log.append(("proposal", feature_stride))
        return torch.rand(B, rpn_post_nms_top_n, 4) * 100, torch.rand(B, rpn_post_nms_top_n, 1)

    def gen_anchor(cls_prob, feature_stride, scales, ratios):
        log.append(("gen_anchor", feature_stride))
        return torch.zeros(cls_prob.shape[2] * cls_prob.shape[3] * len(scales) * len(ratios), 4)

    def gen_proposal_retina(cls_prob, bbox_pred, im_info, anchors, feature_stride, rpn_pre_nms_top_n, rpn_min_size,
                            num_anchors, thresh, anchor_mean, anchor_std, output_one_hot):
        B, AK, H, W = cls_prob.shape
        assert tuple(bbox_pred.shape) == (B, 4 * num_anchors, H, W) and tuple(anchors.shape) == (H * W * num_anchors, 4)
        assert len(anchor_mean) == len(anchor_std) == 4 and 0 <= thresh < 1 and AK % num_anchors == 0
        log.append(("gen_proposal_retina", feature_stride, thresh))
        return torch.zeros(B, rpn_pre_nms_top_n, 4), torch.zeros(B, rpn_pre_nms_top_n, AK // num_anchors + 1)

    def get_top_proposal(bbox, score, top_n):
        assert bbox.shape[:2] == score.shape[:2]
        return bbox[:, :top_n].contiguous(), score[:, :top_n].contiguous()

    def decode_bbox(rois, bbox_pred, im_info, mean, std, class_agnostic):
        assert rois.shape[:2] == bbox_pred.shape[:2] and len(mean) == len(std) == 4
        return rois.clone() if class_agnostic else bbox_pred.clone()

    def post_processing(cls_score, bbox_xyxy, max_det_per_image, min_det_score, nms_type, nms_thr):
        B = cls_score.shape[0]
        assert bbox_xyxy.shape[:2] == cls_score.shape[:2] and nms_type == "nms" and 0 < nms_thr < 1
        log.append(("post", max_det_per_image))
        return (torch.rand(B, max_det_per_image, 1), torch.rand(B, max_det_per_image, 4) * 100,
                torch.zeros(B, max_det_per_image, 1))

    def deform_conv(data, offset, weight, bias, kernel, stride, dilate, pad, num_filter, num_group, num_deformable_group,
                    no_bias):
        B, C, H, W = data.shape
        Ho = (H + 2 * pad[0] - (dilate[0] * (kernel[0] - 1) + 1)) // stride[0] + 1
        Wo = (W + 2 * pad[1] - (dilate[1] * (kernel[1] - 1) + 1)) // stride[1] + 1
        assert tuple(offset.shape) == (B, 2 * num_deformable_group * kernel[0] * kernel[1], Ho, Wo)
        assert tuple(weight.shape) == (num_filter, C // num_group, *kernel) and (bias is None) == no_bias
        log.append(("dcn", C))
        return torch.zeros(B, num_filter, Ho, Wo)

    return {"_contrib_Proposal_v3": proposal, "_contrib_Proposal": proposal, "_contrib_GenAnchor": gen_anchor,
            "_contrib_GenProposalRetina": gen_proposal_retina, "get_top_proposal": get_top_proposal,
            "_contrib_DecodeBBox": decode_bbox, "BboxPostProcessing": post_processing,
            "_contrib_DeformableConvolution": deform_conv}


@pytest.mark.parametrize("name,outs", [
    ("retina_r50v1_fpn_1x", [(1, 5000, 81), (1, 5000, 4)]),
    ("mask_r50v1_fpn_1x", [(1, 100, 1), (1, 100, 4), (1, 100, 1), (100, 81, 28, 28), (1,)]),
    ("faster_dcn_r50v1bc4_c5_512roi_1x", [(1, 300, 81), (1, 300, 4)]),
    # TridentNet: three dilation branches with shared weights stacked into the batch axis (mx.symbol.stack + Reshape),
    # one legacy Proposal over the 3-image batch, 3 x 300 rois through the C5 head
    ("tridentnet_r50v1c4_c5_1x", [(900, 81), (900, 4)]),
    # Cascade R-CNN: three fused FPN RoIAlign stages chained through DecodeBBox, the ensembled score of five head passes
    ("cascade_r50v1_fpn_1x", [(1, 1000, 81), (1, 1000, 4)]),
])
def test_executor_glue_of_the_other_detectors(name, outs, monkeypatch):
    import torch

    from simpledet_b200 import ops

    sym = S.fromjson(open(os.path.join(ROOT, "tests", "golden", name + "_test_symbol.json")).read())
    shapes = dict(data=(1, 3, 128, 192), im_info=(1, 3), im_id=(1,), rec_id=(1,))
    log = []
    for k, fn in _stubs(torch, log).items():
        monkeypatch.setitem(ops.OPS, k, fn)

    def roi_align(data, rois, pooled_size, spatial_scale, with_argmax=True, **kw):
        assert rois.shape[0] == data.shape[0] and rois.shape[2] == 4 and 0 < spatial_scale < 1
        o = torch.zeros(rois.shape[0], rois.shape[1], data.shape[1], *pooled_size)
        return o, o, o

    def fpn_roi_align(feats, rois, strides, out_size, scale0, lvl0, with_argmax=True):
        assert len(feats) == len(strides) and all(f.shape[1] == feats[0].shape[1] for f in feats)
        log.append(("fpn_roi_align", tuple(out_size)))
        return (torch.zeros(rois.shape[0], rois.shape[1], feats[0].shape[1], *out_size),)

    monkeypatch.setattr(ops, "roi_align_v2_raw", roi_align)
    monkeypatch.setattr(ops, "fpn_roi_align_raw", fpn_roi_align)
    ex = E.Executor(sym, device="cpu", channels_last=False).init_params(shapes, rng_std=0.01)
    feed = dict(data=torch.ones(shapes["data"]), im_info=torch.tensor([[128.0, 192.0, 1.0]]), im_id=torch.ones(1),
                rec_id=torch.ones(1))
    with torch.no_grad():
        got = ex.forward(**feed)
    assert [tuple(o.shape) for o in got[3:]] == outs
    kinds = {e[0] for e in log}
    if name.startswith("retina"):
        assert sorted(e[1] for e in log if e[0] == "gen_proposal_retina") == [8, 16, 32, 64, 128]
        assert [e[2] for e in log if e[0] == "gen_proposal_retina" and e[1] == 128] == [0.0]   # builder.py:372
    elif name.startswith("mask"):
        assert ("fpn_roi_align", (7, 7)) in log and ("fpn_roi_align", (14, 14)) in log and "post" in kinds
    elif name.startswith("cascade"):
        assert [e for e in log if e[0] == "fpn_roi_align"] == [("fpn_roi_align", (7, 7))] * 3 and "proposal" in kinds
    elif name.startswith("tridentnet"):
        assert kinds == {"proposal"} and len(log) == 1          # ONE Proposal call for the three branches
    else:
        assert "dcn" in kinds and "proposal" in kinds
