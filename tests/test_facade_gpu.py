"""Façade executor on the device: the graph that the reference's builders produce for faster_r50v1_fpn_1x (committed
fixture) runs end to end - library conv / GEMM for backbone and heads, the C ABI for every detection operator - and a
small FPN graph written against the `mxnext` stand-in checks the fused / channels-last RoIAlign route against the
literal operator-by-operator route."""
import os

import numpy as np
import pytest
import torch

from simpledet_b200 import facade
from simpledet_b200.facade import mxnext_impl as X
from simpledet_b200.facade import symbol as S

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "golden", "faster_r50v1_fpn_1x_test_symbol.json")


def _mini_fpn_graph():
    """conv stem -> 3 pyramid levels (strides 4, 8, 16) -> RPN (Proposal_v3 per level) -> get_top_proposal ->
    assign_layer_fpn -> 3 x ROIAlign_v2 -> add_n -> fc -> softmax / DecodeBBox, built with the X.* helpers exactly the
    way symbol/builder.py and models/FPN/builder.py use them."""
    data, im_info = X.var("data"), X.var("im_info")
    norm = X.normalizer_factory(type="fixbn")
    c = X.convnormrelu(norm, data, "stem", 16, kernel=3, stride=2)
    c = X.max_pool(c, name="pool0")
    feats = {}
    for s in (4, 8, 16):
        if s > 4:
            c = X.convrelu(c, f"down{s}", 16, kernel=3, stride=2)
        feats[s] = X.conv(c, f"P{s}", 16, kernel=3, no_bias=False)
    props, scores = [], []
    for s, f in feats.items():
        r = X.convrelu(f, f"rpn_conv{s}", 16, kernel=3)
        logit = X.conv(r, f"rpn_cls{s}", 6, no_bias=False)
        delta = X.conv(r, f"rpn_reg{s}", 12, no_bias=False)
        sc = X.sym.SoftmaxActivation(data=X.reshape(logit, (0, 2, -1, 0)), mode="channel")
        sc = X.reshape(sc, (0, 6, -1, 0))
        p, ps = X.contrib.Proposal_v3(cls_prob=sc, bbox_pred=delta, im_info=im_info, rpn_pre_nms_top_n=300,
                                      rpn_post_nms_top_n=100, feature_stride=s, output_score=True, scales=(8,),
                                      ratios=(0.5, 1.0, 2.0), rpn_min_size=0, threshold=0.7, iou_loss=False)
        props.append(p)
        scores.append(ps)
    prop, _ = X.tvm_get_top_proposal(None, bbox=X.concat(props, axis=1), score=X.concat(scores, axis=1), top_n=120)
    lv = X.tvm_fpn_roi_assign(None, prop, (4, 8, 16), 56, 3)
    pooled = [X.roi_align(feats[s], rois=lv[i], out_size=7, stride=s, name=f"roi_align{s}") for i, s in enumerate((4, 8, 16))]
    feat = X.reshape(X.add_n(*pooled, name="roi_sum"), (-3, -2))
    h = X.relu(X.fc(X.flatten(feat), "fc6", 32))
    cls = X.reshape(X.softmax(X.fc(h, "cls", 5), axis=-1), (1, -1, 5))
    box = X.decode_bbox(prop, X.reshape(X.fc(h, "reg", 20), (1, -1, 20)), im_info, bbox_mean=(0, 0, 0, 0),
                        bbox_std=(0.1, 0.1, 0.2, 0.2), class_agnostic=False)
    return X.group([cls, box, feat])


def test_fused_channels_last_route_equals_literal_graph(cuda):
    sym = _mini_fpn_graph()
    shapes = dict(data=(1, 3, 128, 160), im_info=(1, 3))
    data = torch.randn(shapes["data"], device=cuda, generator=torch.Generator(device=cuda).manual_seed(1))
    im_info = torch.tensor([[128.0, 160.0, 1.0]], device=cuda)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    outs = {}
    for key, kw in (("literal", dict(channels_last=False, fuse_fpn_roi_align=False)),
                    ("fused", dict(channels_last=False, fuse_fpn_roi_align=True))):
        ex = facade.Executor(sym, cuda, **kw).init_params(shapes, rng_std=0.05)
        assert (len(ex._fusions) == 1) == kw["fuse_fpn_roi_align"]
        outs[key] = ex.forward(data=data, im_info=im_info)
    for a, b in zip(outs["literal"], outs["fused"]):
        assert torch.equal(a, b)   # same convolutions, and the fused FPN RoIAlign is bit-identical to the 3-level graph
    cl = facade.Executor(sym, cuda, channels_last=True).init_params(shapes, rng_std=0.05).forward(data=data, im_info=im_info)
    assert [tuple(o.shape) for o in cl] == [tuple(o.shape) for o in outs["fused"]]
    assert all(torch.isfinite(o).all() for o in cl)
    assert float(cl[2].abs().sum()) > 0


def test_faster_r50v1_fpn_graph_runs(cuda):
    """The full 551-node graph, zero weights (what detection_infer_speed.py times) and random weights."""
    sym = S.fromjson(open(FIX).read())
    shapes = dict(data=(1, 3, 800, 1333), im_info=(1, 3), im_id=(1,), rec_id=(1,))
    feed = dict(data=torch.ones(shapes["data"], device=cuda), im_info=torch.tensor([[400.0, 666.5, 2.0]], device=cuda),
                im_id=torch.ones(1, device=cuda), rec_id=torch.ones(1, device=cuda))
    for std in (None, 0.02):
        ex = facade.Executor(sym, cuda).init_params(shapes, rng_std=std)
        outs = ex.forward(**feed)
        assert [tuple(o.shape) for o in outs] == [(1,), (1,), (1, 3), (1, 1000, 81), (1, 1000, 324)]
        assert all(torch.isfinite(o).all() for o in outs)
        np.testing.assert_allclose(outs[3].sum(-1).cpu().numpy(), 1.0, rtol=1e-4)   # class scores are a softmax


def test_graph_capture_replays_the_same_outputs(cuda):
    """Executor.capture: the whole mini FPN graph (convolutions + every detection operator) recorded once and replayed
    with new inputs gives what the eager pass gives."""
    sym = _mini_fpn_graph()
    shapes = dict(data=(1, 3, 128, 160), im_info=(1, 3))
    gen = torch.Generator(device=cuda).manual_seed(3)
    ex = facade.Executor(sym, cuda).init_params(shapes, rng_std=0.05)
    im_info = torch.tensor([[128.0, 160.0, 1.0]], device=cuda)
    a = torch.randn(shapes["data"], device=cuda, generator=gen)
    b = torch.randn(shapes["data"], device=cuda, generator=gen)
    with torch.no_grad():
        want = [[o.clone() for o in ex.forward(data=x, im_info=im_info)] for x in (a, b)]
    run = ex.capture(data=a, im_info=im_info)
    for x, w in ((b, want[1]), (a, want[0]), (b, want[1])):
        got = run(data=x, im_info=im_info)
        torch.cuda.synchronize()
        for g, t in zip(got, w):
            assert torch.equal(g, t)


def test_module_executor_group_records_and_replays(cuda):
    """The DetModule-facing executor group: the first forward records the pass into a CUDA graph, later ones replay it;
    set_params drops the recording (and the folded BatchNorm weights).  Outputs equal the eager executor's."""
    from simpledet_b200 import facade as F
    from simpledet_b200.facade import module as M
    from simpledet_b200.facade import ndarray as nd

    mx = F.install()
    sym = _mini_fpn_graph()
    shapes = [nd.DataDesc("data", (1, 3, 128, 160)), nd.DataDesc("im_info", (1, 3))]
    names = [n for n in sym.list_arguments() if n not in ("data", "im_info")]
    grp = M.DataParallelExecutorGroup(sym, [mx.gpu(0)], None, shapes, None, names, False, False)
    gen = torch.Generator(device=cuda).manual_seed(5)
    params = {n: nd.NDArray(torch.randn(grp.exe.params[n].shape, device=cuda, generator=gen) * 0.05) for n in names}
    grp.set_params(params, {})
    im_info = torch.tensor([[128.0, 160.0, 1.0]], device=cuda)
    xs = [torch.randn((1, 3, 128, 160), device=cuda, generator=gen) for _ in range(2)]
    eager = facade.Executor(sym, cuda)
    eager.init_params({"data": (1, 3, 128, 160), "im_info": (1, 3)})
    for n in names:
        eager.params[n].copy_(params[n].t)
    eager._folded = None
    for x in (xs[0], xs[1], xs[0]):
        grp.forward(nd.DataBatch([nd.NDArray(x), nd.NDArray(im_info)]))
        got = [o.t for o in grp.get_outputs()]
        torch.cuda.synchronize()
        with torch.no_grad():
            want = eager.forward(data=x, im_info=im_info)
        for g, w in zip(got, want):
            assert torch.equal(g, w)
    assert grp._run is not None and getattr(grp, "_graph_ok", True)
    grp.set_params(params, {})
    assert grp._run is None


@pytest.mark.parametrize("name,outs", [
    ("retina_r50v1_fpn_1x", [(1, 5000, 81), (1, 5000, 4)]),
    ("mask_r50v1_fpn_1x", [(1, 100, 1), (1, 100, 4), (1, 100, 1), (100, 81, 28, 28), (1,)]),
    ("faster_dcn_r50v1bc4_c5_512roi_1x", [(1, 300, 81), (1, 300, 4)]),
])
def test_other_detectors_of_the_reference_run(cuda, name, outs):
    """BASELINE configs 3-5 as the reference's own builders emit them (fixtures written by
    tests/golden/make_golden_graph.py): RetinaNet (GenAnchor + GenProposalRetina per level), Mask R-CNN (Proposal_v3 x5,
    two fused FPN RoIAlign sizes, BboxPostProcessing, the deconvolution mask head) and DCNv1 Faster R-CNN C4 (legacy
    Proposal, DeformableConvolution in the backbone), 800x1333, random weights."""
    sym = S.fromjson(open(os.path.join(ROOT, "tests", "golden", name + "_test_symbol.json")).read())
    shapes = dict(data=(1, 3, 800, 1333), im_info=(1, 3), im_id=(1,), rec_id=(1,))
    gen = torch.Generator(device=cuda).manual_seed(5)
    feed = dict(data=torch.randn(shapes["data"], device=cuda, generator=gen), im_info=torch.tensor([[800.0, 1333.0, 1.0]], device=cuda),
                im_id=torch.ones(1, device=cuda), rec_id=torch.ones(1, device=cuda))
    ex = facade.Executor(sym, cuda).init_params(shapes, rng_std=0.02)
    with torch.no_grad():
        got = ex.forward(**feed)
    torch.cuda.synchronize()
    assert [tuple(o.shape) for o in got[3:]] == outs
    assert all(torch.isfinite(o).all() for o in got)
    if name.startswith("retina"):
        score, box = got[3], got[4]
        assert float(score.max()) <= 1.0 and float(score.min()) >= 0.0
        kept = score.sum(-1) > 0                                  # rows that passed the score threshold are real boxes
        assert int(kept.sum()) > 0 and bool((box[kept][:, 2] >= box[kept][:, 0] - 1).all())
    elif name.startswith("mask"):
        assert float(got[6].min()) >= 0.0 and float(got[6].max()) <= 1.0     # mask_prob is a sigmoid
    else:
        np.testing.assert_allclose(got[3].sum(-1).cpu().numpy(), 1.0, rtol=1e-4)
