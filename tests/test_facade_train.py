"""Training side of the façade (simpledet_b200/facade/train.py) on the CPU: MXNet's loss-operator gradients against
their formulas written out in numpy, MXNet's SGD update, and the reference's OWN Faster R-CNN FPN / RetinaNet TRAIN
graphs (fixtures written by tests/golden/make_golden_graph.py) run forward + backward through the Trainer with every
detection operator replaced by an argument-checking stub - the plumbing (shape inference incl. label shapes, loss
heads, BlockGrad, fixed parameters, gradient flow into backbone / neck / heads, the flat all-reduce bucket, the
update).  The device run of the same graphs with the real operators is tests/test_zz_late_gpu.py."""
import os

import numpy as np
import pytest
import torch

from simpledet_b200.facade import executor as E
from simpledet_b200.facade import symbol as S
from simpledet_b200.facade import train as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _softmax(z, axis):
    e = np.exp(z - z.max(axis, keepdims=True))
    return e / e.sum(axis, keepdims=True)


def test_softmax_output_gradients():
    rng = np.random.default_rng(0)
    # multi_output, valid normalisation with ignored labels (the RPN classification loss, models/FPN/builder.py:214-223)
    n, k, a, s = 2, 2, 3, 17
    z = rng.standard_normal((n, k, a, s)).astype(np.float32)
    lab = rng.integers(-1, 2, (n, a, s)).astype(np.float32)
    data = torch.tensor(z, requires_grad=True)
    out = T.softmax_output(data, torch.tensor(lab), True, "valid", True, -1, 2.0)
    out.backward(torch.full_like(out, 123.0))                       # the head gradient must not matter
    p = _softmax(z.reshape(n, k, -1), 1)
    l2 = lab.reshape(n, -1).astype(int)
    onehot = np.stack([(l2 == c) for c in range(k)], 1).astype(np.float32)
    want = (p - onehot) * (l2 != -1)[:, None] * (2.0 / max((l2 != -1).sum(), 1))
    np.testing.assert_allclose(data.grad.numpy().reshape(n, k, -1), want, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(out.detach().numpy().reshape(n, k, -1), p, rtol=1e-6)
    # multi_output, batch and null: divided by the spatial size too
    for norm, div in (("batch", a * s * n), ("null", a * s)):
        d = torch.tensor(z, requires_grad=True)
        T.softmax_output(d, torch.tensor(np.abs(lab)), True, norm, False, -1, 1.0).sum().backward()
        oh = np.stack([(np.abs(l2) == c) for c in range(k)], 1).astype(np.float32)
        np.testing.assert_allclose(d.grad.numpy().reshape(n, k, -1), (p - oh) / div, rtol=1e-5, atol=1e-7)
    # 2-D, batch normalisation (the R-CNN classification loss, symbol/builder.py bbox head)
    z2 = rng.standard_normal((12, 7)).astype(np.float32)
    l1 = rng.integers(0, 7, 12).astype(np.float32)
    d = torch.tensor(z2, requires_grad=True)
    T.softmax_output(d, torch.tensor(l1), False, "batch", False, -1, 1.0).sum().backward()
    want = (_softmax(z2, 1) - np.eye(7, dtype=np.float32)[l1.astype(int)]) / 12
    np.testing.assert_allclose(d.grad.numpy(), want, rtol=1e-5, atol=1e-7)


def test_make_loss_smooth_l1_blockgrad():
    x = torch.tensor([[-2.0, -0.05, 0.0, 0.08, 0.4, 3.0]], requires_grad=True)
    y = T.make_loss(T.smooth_l1(x, 3.0), grad_scale=0.25)
    y.backward(torch.full_like(y, 9.0))
    v = x.detach().numpy()
    want_f = np.where(np.abs(v) < 1 / 9, 0.5 * 9 * v * v, np.abs(v) - 0.5 / 9)
    want_g = np.where(np.abs(v) < 1 / 9, 9 * v, np.sign(v)) * 0.25
    np.testing.assert_allclose(y.detach().numpy(), want_f, rtol=1e-6)
    np.testing.assert_allclose(x.grad.numpy(), want_g, rtol=1e-6)
    x2 = torch.tensor([0.5, -1.0, 2.0, 0.0], requires_grad=True)
    T.make_loss(x2, 2.0, "valid", 0.0).backward(torch.ones(4))
    np.testing.assert_allclose(x2.grad.numpy(), np.full(4, 2.0 / 2))       # two entries above valid_thresh


def test_sgd_update_is_mxnets():
    data = S.Variable("data")
    w = S.Variable("fc_weight", lr_mult=2.0, wd_mult=0.5)
    from simpledet_b200.facade import mxnext_impl as X

    out = X.loss(X.fc(data, "fc", 3, weight=w, no_bias=True), grad_scale=1.0, name="l") if hasattr(X, "loss") else None
    assert out is not None
    tr = T.Trainer(out, dict(data=(4, 5)), device="cpu", rng_std=0.1)
    x = torch.arange(20, dtype=torch.float32).reshape(4, 5) / 10
    w0 = tr.ex.params["fc_weight"].detach().clone()
    tr.forward_backward(data=x)
    g = tr.grads()["fc_weight"].clone()
    np.testing.assert_allclose(g.numpy(), np.tile(x.sum(0).numpy(), (3, 1)), rtol=1e-6)   # d(sum of outputs)/dW
    tr.update(lr=0.1, momentum=0.9, wd=0.01, rescale_grad=0.5, clip_gradient=1.5)
    step = np.clip(0.5 * g.numpy(), -1.5, 1.5) + 0.01 * 0.5 * w0.numpy()
    m1 = -0.1 * 2.0 * step
    np.testing.assert_allclose(tr.ex.params["fc_weight"].detach().numpy(), w0.numpy() + m1, rtol=1e-6, atol=1e-7)
    tr.forward_backward(data=x)
    w1 = tr.ex.params["fc_weight"].detach().clone()
    tr.update(lr=0.1, momentum=0.9, wd=0.01, rescale_grad=0.5, clip_gradient=1.5)
    m2 = 0.9 * m1 - 0.1 * 2.0 * (np.clip(0.5 * g.numpy(), -1.5, 1.5) + 0.01 * 0.5 * w1.numpy())
    np.testing.assert_allclose(tr.ex.params["fc_weight"].detach().numpy(), w1.numpy() + m2, rtol=1e-5, atol=1e-7)


# ---- the reference's train graphs with stubbed detection operators -------------------------------------------------
def _train_stubs(monkeypatch, log):
    from simpledet_b200 import ops

    def proposal(cls_prob, bbox_pred, im_info, rpn_pre_nms_top_n, rpn_post_nms_top_n, threshold, rpn_min_size, scales,
                 ratios, feature_stride, output_score, iou_loss):
        B = cls_prob.shape[0]
        assert not cls_prob.requires_grad or True
        log.append(("proposal", feature_stride))
        g = torch.Generator().manual_seed(feature_stride)
        xy = torch.rand(B, rpn_post_nms_top_n, 2, generator=g) * 100
        return torch.cat([xy, xy + 20 + torch.rand(B, rpn_post_nms_top_n, 2, generator=g) * 60], 2), \
            torch.rand(B, rpn_post_nms_top_n, 1, generator=g)

    def get_top_proposal(bbox, score, top_n):
        return bbox[:, :top_n].contiguous(), score[:, :top_n].contiguous()

    def proposal_target(rois, gt_boxes, num_classes, batch_images, image_rois, fg_thresh, bg_thresh_hi, bg_thresh_lo,
                        proposal_without_gt, fg_fraction, class_agnostic, output_iou, bbox_mean, bbox_std, bbox_weight):
        assert not rois.requires_grad and tuple(gt_boxes.shape[::2]) == (batch_images, 5)
        log.append(("proposal_target", image_rois))
        g = torch.Generator().manual_seed(1)
        B, R = batch_images, image_rois
        lab = torch.randint(0, num_classes, (B, R), generator=g).float()
        return (rois[:, :R].contiguous(), lab, torch.randn(B, R, 4 * num_classes, generator=g),
                (torch.rand(B, R, 4 * num_classes, generator=g) < 0.05).float())

    def fpn_roi_align(feats, rois, strides, out_size, scale0, lvl0):
        assert len(feats) == len(strides) == 4 and all(f.requires_grad for f in feats) and not rois.requires_grad
        log.append(("fpn_roi_align", tuple(out_size), rois.shape[1]))
        pooled = sum(torch.nn.functional.adaptive_avg_pool2d(f, out_size) for f in feats)     # (B, C, ph, pw)
        return pooled[:, None].expand(-1, rois.shape[1], -1, -1, -1) * (1 + 0 * rois.sum(-1)[..., None, None, None])

    def focal_loss(data, label, alpha, gamma, normalization, grad_scale):
        assert data.requires_grad and data.shape[:2] == label.shape and normalization == "valid"
        log.append(("focal", tuple(data.shape)))
        return T.make_loss(torch.sigmoid(data), grad_scale)

    def bbox_norm(data, label, normalization):
        log.append(("bbox_norm", tuple(data.shape)))
        return data

    def proposal_mask_target(rois, gt_boxes, gt_polys, num_classes, batch_images, image_rois, mask_size, fg_thresh,
                             bg_thresh_hi, bg_thresh_lo, proposal_without_gt, fg_fraction, class_agnostic, output_iou,
                             output_ratio, bbox_mean, bbox_std, bbox_weight):
        assert output_iou and gt_polys.shape[:2] == gt_boxes.shape[:2]
        base = proposal_target(rois, gt_boxes, num_classes, batch_images, image_rois, fg_thresh, bg_thresh_hi, bg_thresh_lo,
                               proposal_without_gt, fg_fraction, class_agnostic, output_iou, bbox_mean, bbox_std, bbox_weight)
        nfg = int(image_rois * fg_fraction)
        log.append(("proposal_mask_target", nfg, mask_size))
        g = torch.Generator().manual_seed(2)
        tgt = torch.randint(-1, 2, (batch_images, nfg, mask_size, mask_size), generator=g).float()
        out = (*base, torch.rand(batch_images, image_rois, generator=g), tgt)
        if output_ratio:       # Mask Scoring R-CNN: the share of the instance inside the roi
            log.append(("mask_ratio", nfg))
            out += (torch.rand(batch_images, nfg, generator=g).clamp(min=0.05),)
        return out

    def sigmoid_ce(data, label, grad_scale):
        assert data.requires_grad and data.shape == label.shape and data.dim() == 2
        log.append(("sigmoid_ce", tuple(data.shape)))
        keep = (label >= 0).float()
        loss = torch.nn.functional.binary_cross_entropy_with_logits(data, label.clamp(min=0), weight=keep, reduction="none")
        return T.make_loss(loss.sum(1) / keep.sum(1).clamp(min=1), grad_scale)

    def roi_align(data, rois, pooled_size, spatial_scale):
        assert data.requires_grad and rois.shape[0] == data.shape[0] and rois.shape[2] == 4 and 0 < spatial_scale < 1
        log.append(("roi_align", tuple(pooled_size)))
        pooled = torch.nn.functional.adaptive_avg_pool2d(data, tuple(pooled_size))
        return pooled[:, None].expand(-1, rois.shape[1], -1, -1, -1) * (1 + 0 * rois.sum(-1)[..., None, None, None])

    def deform_conv(data, offset, weight, bias, kernel, stride, dilate, pad, num_filter, num_group, num_deformable_group,
                    no_bias):
        assert offset.shape[1] == 2 * num_deformable_group * kernel[0] * kernel[1] and offset.requires_grad
        log.append(("dcn", data.shape[1]))
        return torch.nn.functional.conv2d(data, weight, bias, stride, pad, dilate, num_group) + 0 * offset.sum()

    monkeypatch.setitem(ops.OPS, "_contrib_ROIAlign_v2", roi_align)
    monkeypatch.setitem(ops.OPS, "_contrib_DeformableConvolution", deform_conv)
    monkeypatch.setitem(ops.OPS, "_contrib_Proposal", proposal)
    monkeypatch.setitem(ops.OPS, "ProposalMaskTarget", proposal_mask_target)
    monkeypatch.setitem(ops.OPS, "_contrib_SigmoidCrossEntropy", sigmoid_ce)
    for k, fn in {"_contrib_Proposal_v3": proposal, "get_top_proposal": get_top_proposal, "ProposalTarget": proposal_target,
                  "_contrib_FocalLoss": focal_loss, "_contrib_BBoxNorm": bbox_norm}.items():
        monkeypatch.setitem(ops.OPS, k, fn)
    monkeypatch.setattr(ops, "fpn_roi_align", fpn_roi_align)


def _fpn_label_shapes(h, w, strides, num_anchors):
    s = sum(-(-h // st) * -(-w // st) for st in strides)
    return s, num_anchors


def test_faster_rcnn_fpn_train_graph_through_the_trainer(monkeypatch):
    sym = S.fromjson(open(os.path.join(ROOT, "tests", "golden", "faster_r50v1_fpn_1x_train_symbol.json")).read())
    B, H, W = 2, 128, 192
    shapes = dict(data=(B, 3, H, W), im_info=(B, 3), gt_bbox=(B, 100, 5))
    args, outs, _ = E.infer_shapes(sym, shapes)
    byname = dict(zip(sym.list_arguments(), args))
    s_total = sum((H // st) * (W // st) for st in (4, 8, 16, 32, 64))
    assert byname["rpn_cls_label"] == (B, 3, s_total) and byname["rpn_reg_target"] == (B, 12, s_total) == byname["rpn_reg_weight"]
    assert outs == [(B, 2, 3, s_total), (B, 12, s_total), (B, 3, s_total), (B * 512, 81), (B * 512, 324), (B, 512)]
    log = []
    _train_stubs(monkeypatch, log)
    fixed = ("conv0", "stage1", "gamma", "beta")                                # config/faster_r50v1_fpn_1x.py:147
    labels = ("rpn_cls_label", "rpn_reg_target", "rpn_reg_weight")               # config/faster_r50v1_fpn_1x.py label_name
    tr = T.Trainer(sym, shapes, device="cpu", fixed_param=fixed, rng_std=0.02, label_names=labels)
    with pytest.raises(KeyError):
        tr.ex.forward(data=torch.zeros(shapes["data"]), im_info=torch.ones(B, 3), gt_bbox=torch.zeros(B, 100, 5))
    assert not any(any(f in n for f in fixed) for n in tr.trainable) and not set(tr.trainable) & (set(shapes) | set(labels))
    assert "stage2_unit1_conv1_weight" in tr.trainable and "bbox_cls_logit_weight" in tr.trainable
    g = torch.Generator().manual_seed(0)
    feed = dict(data=torch.randn(shapes["data"], generator=g), im_info=torch.tensor([[H, W, 1.0]] * B),
                gt_bbox=torch.full((B, 100, 5), -1.0),
                rpn_cls_label=torch.randint(-1, 2, (B, 3, s_total), generator=g).float(),
                rpn_reg_target=torch.randn(B, 12, s_total, generator=g),
                rpn_reg_weight=(torch.rand(B, 12, s_total, generator=g) < 0.1).float())
    res = tr.forward_backward(**feed)
    assert [tuple(o.shape) for o in res] == outs and all(torch.isfinite(o).all() for o in res)
    assert sorted(e[1] for e in log if e[0] == "proposal") == [4, 8, 16, 32, 64]
    assert ("proposal_target", 512) in log and ("fpn_roi_align", (7, 7), 512) in log
    grads = tr.grads()
    assert set(grads) == set(tr.trainable)                                      # every trainable parameter got one
    for name in ("stage2_unit1_conv1_weight", "stage4_unit3_conv3_weight", "P2_lateral_weight", "rpn_conv_weight",
                 "rpn_conv_cls_weight", "rpn_conv_bbox_weight", "bbox_fc1_weight", "bbox_cls_logit_weight",
                 "bbox_reg_delta_weight"):
        assert name in grads and float(grads[name].abs().sum()) > 0, name
    assert tr.ex.params["conv0_weight"].grad is None and tr.ex.params["stage2_unit1_conv1_bn_gamma"].grad is None
    # the label outputs are BlockGrad heads; the RPN classification output is the softmax
    assert not res[2].requires_grad and torch.allclose(res[0].sum(1), torch.ones(B, 3, s_total), atol=1e-5)
    before = tr.ex.params["bbox_fc1_weight"].detach().clone()
    assert tr.allreduce_grads() is None                                         # no process group: a no-op
    tr.update(lr=0.01, momentum=0.9, wd=1e-4, rescale_grad=1.0)
    assert not torch.equal(before, tr.ex.params["bbox_fc1_weight"].detach())


def test_retinanet_train_graph_through_the_trainer(monkeypatch):
    sym = S.fromjson(open(os.path.join(ROOT, "tests", "golden", "retina_r50v1_fpn_1x_train_symbol.json")).read())
    B, H, W = 2, 128, 192
    shapes = dict(data=(B, 3, H, W))
    args, outs, _ = E.infer_shapes(sym, shapes)
    byname = dict(zip(sym.list_arguments(), args))
    labels = {n: s for n, s in byname.items() if not n.endswith(("_weight", "_bias", "_gamma", "_beta")) and n != "data"}
    labels.update({n: s for n, s in byname.items() if n in ("rpn_reg_weight",)})
    assert all(s is not None for s in labels.values()), labels
    log = []
    _train_stubs(monkeypatch, log)
    tr = T.Trainer(sym, shapes, device="cpu", fixed_param=("conv0", "stage1", "gamma", "beta"), rng_std=0.02,
                   label_names=tuple(labels))
    g = torch.Generator().manual_seed(0)
    feed = dict(data=torch.randn(shapes["data"], generator=g))
    for n, s in labels.items():
        feed[n] = (torch.rand(s, generator=g) < 0.05).float()
    res = tr.forward_backward(**feed)
    assert [tuple(o.shape) for o in res] == outs
    assert any(e[0] == "focal" for e in log) and any(e[0] == "bbox_norm" for e in log)
    grads = tr.grads()
    assert set(grads) == set(tr.trainable)
    assert sum(float(v.abs().sum()) > 0 for v in grads.values()) > 0.9 * len(grads)


def test_mask_rcnn_train_graph_through_the_trainer(monkeypatch):
    """models/maskrcnn/builder.py:184-320: ProposalMaskTarget's six outputs, the first 128 rois of each image into the
    14x14 RoIAlign, split / arange / stack / gather_nd picking each roi's class channel, SigmoidCrossEntropy."""
    sym = S.fromjson(open(os.path.join(ROOT, "tests", "golden", "mask_r50v1_fpn_1x_train_symbol.json")).read())
    B, H, W = 2, 128, 192
    shapes = dict(data=(B, 3, H, W), im_info=(B, 3), gt_bbox=(B, 100, 5), gt_poly=(B, 100, 2500))
    _, outs, _ = E.infer_shapes(sym, shapes)
    assert outs[-1] == (1,) and outs[3] == (B * 512, 81)
    log = []
    _train_stubs(monkeypatch, log)
    labels = ("rpn_cls_label", "rpn_reg_target", "rpn_reg_weight")
    tr = T.Trainer(sym, shapes, device="cpu", fixed_param=("conv0", "stage1", "gamma", "beta"), rng_std=0.02,
                   label_names=labels)
    s_total = sum((H // st) * (W // st) for st in (4, 8, 16, 32, 64))
    g = torch.Generator().manual_seed(0)
    feed = dict(data=torch.randn(shapes["data"], generator=g), im_info=torch.tensor([[H, W, 1.0]] * B),
                gt_bbox=torch.full((B, 100, 5), -1.0), gt_poly=torch.full((B, 100, 2500), -1.0),
                rpn_cls_label=torch.randint(-1, 2, (B, 3, s_total), generator=g).float(),
                rpn_reg_target=torch.randn(B, 12, s_total, generator=g),
                rpn_reg_weight=(torch.rand(B, 12, s_total, generator=g) < 0.1).float())
    res = tr.forward_backward(**feed)
    assert [tuple(o.shape) for o in res] == outs and all(torch.isfinite(o).all() for o in res)
    assert ("proposal_mask_target", 128, 28) in log and ("sigmoid_ce", (1, B * 128 * 28 * 28)) in log
    assert ("fpn_roi_align", (7, 7), 512) in log and ("fpn_roi_align", (14, 14), 128) in log
    grads = tr.grads()
    assert set(grads) == set(tr.trainable)
    for name in ("mask_fcn_logit_weight", "bbox_fc1_weight", "P2_lateral_weight", "stage3_unit2_conv2_weight"):
        assert float(grads[name].abs().sum()) > 0, name


def test_dcn_c4_train_graph_through_the_trainer(monkeypatch):
    """config/dcn/faster_dcn_r50v1bc4_c5_512roi_1x.py: legacy Proposal, a single-level ROIAlign_v2 through the
    autograd operator, three DeformableConvolution blocks whose offset branches must receive gradients too."""
    sym = S.fromjson(open(os.path.join(ROOT, "tests", "golden", "faster_dcn_r50v1bc4_c5_512roi_1x_train_symbol.json")).read())
    for node in sym._topo():           # 512 rois per image through a ResNet stage is minutes on a CPU: sample 24
        if node.op == "ProposalTarget":
            node.attrs["image_rois"] = 24
    B, H, W = 2, 128, 192
    shapes = dict(data=(B, 3, H, W), im_info=(B, 3), gt_bbox=(B, 100, 5))
    args, outs, _ = E.infer_shapes(sym, shapes)
    byname = dict(zip(sym.list_arguments(), args))
    assert outs[2:] == [(B * 24, 81), (B * 24, 8), (B, 24)]
    log = []
    _train_stubs(monkeypatch, log)
    labels = ("rpn_cls_label", "rpn_reg_target", "rpn_reg_weight")
    tr = T.Trainer(sym, shapes, device="cpu", fixed_param=("conv0", "stage1", "gamma", "beta"), rng_std=0.02,
                   label_names=labels)
    g = torch.Generator().manual_seed(0)
    feed = dict(data=torch.randn(shapes["data"], generator=g), im_info=torch.tensor([[H, W, 1.0]] * B),
                gt_bbox=torch.full((B, 100, 5), -1.0),
                rpn_cls_label=torch.randint(-1, 2, byname["rpn_cls_label"], generator=g).float(),
                rpn_reg_target=torch.randn(byname["rpn_reg_target"], generator=g),
                rpn_reg_weight=(torch.rand(byname["rpn_reg_weight"], generator=g) < 0.1).float())
    res = tr.forward_backward(**feed)
    assert [tuple(o.shape) for o in res] == outs and all(torch.isfinite(o).all() for o in res)
    kinds = [e[0] for e in log]
    assert kinds.count("dcn") == 3 and "roi_align" in kinds and "proposal" in kinds and "proposal_target" in kinds
    grads = tr.grads()
    assert set(grads) == set(tr.trainable)
    offs = [n for n in grads if "offset" in n and n.endswith("weight")]
    assert len(offs) == 3 and all(n in grads for n in offs)


def test_crowdhuman_train_graph_hands_bbox_target_its_attributes(monkeypatch):
    """config/crowdhuman/faster_r50v1b_fpn_1x.py: mx.sym.Custom(op_type='bbox_target') (models/crowdhuman/builder.py:
    380-396) reaches OPS['bbox_target'] with the attributes the CustomOpProp would parse."""
    from simpledet_b200 import ops

    sym = S.fromjson(open(os.path.join(ROOT, "tests", "golden", "crowdhuman_faster_r50v1b_fpn_1x_train_symbol.json")).read())
    B, H, W = 2, 128, 192
    shapes = dict(data=(B, 3, H, W), im_info=(B, 3), gt_bbox=(B, 100, 5))
    _, outs, _ = E.infer_shapes(sym, shapes)
    assert outs[3:] == [(B * 512, 2), (B * 512, 8), (B, 512)]
    log = []
    _train_stubs(monkeypatch, log)

    def bbox_target(proposal, gt_bbox, num_class, add_gt_to_proposal, image_rois, fg_fraction, fg_thresh, bg_thresh_hi,
                    bg_thresh_lo, bbox_target_std):
        assert (num_class, add_gt_to_proposal, image_rois, fg_fraction) == (2, True, 512, 0.5)
        assert (fg_thresh, bg_thresh_hi, bg_thresh_lo, bbox_target_std) == (0.5, 0.5, 0.0, (0.1, 0.1, 0.2, 0.2))
        assert tuple(gt_bbox.shape) == (B, 100, 5) and proposal.shape[2] == 4
        log.append(("bbox_target", image_rois))
        g = torch.Generator().manual_seed(5)
        return (proposal[:, :image_rois].contiguous(), torch.randint(0, 2, (B, image_rois), generator=g).float(),
                torch.randn(B, image_rois, 8, generator=g), (torch.rand(B, image_rois, 8, generator=g) < 0.2).float())

    monkeypatch.setitem(ops.OPS, "bbox_target", bbox_target)
    tr = T.Trainer(sym, shapes, device="cpu", fixed_param=("conv0", "stage1", "gamma", "beta"), rng_std=0.02,
                   label_names=("rpn_cls_label", "rpn_reg_target", "rpn_reg_weight"))
    s_total = sum((H // st) * (W // st) for st in (4, 8, 16, 32, 64))
    g = torch.Generator().manual_seed(0)
    res = tr.forward_backward(data=torch.randn(shapes["data"], generator=g), im_info=torch.tensor([[H, W, 1.0]] * B),
                              gt_bbox=torch.full((B, 100, 5), -1.0),
                              rpn_cls_label=torch.randint(-1, 2, (B, 3, s_total), generator=g).float(),
                              rpn_reg_target=torch.randn(B, 12, s_total, generator=g),
                              rpn_reg_weight=(torch.rand(B, 12, s_total, generator=g) < 0.1).float())
    assert [tuple(o.shape) for o in res] == outs and ("bbox_target", 512) in log
    assert set(tr.grads()) == set(tr.trainable)


def test_cascade_rcnn_train_graph_through_the_trainer(monkeypatch):
    """config/cascade_r50v1_fpn_1x.py: three ProposalTarget -> fused RoIAlign -> head -> DecodeBBox stages, nine loss
    heads; every stage's head parameters must receive gradients."""
    sym = S.fromjson(open(os.path.join(ROOT, "tests", "golden", "cascade_r50v1_fpn_1x_train_symbol.json")).read())
    for node in sym._topo():               # 3 x 512 rois per image through two 1024-wide FC layers: keep the CPU run short
        if node.op == "ProposalTarget":
            node.attrs["image_rois"] = 64
    B, H, W = 2, 128, 192
    shapes = dict(data=(B, 3, H, W), im_info=(B, 3), gt_bbox=(B, 100, 5))
    _, outs, _ = E.infer_shapes(sym, shapes)
    assert len(outs) == 12 and outs[3:6] == outs[6:9] == outs[9:12] == [(B * 64, 81), (B * 64, 8), (B, 64)]
    log = []
    _train_stubs(monkeypatch, log)

    def decode_bbox(rois, bbox_pred, im_info, mean, std, class_agnostic):
        assert class_agnostic and rois.shape[:2] == bbox_pred.shape[:2] and not rois.requires_grad
        log.append(("decode", tuple(std)))
        return rois.detach().clone()

    from simpledet_b200 import ops
    monkeypatch.setitem(ops.OPS, "_contrib_DecodeBBox", decode_bbox)
    tr = T.Trainer(sym, shapes, device="cpu", fixed_param=("conv0", "stage1", "gamma", "beta"), rng_std=0.02,
                   label_names=("rpn_cls_label", "rpn_reg_target", "rpn_reg_weight"))
    s_total = sum((H // st) * (W // st) for st in (4, 8, 16, 32, 64))
    g = torch.Generator().manual_seed(0)
    res = tr.forward_backward(data=torch.randn(shapes["data"], generator=g), im_info=torch.tensor([[H, W, 1.0]] * B),
                              gt_bbox=torch.full((B, 100, 5), -1.0),
                              rpn_cls_label=torch.randint(-1, 2, (B, 3, s_total), generator=g).float(),
                              rpn_reg_target=torch.randn(B, 12, s_total, generator=g),
                              rpn_reg_weight=(torch.rand(B, 12, s_total, generator=g) < 0.1).float())
    assert [tuple(o.shape) for o in res] == outs
    assert [e[1] for e in log if e[0] == "proposal_target"] == [64, 64, 64]
    assert len([e for e in log if e[0] == "fpn_roi_align"]) == 3 and len([e for e in log if e[0] == "decode"]) == 2
    grads = tr.grads()
    assert set(grads) == set(tr.trainable)
    heads = [n for n in grads if n.startswith("bbox_") and n.endswith("weight")]
    assert len(heads) >= 12 and all(float(grads[n].abs().sum()) > 0 for n in heads), heads


def test_tridentnet_train_graph_reaches_proposal_v2_and_proposal_target_v2(monkeypatch):
    """config/tridentnet_r50v1c4_c5_1x.py: three branches in the batch axis, `_contrib_Proposal_v2` and
    `ProposalTarget_v2` with valid_ranges / filter_scales (models/tridentnet/builder.py:239, :377-398); the shared
    moving statistics handed to BatchNorm as variables are auxiliary states, not trainable arguments."""
    from simpledet_b200 import ops

    sym = S.fromjson(open(os.path.join(ROOT, "tests", "golden", "tridentnet_r50v1c4_c5_1x_train_symbol.json")).read())
    assert len(sym.list_auxiliary_states()) == 106 and not any("moving" in n for n in sym.list_arguments())
    for node in sym._topo():
        if node.op == "ProposalTarget_v2":
            node.attrs["image_rois"] = 16
    B, H, W, NB = 2, 128, 192, 3
    shapes = dict(data=(B, 3, H, W), im_info=(B, 3), gt_bbox=(B, 100, 5), valid_ranges=(B, NB, 2),
                  rpn_cls_label=(B, NB, 15, 8, 12), rpn_reg_target=(B, NB, 60, 8, 12), rpn_reg_weight=(B, NB, 60, 8, 12))
    _, outs, _ = E.infer_shapes(sym, shapes)
    assert outs == [(B * NB, 2, 15, 8, 12), (B * NB, 60, 8, 12), (B * NB * 16, 81), (B * NB * 16, 8), (B * NB, 16)]
    log = []
    _train_stubs(monkeypatch, log)

    def proposal_v2(cls_prob, bbox_pred, im_info, valid_ranges, rpn_pre_nms_top_n, rpn_post_nms_top_n, threshold,
                    rpn_min_size, scales, ratios, feature_stride, output_score, iou_loss, filter_scales):
        n = cls_prob.shape[0]
        assert n == B * NB and tuple(valid_ranges.shape) == (n, 2) and tuple(im_info.shape) == (n, 3) and filter_scales
        log.append(("proposal_v2", rpn_post_nms_top_n))
        g = torch.Generator().manual_seed(3)
        xy = torch.rand(n, rpn_post_nms_top_n, 2, generator=g) * 100
        return torch.cat([xy, xy + 30], 2), torch.rand(n, rpn_post_nms_top_n, 1, generator=g)

    def proposal_target_v2(rois, gt_boxes, valid_ranges, num_classes, batch_images, image_rois, fg_thresh, bg_thresh_hi,
                           bg_thresh_lo, proposal_without_gt, fg_fraction, class_agnostic, output_iou, bbox_mean, bbox_std,
                           bbox_weight, filter_scales):
        assert batch_images == B * NB and tuple(valid_ranges.shape) == (B * NB, 2) and filter_scales and class_agnostic
        assert tuple(gt_boxes.shape) == (B * NB, 100, 5)
        log.append(("proposal_target_v2", image_rois))
        g = torch.Generator().manual_seed(4)
        return (rois[:, :image_rois].contiguous(), torch.randint(0, 81, (batch_images, image_rois), generator=g).float(),
                torch.randn(batch_images, image_rois, 8, generator=g), (torch.rand(batch_images, image_rois, 8, generator=g) < 0.3).float())

    monkeypatch.setitem(ops.OPS, "_contrib_Proposal_v2", proposal_v2)
    monkeypatch.setitem(ops.OPS, "ProposalTarget_v2", proposal_target_v2)
    labels = ("rpn_cls_label", "rpn_reg_target", "rpn_reg_weight")
    data_shapes = {k: v for k, v in shapes.items() if k not in labels}
    tr = T.Trainer(sym, shapes, device="cpu", fixed_param=("conv0", "stage1", "gamma", "beta"), rng_std=0.02)
    assert not set(tr.trainable) & set(shapes)
    g = torch.Generator().manual_seed(0)
    feed = dict(data=torch.randn(shapes["data"], generator=g), im_info=torch.tensor([[H, W, 1.0]] * B),
                gt_bbox=torch.full((B, 100, 5), -1.0), valid_ranges=torch.tensor([[[0.0, 90.0], [30.0, 160.0], [90.0, -1.0]]] * B),
                rpn_cls_label=torch.randint(-1, 2, shapes["rpn_cls_label"], generator=g).float(),
                rpn_reg_target=torch.randn(shapes["rpn_reg_target"], generator=g),
                rpn_reg_weight=(torch.rand(shapes["rpn_reg_weight"], generator=g) < 0.1).float())
    res = tr.forward_backward(**feed)
    assert [tuple(o.shape) for o in res] == outs and ("proposal_v2", 500) in log and ("proposal_target_v2", 16) in log
    grads = tr.grads()
    assert set(grads) == set(tr.trainable) and len(data_shapes) == 4


def test_mask_scoring_rcnn_train_graph_through_the_trainer(monkeypatch):
    """config/ms_r50v1_fpn_1x.py: ProposalMaskTarget(output_ratio=True) -> CustomOp 'maskiou_compute' (the real
    host-composed operator, CUDA tensor check lifted) -> the MaskIoU head's L2 loss written with sum / maximum / **."""
    from simpledet_b200 import ops

    sym = S.fromjson(open(os.path.join(ROOT, "tests", "golden", "ms_r50v1_fpn_1x_train_symbol.json")).read())
    B, H, W = 2, 128, 192
    shapes = dict(data=(B, 3, H, W), im_info=(B, 3), gt_bbox=(B, 100, 5), gt_poly=(B, 100, 2500))
    _, outs, _ = E.infer_shapes(sym, shapes)
    assert outs[-2:] == [(1,), (1,)]
    log = []
    _train_stubs(monkeypatch, log)
    monkeypatch.setattr(ops, "_dev", lambda t, name, dtype=torch.float32: t.contiguous())
    tr = T.Trainer(sym, shapes, device="cpu", fixed_param=("conv0", "stage1", "gamma", "beta"), rng_std=0.02,
                   label_names=("rpn_cls_label", "rpn_reg_target", "rpn_reg_weight"))
    s_total = sum((H // st) * (W // st) for st in (4, 8, 16, 32, 64))
    g = torch.Generator().manual_seed(0)
    res = tr.forward_backward(data=torch.randn(shapes["data"], generator=g), im_info=torch.tensor([[H, W, 1.0]] * B),
                              gt_bbox=torch.full((B, 100, 5), -1.0), gt_poly=torch.full((B, 100, 2500), -1.0),
                              rpn_cls_label=torch.randint(-1, 2, (B, 3, s_total), generator=g).float(),
                              rpn_reg_target=torch.randn(B, 12, s_total, generator=g),
                              rpn_reg_weight=(torch.rand(B, 12, s_total, generator=g) < 0.1).float())
    assert [tuple(o.shape) for o in res] == outs and all(torch.isfinite(o).all() for o in res)
    assert ("mask_ratio", 128) in log and float(res[-1]) >= 0
    grads = tr.grads()
    assert set(grads) == set(tr.trainable)
    iou_head = [n for n in grads if "iou" in n and n.endswith("weight")]
    assert iou_head and all(float(grads[n].abs().sum()) > 0 for n in iou_head), iou_head
