"""Shape inference and torch execution of façade graphs.

Generic NN operators run as library calls (torch -> cuDNN / cuBLAS: the tensor-core part of the model, not this
repository's product); detection operators are dispatched BY THE REFERENCE'S REGISTRATION STRING to
`simpledet_b200.ops.OPS`, i.e. to the C ABI.  Parameters the caller did not supply are zero-initialised, which is what
`detection_infer_speed.py` times (`mod.set_params({}, {}, True)`, core/detection_module.py:374-383)."""
from __future__ import annotations

import ast
import math

from . import symbol as S


def _t(v):
    if isinstance(v, str):
        try:
            return ast.literal_eval(v)
        except (ValueError, SyntaxError):
            return v
    return v


def _tup(v, n=2):
    v = _t(v)
    return tuple(int(x) for x in v) if isinstance(v, (tuple, list)) else (int(v),) * n


def _b(v):
    return _t(v) in (True, 1, "True", "true")


def mx_reshape(shape, target):
    """MXNet Reshape with the special codes 0 (copy), -1 (infer), -2 (copy the rest), -3 (merge two), -4 (split)."""
    shape, target = list(shape), list(_t(target))
    out, i, j, infer = [], 0, 0, None
    while j < len(target):
        t = target[j]
        if t == 0:
            out.append(shape[i]); i += 1
        elif t == -1:
            infer = len(out); out.append(-1); i += 1
        elif t == -2:
            out.extend(shape[i:]); i = len(shape)
        elif t == -3:
            out.append(shape[i] * shape[i + 1]); i += 2
        elif t == -4:
            a, b = target[j + 1], target[j + 2]
            if a == -1:
                a = shape[i] // b
            if b == -1:
                b = shape[i] // a
            out.extend([a, b]); i += 1; j += 2
        else:
            out.append(t); i += 1
        j += 1
    if infer is not None:
        total = math.prod(shape)
        known = math.prod(x for x in out if x != -1)
        out[infer] = total // max(known, 1)
    return tuple(out)


def _conv_out(x, k, s, p, d):
    return (x + 2 * p - (d * (k - 1) + 1)) // s + 1


# ------------------------------------------------------------------------------------------------
# shape rules: fn(attrs, in_shapes: list[tuple|None], arg_names) -> (in_shapes completed, out_shapes)
# ------------------------------------------------------------------------------------------------
def _shape_conv(a, ins, names):
    x = ins[0]
    k, s, p, d = _tup(a["kernel"]), _tup(a.get("stride", 1)), _tup(a.get("pad", 0)), _tup(a.get("dilate", 1))
    nf, g = int(_t(a["num_filter"])), int(_t(a.get("num_group", 1)))
    ins = list(ins)
    for i, n in enumerate(names):
        if n == "weight":
            ins[i] = (nf, x[1] // g, k[0], k[1])
        elif n == "bias":
            ins[i] = (nf,)
    return ins, [(x[0], nf, _conv_out(x[2], k[0], s[0], p[0], d[0]), _conv_out(x[3], k[1], s[1], p[1], d[1]))]


def _shape_deconv(a, ins, names):
    x = ins[0]
    k, s, p = _tup(a["kernel"]), _tup(a.get("stride", 1)), _tup(a.get("pad", 0))
    adj = _tup(a.get("adj", 0))
    nf, g = int(_t(a["num_filter"])), int(_t(a.get("num_group", 1)))
    ins = list(ins)
    for i, n in enumerate(names):
        if n == "weight":
            ins[i] = (x[1], nf // g, k[0], k[1])     # MXNet Deconvolution: (in_channels, num_filter / group, kh, kw)
        elif n == "bias":
            ins[i] = (nf,)
    o = lambda n, j: (n - 1) * s[j] - 2 * p[j] + k[j] + adj[j]
    return ins, [(x[0], nf, o(x[2], 0), o(x[3], 1))]


def _shape_deform_conv(a, ins, names):
    ins2, outs = _shape_conv(a, ins, names)          # data, offset, weight[, bias]: the output is a convolution's
    k, dg = _tup(a["kernel"]), int(_t(a.get("num_deformable_group", 1)))
    o = outs[0]
    ins2[1] = (o[0], 2 * dg * k[0] * k[1], o[2], o[3])
    return ins2, outs


def _shape_gen_anchor(a, ins, names):
    c = ins[0]                                       # cls_prob (B, A*K, H, W): generate_anchor-inl.h InferShape
    A = len(_t(a["scales"])) * len(_t(a["ratios"]))
    return ins, [(c[2] * c[3] * A, 4)]


def _shape_gen_proposal_retina(a, ins, names):
    c = ins[0]
    A, pre = int(_t(a["num_anchors"])), int(_t(a.get("rpn_pre_nms_top_n", 6000)))
    oc = c[1] // A + 1 if _b(a.get("output_one_hot", True)) else 1
    return ins, [(c[0], pre, 4), (c[0], pre, oc)]


def _shape_fc(a, ins, names):
    x = ins[0]
    nh = int(_t(a["num_hidden"]))
    flat = _b(a.get("flatten", True))
    k = math.prod(x[1:]) if flat else x[-1]
    ins = list(ins)
    for i, n in enumerate(names):
        if n == "weight":
            ins[i] = (nh, k)
        elif n == "bias":
            ins[i] = (nh,)
    return ins, [(x[0], nh) if flat else tuple(x[:-1]) + (nh,)]


def _shape_bn(a, ins, names):
    c = ins[0][1]
    return [ins[0]] + [(c,)] * (len(ins) - 1), [ins[0]]


def _shape_pool(a, ins, names):
    x = ins[0]
    if _b(a.get("global_pool", False)):
        return ins, [(x[0], x[1], 1, 1)]
    k, s, p = _tup(a["kernel"]), _tup(a.get("stride", 1)), _tup(a.get("pad", 0))
    full = str(a.get("pooling_convention", "valid")) == "full"
    def o(n, kk, ss, pp):
        v = (n + 2 * pp - kk)
        return (-(-v // ss) if full else v // ss) + 1
    return ins, [(x[0], x[1], o(x[2], k[0], s[0], p[0]), o(x[3], k[1], s[1], p[1]))]


def _same(a, ins, names):
    return ins, [ins[0]]


def _shape_elemwise(a, ins, names):
    """Binary elementwise: an input whose shape is not known yet (a label variable) takes the other's."""
    known = next((i for i in ins if i is not None), None)
    ins = [known if i is None else i for i in ins]
    return ins, [known]


def _shape_softmax_output(a, ins, names):
    d = ins[0]
    if len(ins) > 1 and ins[1] is None:  # SoftmaxOutputProp::InferShape: the label's shape follows from the data's
        ins = [d, (d[0],) + tuple(d[2:]) if _b(a.get("multi_output", False)) else (d[0],)] + list(ins[2:])
    return ins, [d]


def _shape_focal_loss(a, ins, names):
    d = ins[0]
    if len(ins) > 1 and ins[1] is None:  # focal_loss-inl.h InferShape: label (batch, anchors) from data (batch, anchors, classes)
        ins = [d, (d[0], d[1])] + list(ins[2:])
    return ins, [d]


def _shape_proposal_mask_target(a, ins, names):
    B, R, C = int(_t(a["batch_images"])), int(_t(a["image_rois"])), int(_t(a["num_classes"]))
    M, nfg = int(_t(a["mask_size"])), int(int(_t(a["image_rois"])) * float(_t(a.get("fg_fraction", 0.25))))
    outs = [(B, R, 4), (B, R), (B, R, 4 * C), (B, R, 4 * C)]
    if _b(a.get("output_iou", False)):
        outs.append((B, R))
    outs.append((B, nfg, M, M))
    if _b(a.get("output_ratio", False)):
        outs.append((B, nfg))
    return ins, outs


def _shape_split(a, ins, names):
    n, ax = int(_t(a["num_outputs"])), int(_t(a.get("axis", 1)))
    d = list(ins[0])
    ax %= len(d)
    d[ax] //= n
    if _b(a.get("squeeze_axis", False)):
        d.pop(ax)
    return ins, [tuple(d)] * n


def _arange_len(a):
    start, stop = float(_t(a.get("start", 0))), a.get("stop")
    stop = None if stop in (None, "None") else float(_t(stop))
    if stop is None:
        start, stop = 0.0, start
    return start, stop, float(_t(a.get("step", 1.0))), int(_t(a.get("repeat", 1)))


def _shape_arange(a, ins, names):
    start, stop, step, rep = _arange_len(a)
    return ins, [(max(0, math.ceil((stop - start) / step)) * rep,)]


def _shape_sum(a, ins, names):
    ax = a.get("axis")
    d = list(ins[0])
    if ax in (None, "None", ()):
        return ins, [(1,)]
    axes = [int(v) % len(d) for v in (_t(ax) if isinstance(_t(ax), (tuple, list)) else (_t(ax),))]
    keep = _b(a.get("keepdims", False))
    out = [1 if i in axes else v for i, v in enumerate(d)] if keep else [v for i, v in enumerate(d) if i not in axes]
    return ins, [tuple(out) or (1,)]


def _shape_expand_dims(a, ins, names):
    d = list(ins[0])
    ax = int(_t(a["axis"]))
    d.insert(ax if ax >= 0 else len(d) + 1 + ax, 1)
    return ins, [tuple(d)]


def _shape_stack(a, ins, names):
    ax = int(_t(a.get("axis", 0)))
    d = list(ins[0])
    d.insert(ax % (len(d) + 1), len(ins))
    return ins, [tuple(d)]


def _shape_gather_nd(a, ins, names):
    d, i = ins
    return ins, [tuple(i[1:]) + tuple(d[i[0]:])]


def _shape_proposal_target(a, ins, names):
    B, R, C = int(_t(a["batch_images"])), int(_t(a["image_rois"])), int(_t(a["num_classes"]))
    outs = [(B, R, 4), (B, R), (B, R, 4 * C), (B, R, 4 * C)]
    if _b(a.get("output_iou", False)):
        outs.append((B, R))
    return ins, outs


def _shape_concat(a, ins, names):
    d = int(_t(a.get("dim", 1)))
    out = list(ins[0])
    out[d] = sum(s[d] for s in ins)
    return ins, [tuple(out)]


def _shape_upsample(a, ins, names):
    x, sc = ins[0], int(_t(a["scale"]))
    return ins, [(x[0], x[1], x[2] * sc, x[3] * sc)]


def _shape_slice_like(a, ins, names):
    axes = _t(a.get("axes", ()))
    out = list(ins[0])
    for ax in (axes if axes else range(len(out))):
        out[ax] = ins[1][ax]
    return ins, [tuple(out)]


def _shape_slice_axis(a, ins, names):
    ax, b, e = int(_t(a["axis"])), int(_t(a["begin"])), _t(a["end"])
    out = list(ins[0])
    e = out[ax] if e in (None, "None") else int(e)
    if e < 0:
        e += out[ax]
    out[ax] = e - b
    return ins, [tuple(out)]


def _shape_proposal(a, ins, names):
    B = ins[0][0]
    post = int(_t(a.get("rpn_post_nms_top_n", 300)))
    return ins, [(B, post, 4), (B, post, 1)]


def _shape_roialign(a, ins, names):
    d, r = ins[0], ins[1]
    ph, pw = _tup(a["pooled_size"])
    s = (r[0], r[1], d[1], ph, pw)
    return ins, [s, s, s]


def _shape_decode(a, ins, names):
    r, d = ins[0], ins[1]
    return ins, [(r[0], r[1], 4) if _b(a.get("class_agnostic", True)) else tuple(d)]


def _shape_custom(a, ins, names):
    t = a.get("op_type")
    if t == "get_top_proposal":
        B, n = ins[0][0], int(_t(a["top_n"]))
        return ins, [(B, n, 4), (B, n, 1)]
    if t == "assign_layer_fpn":
        return ins, [ins[0]] * len(_t(a["rcnn_stride"]))
    if t == "BboxPostProcessing":
        B, m = ins[0][0], int(_t(a["max_det_per_image"]))
        return ins, [(B, m, 1), (B, m, 4), (B, m, 1)]
    if t == "maskiou_compute":
        return ins, [(ins[1][0], 1), (ins[1][0], 1)]
    if t == "bbox_target":
        B, R, C = ins[0][0], int(_t(a["image_rois"])), int(_t(a["num_class"]))
        return ins, [(B, R, 4), (B, R), (B, R, 4 * C), (B, R, 4 * C)]
    if t == "decode_retina":
        n_in = len(ins)
        raise NotImplementedError("decode_retina through the symbol graph (the builder prefers GenProposalRetina)")
    raise NotImplementedError(f"Custom op_type {t!r}")


SHAPE_RULES = {
    "Convolution": _shape_conv, "FullyConnected": _shape_fc, "BatchNorm": _shape_bn, "Pooling": _shape_pool,
    "Activation": _same, "relu": _same, "Cast": _same, "BlockGrad": _same, "softmax": _same, "SoftmaxActivation": _same,
    "SoftmaxOutput": _shape_softmax_output, "Dropout": _same, "MakeLoss": _same, "smooth_l1": _same, "identity": _same,
    "_plus_scalar": _same, "_minus_scalar": _same, "_mul_scalar": _same, "_div_scalar": _same,
    "elemwise_add": _shape_elemwise, "elemwise_sub": _shape_elemwise, "elemwise_mul": _shape_elemwise,
    "elemwise_div": _shape_elemwise, "add_n": _same, "ProposalTarget": _shape_proposal_target,
    "_contrib_FocalLoss": _shape_focal_loss, "_contrib_BBoxNorm": _same,
    "ProposalMaskTarget": _shape_proposal_mask_target, "_contrib_SigmoidCrossEntropy": lambda a, ins, n: (ins, [(ins[0][0],)]),
    "sum": _shape_sum, "expand_dims": _shape_expand_dims, "_power_scalar": _same, "maximum": _shape_elemwise,
    "minimum": _shape_elemwise, "split": _shape_split, "SliceChannel": _shape_split, "arange": _shape_arange, "_arange": _shape_arange,
    "stack": _shape_stack, "gather_nd": _shape_gather_nd, "concat": _shape_concat,
    "broadcast_add": _same, "broadcast_mul": _same,
    "Concat": _shape_concat, "UpSampling": _shape_upsample, "slice_like": _shape_slice_like, "slice_axis": _shape_slice_axis,
    "Reshape": lambda a, ins, n: (ins, [mx_reshape(ins[0], a["shape"])]),
    "reshape": lambda a, ins, n: (ins, [mx_reshape(ins[0], a["shape"])]),
    "Flatten": lambda a, ins, n: (ins, [(ins[0][0], math.prod(ins[0][1:]))]),
    "_contrib_Proposal_v3": _shape_proposal, "_contrib_Proposal": _shape_proposal, "_contrib_Proposal_v2": _shape_proposal,
    "ProposalTarget_v2": _shape_proposal_target,
    "_contrib_ROIAlign_v2": _shape_roialign, "_contrib_DecodeBBox": _shape_decode, "Custom": _shape_custom,
    "Deconvolution": _shape_deconv, "_contrib_DeformableConvolution": _shape_deform_conv,
    "_contrib_GenAnchor": _shape_gen_anchor, "_contrib_GenProposalRetina": _shape_gen_proposal_retina,
    "_full": lambda a, ins, n: (ins, [tuple(_t(a["shape"]))]), "full": lambda a, ins, n: (ins, [tuple(_t(a["shape"]))]),
    "transpose": lambda a, ins, n: (ins, [tuple(ins[0][i] for i in _t(a["axes"]))]),
}


def infer_shapes(sym: S.Symbol, known: dict):
    """mx Symbol.infer_shape: (arg_shapes, out_shapes, aux_shapes) in list_arguments / list_outputs / aux order."""
    shapes: dict[int, list] = {}
    var_shape: dict[str, tuple] = {k: tuple(v) for k, v in known.items()}
    for node in sym._topo():
        if node.op is None:
            s = var_shape.get(node.name) or node.attrs.get("__shape__")
            shapes[id(node)] = [tuple(s) if s is not None else None]
            continue
        ins = [shapes[id(n)][i] for s_ in node.inputs for n, i in s_.entries]
        names = _flat_names(node)
        rule = SHAPE_RULES.get(node.op)
        if rule is None:
            raise NotImplementedError(f"shape rule for operator {node.op!r} ({node.name})")
        ins2, outs = rule(node.attrs, ins, names)
        k = 0
        for s_ in node.inputs:
            for n, i in s_.entries:
                if n.op is None and shapes[id(n)][0] is None and ins2[k] is not None:
                    shapes[id(n)][0] = tuple(ins2[k])
                    var_shape[n.name] = tuple(ins2[k])
                k += 1
        shapes[id(node)] = [tuple(o) for o in outs]
    args = [var_shape.get(n) for n in sym.list_arguments()]
    aux = [var_shape.get(n) for n in sym.list_auxiliary_states()]
    outs = [shapes[id(n)][i] for n, i in sym.entries]
    return args, outs, aux


def _flat_names(node):
    names = []
    k = 0
    for s_ in node.inputs:
        for _ in s_.entries:
            names.append(node.arg_names[k] if k < len(node.arg_names) else None)
        k += 1
    return names


# ------------------------------------------------------------------------------------------------
# execution
# ------------------------------------------------------------------------------------------------
class Executor:
    """Evaluates a façade Symbol with torch on one device.  `channels_last=True` keeps 4-D activations in torch's
    channels_last memory format (what the tensor-core convolutions prefer) and hands the FPN levels to the fused
    RoIAlign as NHWC - no re-layout pass."""

    def __init__(self, sym: S.Symbol, device="cuda:0", channels_last=True, fuse_fpn_roi_align=True, fold_bn=True,
                 is_train=False):
        import torch

        self.sym, self.device = sym, torch.device(device)
        # is_train: MXNet's forward(is_train=True) - loss operators get their MXNet backward (facade/train.py),
        # RoIAlign goes through the autograd operators, BatchNorm without use_global_stats uses batch statistics,
        # frozen BatchNorm is NOT folded into the convolution (the convolution's weights are being trained)
        self.is_train = bool(is_train)
        fold_bn = fold_bn and not self.is_train
        self.channels_last, self.fuse, self.fold_bn = channels_last, fuse_fpn_roi_align, fold_bn
        self.params: dict = {}
        self.order = sym._topo()
        self._fusions = self._find_fpn_roi_align() if fuse_fpn_roi_align else {}
        self._bn_of_conv = self._find_conv_bn() if fold_bn else {}
        self._folded = None  # conv node id -> (weight, bias) with the BatchNorm folded in; rebuilt after init_params

    # ---- pattern: Convolution -> BatchNorm on moving statistics (the frozen BN of the detection backbones) ==> the
    # affine map is folded into the convolution's weights once, the BatchNorm node becomes a pass-through
    def _find_conv_bn(self):
        uses: dict[int, int] = {}
        for node in self.order:
            for s_ in node.inputs:
                for n, _ in s_.entries:
                    uses[id(n)] = uses.get(id(n), 0) + 1
        for n, _ in self.sym.entries:
            uses[id(n)] = uses.get(id(n), 0) + 1
        pairs = {}
        for node in self.order:
            if node.op != "BatchNorm":
                continue
            names = _flat_names(node)
            ins = [e for s_ in node.inputs for e in s_.entries]
            byname = dict(zip(names, ins))
            src = byname.get("data", ins[0])[0]
            stat = [byname.get(k) for k in ("gamma", "beta", "moving_mean", "moving_var")]
            if src.op != "Convolution" or uses.get(id(src), 0) != 1 or any(e is None or e[0].op is not None for e in stat):
                continue
            cn = dict(zip(_flat_names(src), [e for s_ in src.inputs for e in s_.entries]))
            if cn.get("weight") is None or cn["weight"][0].op is not None or ("bias" in cn and cn["bias"][0].op is not None):
                continue
            pairs[id(src)] = node
        return pairs

    def _fold(self, torch):
        folded = {}
        for node in self.order:
            bn = self._bn_of_conv.get(id(node))
            if bn is None:
                continue
            cn = dict(zip(_flat_names(node), [e for s_ in node.inputs for e in s_.entries]))
            bnn = dict(zip(_flat_names(bn), [e for s_ in bn.inputs for e in s_.entries]))
            P = lambda e: self.params[e[0].name]  # noqa: E731
            w = P(cn["weight"]).double()
            b = P(cn["bias"]).double() if "bias" in cn else torch.zeros(w.shape[0], device=w.device, dtype=torch.float64)
            gamma = torch.ones_like(P(bnn["gamma"])) if _b(bn.attrs.get("fix_gamma", True)) else P(bnn["gamma"])
            scale = gamma.double() / torch.sqrt(P(bnn["moving_var"]).double() + float(_t(bn.attrs.get("eps", 1e-3))))
            wf = (w * scale.view(-1, 1, 1, 1)).float()
            bf = ((b - P(bnn["moving_mean"]).double()) * scale + P(bnn["beta"]).double()).float()
            if self.channels_last:
                wf = wf.contiguous(memory_format=torch.channels_last)
            folded[id(node)] = (wf, bf)
        self._folded = folded
        self._folded_bn = {id(bn) for bn in self._bn_of_conv.values()}

    # ---- parameters
    def init_params(self, input_shapes: dict, arg_params=None, aux_params=None, rng_std=None):
        """Zero-initialise every parameter the caller did not give (detection_infer_speed.py's setting); `rng_std`
        draws N(0, rng_std) weights instead (BatchNorm: gamma 1, var 1)."""
        import torch

        args, _, aux = infer_shapes(self.sym, input_shapes)
        given = dict(arg_params or {})
        given.update(aux_params or {})
        g = torch.Generator(device=self.device).manual_seed(0)
        for name, shp in list(zip(self.sym.list_arguments(), args)) + list(zip(self.sym.list_auxiliary_states(), aux)):
            if name in input_shapes:
                continue
            if name in given:
                v = given[name]
                v = v.t if hasattr(v, "t") and not callable(v.t) else torch.as_tensor(v)
                self.params[name] = v.to(self.device, torch.float32)
            elif shp is None:
                raise ValueError(f"cannot infer the shape of parameter {name}")
            elif rng_std is None:
                self.params[name] = torch.zeros(shp, device=self.device)
            elif name.endswith(("gamma", "moving_var")):
                self.params[name] = torch.ones(shp, device=self.device)
            elif name.endswith(("beta", "moving_mean", "bias")):
                self.params[name] = torch.zeros(shp, device=self.device)
            else:
                self.params[name] = torch.randn(shp, device=self.device, generator=g) * rng_std
        self._folded = None
        return self

    # ---- pattern: assign_layer_fpn -> L x ROIAlign_v2 -> add_n  ==> one fused FPN RoIAlign
    def _find_fpn_roi_align(self):
        fus = {}
        for node in self.order:
            if node.op != "add_n":
                continue
            ins = [e for s_ in node.inputs for e in s_.entries]
            # models/FPN/builder.py:588-605 reshapes every level's (B,N,C,ph,pw) result with (-3,-2) before the add_n
            reshape = None
            if ins and all(n.op == "Reshape" and len(n.inputs) == 1 for n, _ in ins):
                shp = {str(_t(n.attrs["shape"])) for n, _ in ins}
                if len(shp) == 1:
                    reshape = _t(ins[0][0].attrs["shape"])
                    self._skipped_reshapes = getattr(self, "_skipped_reshapes", set()) | {id(n) for n, _ in ins}
                    ins = [n.inputs[0].entries[0] for n, _ in ins]
            if not ins or any(n.op != "_contrib_ROIAlign_v2" or i != 0 for n, i in ins):
                continue
            ras = [n for n, _ in ins]
            roi_srcs = [ra.inputs[1].entries[0] for ra in ras]
            asg = roi_srcs[0][0]
            if asg.op != "Custom" or asg.attrs.get("op_type") != "assign_layer_fpn":
                continue
            if any(n is not asg for n, _ in roi_srcs) or sorted(i for _, i in roi_srcs) != list(range(len(ras))):
                continue
            fus[id(node)] = (asg, ras, reshape)
        return fus

    def forward(self, **inputs):
        import torch

        from .. import ops

        vals: dict[int, list] = {}
        if self.fold_bn and self._folded is None:
            self._fold(torch)
        for node in self.order:
            if node.op is None:
                if node.name in inputs:
                    v = inputs[node.name]
                    v = v.t if hasattr(v, "t") and not callable(getattr(v, "t")) else torch.as_tensor(v)
                    vals[id(node)] = [v.to(self.device, torch.float32)]
                elif node.name in self.params:
                    vals[id(node)] = [self.params[node.name]]
                else:
                    raise KeyError(f"no value for variable {node.name}")
                continue
            if id(node) in self._fusions:
                asg, ras, reshape = self._fusions[id(node)]
                rois = vals[id(asg.inputs[0].entries[0][0])][asg.inputs[0].entries[0][1]]
                strides = [int(s) for s in _t(asg.attrs["rcnn_stride"])]
                feats = []
                for k in range(len(ras)):
                    ra = next(r for r in ras if r.inputs[1].entries[0][1] == k)
                    fn, fi = ra.inputs[0].entries[0]
                    feats.append(vals[id(fn)][fi])
                ph, pw = _tup(ras[0].attrs["pooled_size"])
                scale0, lvl0 = int(_t(asg.attrs["roi_canonical_scale"])), int(_t(asg.attrs["roi_canonical_level"]))
                if self.is_train:   # the autograd operator: gradients flow into every FPN level
                    out = ops.fpn_roi_align([f.contiguous() for f in feats], rois.contiguous(), tuple(strides), (ph, pw),
                                            scale0, lvl0)
                elif self.channels_last and all(f.is_contiguous(memory_format=torch.channels_last) for f in feats):
                    out, _ = ops.fpn_roi_align_nhwc([f.permute(0, 2, 3, 1) for f in feats], rois.contiguous(), strides,
                                                    (ph, pw), scale0, lvl0)
                else:
                    out = ops.fpn_roi_align_raw([f.contiguous() for f in feats], rois.contiguous(), strides, (ph, pw),
                                                scale0, lvl0, with_argmax=False)[0]
                if reshape is not None:
                    out = out.reshape(mx_reshape(tuple(out.shape), reshape))
                vals[id(node)] = [out]
                continue
            if node.op in ("_contrib_ROIAlign_v2",) and self._is_fused_member(node):
                continue
            if node.op == "Reshape" and id(node) in getattr(self, "_skipped_reshapes", ()) and self._feeds_fused(node):
                continue
            if node.op == "Custom" and node.attrs.get("op_type") == "assign_layer_fpn" and self._only_feeds_fusion(node):
                continue
            ins = [vals[id(n)][i] for s_ in node.inputs for n, i in s_.entries]
            vals[id(node)] = self._run(node, ins, ops, torch)
        return [vals[id(n)][i] for n, i in self.sym.entries]

    def capture(self, **inputs):
        """Record one forward pass into a CUDA graph and return `run(**inputs) -> outputs` that replays it (inputs are
        copied into the graph's static buffers; the outputs are the graph's static output tensors).  The ~700 small
        launches of a detection graph are host-bound when issued one by one; every detection operator behind `OPS`
        is capture-safe (no allocation, no host synchronisation: tests/test_cuda_graph_gpu.py)."""
        import torch

        static = {k: (v.t if hasattr(v, "t") and not callable(getattr(v, "t")) else torch.as_tensor(v)).to(
            self.device, torch.float32).clone() for k, v in inputs.items()}
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(3):  # cuDNN autotuning, lazy workspaces, BatchNorm folding
                self.forward(**static)
        torch.cuda.current_stream(self.device).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph), torch.no_grad():
            outs = self.forward(**static)

        def run(**new_inputs):
            for k, v in new_inputs.items():
                static[k].copy_(v.t if hasattr(v, "t") and not callable(getattr(v, "t")) else torch.as_tensor(v))
            graph.replay()
            return outs

        run.graph = graph
        return run

    def _is_fused_member(self, node):
        return any(node in ras for _, ras, _r in self._fusions.values())

    def _feeds_fused(self, node):
        src = node.inputs[0].entries[0][0]
        return self._is_fused_member(src)

    def _only_feeds_fusion(self, node):
        return any(asg is node for asg, _, _r in self._fusions.values())

    def _run(self, node, x, ops, torch):
        F = torch.nn.functional
        a, op = node.attrs, node.op
        names = _flat_names(node)
        arg = {n: v for n, v in zip(names, x) if n}
        if op == "Convolution":
            data = arg.get("data", x[0])
            if self.channels_last and data.dim() == 4:
                data = data.contiguous(memory_format=torch.channels_last)
            w, bias = arg["weight"], arg.get("bias")
            if self._folded and id(node) in self._folded:
                w, bias = self._folded[id(node)]
            return [F.conv2d(data, w, bias, _tup(a.get("stride", 1)), _tup(a.get("pad", 0)),
                             _tup(a.get("dilate", 1)), int(_t(a.get("num_group", 1))))]
        if op == "FullyConnected":
            data = arg.get("data", x[0])
            if _b(a.get("flatten", True)):
                data = data.reshape(data.shape[0], -1)
            return [F.linear(data, arg["weight"], arg.get("bias"))]
        if op == "BatchNorm":  # inference form (use_global_stats / is_train=False): moving statistics
            data = arg.get("data", x[0])
            if self._folded and id(node) in self._folded_bn:
                return [data]  # already applied by the convolution that feeds it
            g = torch.ones_like(arg["gamma"]) if _b(a.get("fix_gamma", True)) else arg["gamma"]
            if self.is_train and not _b(a.get("use_global_stats", False)):
                # batch statistics; MXNet's momentum weighs the OLD moving value, torch's the new batch value
                return [F.batch_norm(data, arg["moving_mean"], arg["moving_var"], g, arg["beta"], True,
                                     1.0 - float(_t(a.get("momentum", 0.9))), float(_t(a.get("eps", 1e-3))))]
            return [F.batch_norm(data, arg["moving_mean"], arg["moving_var"], g, arg["beta"], False, 0.0,
                                 float(_t(a.get("eps", 1e-3))))]
        if op in ("Activation", "relu"):
            t = str(a.get("act_type", "relu"))
            return [{"relu": F.relu, "sigmoid": torch.sigmoid, "tanh": torch.tanh}[t](x[0])]
        if op == "Pooling":
            if _b(a.get("global_pool", False)):
                return [x[0].mean((2, 3), keepdim=True) if str(a.get("pool_type", "max")) == "avg" else x[0].amax((2, 3), keepdim=True)]
            k, s, p = _tup(a["kernel"]), _tup(a.get("stride", 1)), _tup(a.get("pad", 0))
            ceil = str(a.get("pooling_convention", "valid")) == "full"
            if str(a.get("pool_type", "max")) == "max":
                return [F.max_pool2d(x[0], k, s, p, ceil_mode=ceil)]
            return [F.avg_pool2d(x[0], k, s, p, ceil_mode=ceil)]
        if self.is_train and op in ("SoftmaxOutput", "MakeLoss", "BlockGrad", "smooth_l1"):
            from . import train as T

            if op == "SoftmaxOutput":
                return [T.softmax_output(arg.get("data", x[0]), arg.get("label", x[1]), _b(a.get("multi_output", False)),
                                         str(a.get("normalization", "null")), _b(a.get("use_ignore", False)),
                                         float(_t(a.get("ignore_label", -1))), float(_t(a.get("grad_scale", 1.0))))]
            if op == "MakeLoss":
                return [T.make_loss(x[0], float(_t(a.get("grad_scale", 1.0))), str(a.get("normalization", "null")),
                                    float(_t(a.get("valid_thresh", 0.0))))]
            if op == "BlockGrad":
                return [x[0].detach()]
            return [T.smooth_l1(x[0], float(_t(a.get("scalar", 1.0))))]
        if op == "smooth_l1":
            from . import train as T

            return [T.smooth_l1(x[0], float(_t(a.get("scalar", 1.0))))]
        if op in ("Cast", "BlockGrad", "identity", "Dropout", "MakeLoss"):
            return [x[0]]
        if op == "UpSampling":
            sc = int(_t(a["scale"]))
            return [F.interpolate(x[0], scale_factor=sc, mode="nearest")]
        if op == "slice_like":
            axes = _t(a.get("axes", ())) or range(x[0].dim())
            sl = [slice(None)] * x[0].dim()
            for ax in axes:
                sl[ax] = slice(0, x[1].shape[ax])
            return [x[0][tuple(sl)]]
        if op == "slice_axis":
            ax, b = int(_t(a["axis"])), int(_t(a["begin"]))
            e = _t(a["end"])
            sl = [slice(None)] * x[0].dim()
            sl[ax] = slice(b, None if e in (None, "None") else int(e))
            return [x[0][tuple(sl)]]
        if op in ("add_n", "elemwise_add", "broadcast_add"):
            out = x[0]
            for v in x[1:]:
                out = out + v
            return [out]
        if op in ("elemwise_sub",):
            return [x[0] - x[1]]
        if op in ("elemwise_mul", "broadcast_mul"):
            return [x[0] * x[1]]
        if op == "elemwise_div":
            return [x[0] / x[1]]
        if op == "_power_scalar":
            return [x[0] ** float(a["scalar"])]
        if op in ("maximum", "minimum"):
            if len(x) == 2:
                return [torch.maximum(x[0], x[1]) if op == "maximum" else torch.minimum(x[0], x[1])]
            return [x[0].clamp(min=float(_t(a["scalar"]))) if op == "maximum" else x[0].clamp(max=float(_t(a["scalar"])))]
        if op == "sum":
            ax = a.get("axis")
            if ax in (None, "None", ()):
                return [x[0].sum().reshape(1)]
            axes = _t(ax) if isinstance(_t(ax), (tuple, list)) else (_t(ax),)
            return [x[0].sum(tuple(int(v) for v in axes), keepdim=_b(a.get("keepdims", False)))]
        if op == "expand_dims":
            return [x[0].unsqueeze(int(_t(a["axis"])))]
        if op.endswith("_scalar"):
            s, rev = float(a["scalar"]), bool(a.get("__rev__", False))
            return [{"_plus_scalar": lambda v: v + s, "_mul_scalar": lambda v: v * s,
                     "_minus_scalar": lambda v: (s - v) if rev else (v - s),
                     "_div_scalar": lambda v: (s / v) if rev else (v / s)}[op](x[0])]
        if op in ("Concat", "concat"):
            return [torch.cat(x, int(_t(a.get("dim", 1))))]
        if op in ("split", "SliceChannel"):
            n, ax = int(_t(a["num_outputs"])), int(_t(a.get("axis", 1)))
            parts = torch.chunk(x[0], n, ax)
            return [p.squeeze(ax) for p in parts] if _b(a.get("squeeze_axis", False)) else list(parts)
        if op in ("arange", "_arange"):
            start, stop, step, rep = _arange_len(a)
            return [torch.arange(start, stop, step, device=self.device, dtype=torch.float32).repeat_interleave(rep)]
        if op == "stack":
            return [torch.stack(x, int(_t(a.get("axis", 0))))]
        if op == "gather_nd":   # data[indices[0], indices[1], ...]: the mask channel of each roi's class
            idx = x[1].to(torch.long)
            return [x[0][tuple(idx[i] for i in range(idx.shape[0]))]]
        if op in ("Reshape", "reshape"):
            return [x[0].reshape(mx_reshape(tuple(x[0].shape), a["shape"]))]
        if op == "Flatten":
            return [x[0].reshape(x[0].shape[0], -1)]
        if op == "softmax":
            return [torch.softmax(x[0], int(_t(a.get("axis", -1))))]
        if op == "SoftmaxActivation":
            return [torch.softmax(x[0], 1 if str(a.get("mode", "instance")) == "channel" else -1)]
        if op == "SoftmaxOutput":
            return [torch.softmax(x[0], 1 if _b(a.get("multi_output", False)) else -1)]
        if op == "Deconvolution":
            data = arg.get("data", x[0])
            return [F.conv_transpose2d(data, arg["weight"], arg.get("bias"), _tup(a.get("stride", 1)), _tup(a.get("pad", 0)),
                                       _tup(a.get("adj", 0)), int(_t(a.get("num_group", 1))))]
        if op in ("full", "_full"):
            return [torch.full(tuple(_t(a["shape"])), float(_t(a["value"])), device=self.device, dtype=torch.float32)]
        if op == "transpose":
            return [x[0].permute(*_t(a["axes"]))]
        # ---- detection operators: by registration string, through the C ABI
        if op == "_contrib_DeformableConvolution":
            data, offset = arg.get("data", x[0]), arg.get("offset", x[1])
            return [ops.OPS[op](data, offset.contiguous(), arg["weight"], arg.get("bias"), kernel=_tup(a["kernel"]),
                                stride=_tup(a.get("stride", 1)), dilate=_tup(a.get("dilate", 1)), pad=_tup(a.get("pad", 0)),
                                num_filter=int(_t(a["num_filter"])), num_group=int(_t(a.get("num_group", 1))),
                                num_deformable_group=int(_t(a.get("num_deformable_group", 1))),
                                no_bias=_b(a.get("no_bias", False)))]
        if op == "_contrib_GenAnchor":
            return [ops.OPS[op](x[0], feature_stride=int(_t(a["feature_stride"])),
                                scales=tuple(float(v) for v in _t(a["scales"])),
                                ratios=tuple(float(v) for v in _t(a["ratios"])))]
        if op == "_contrib_GenProposalRetina":
            return list(ops.OPS[op](arg["cls_prob"].contiguous(), arg["bbox_pred"].contiguous(), arg["im_info"].contiguous(),
                                    arg["anchors"].contiguous(), feature_stride=int(_t(a.get("feature_stride", 16))),
                                    rpn_pre_nms_top_n=int(_t(a.get("rpn_pre_nms_top_n", 6000))),
                                    rpn_min_size=int(_t(a.get("rpn_min_size", 16))), num_anchors=int(_t(a["num_anchors"])),
                                    thresh=float(_t(a.get("thresh", 0.0))),
                                    anchor_mean=tuple(float(v) for v in _t(a["anchor_mean"])),
                                    anchor_std=tuple(float(v) for v in _t(a["anchor_std"])),
                                    output_one_hot=_b(a.get("output_one_hot", True))))
        if op == "_contrib_Proposal":
            r = ops.OPS[op](arg["cls_prob"].contiguous(), arg["bbox_pred"].contiguous(), arg["im_info"].contiguous(),
                            rpn_pre_nms_top_n=int(_t(a.get("rpn_pre_nms_top_n", 6000))),
                            rpn_post_nms_top_n=int(_t(a.get("rpn_post_nms_top_n", 300))),
                            threshold=float(_t(a.get("threshold", 0.7))), rpn_min_size=int(_t(a.get("rpn_min_size", 16))),
                            scales=tuple(float(s) for s in _t(a.get("scales", (4, 8, 16, 32)))),
                            ratios=tuple(float(s) for s in _t(a.get("ratios", (0.5, 1, 2)))),
                            feature_stride=int(_t(a.get("feature_stride", 16))), output_score=True,
                            iou_loss=_b(a.get("iou_loss", False)))
            return list(r)
        if op == "_contrib_Proposal_v3":
            r = ops.OPS[op](arg["cls_prob"].contiguous(), arg["bbox_pred"].contiguous(), arg["im_info"].contiguous(),
                            rpn_pre_nms_top_n=int(_t(a.get("rpn_pre_nms_top_n", 6000))),
                            rpn_post_nms_top_n=int(_t(a.get("rpn_post_nms_top_n", 300))),
                            threshold=float(_t(a.get("threshold", 0.7))), rpn_min_size=int(_t(a.get("rpn_min_size", 16))),
                            scales=tuple(float(s) for s in _t(a.get("scales", (4, 8, 16, 32)))),
                            ratios=tuple(float(s) for s in _t(a.get("ratios", (0.5, 1, 2)))),
                            feature_stride=int(_t(a.get("feature_stride", 16))), output_score=True,
                            iou_loss=_b(a.get("iou_loss", False)))
            return list(r)
        if op == "_contrib_ROIAlign_v2" and self.is_train:
            out = ops.OPS[op](arg["data"].contiguous(), arg["rois"].contiguous(), _tup(a["pooled_size"]),
                              float(_t(a["spatial_scale"])))
            return [out, out, out]   # argmax_x / argmax_y are not visible outputs of the symbol
        if op == "_contrib_Proposal_v2":   # TridentNet: Proposal + valid_ranges (models/tridentnet/builder.py:239)
            return list(ops.OPS[op](arg["cls_prob"].contiguous(), arg["bbox_pred"].contiguous(), arg["im_info"].contiguous(),
                                    arg["valid_ranges"].contiguous(),
                                    rpn_pre_nms_top_n=int(_t(a.get("rpn_pre_nms_top_n", 6000))),
                                    rpn_post_nms_top_n=int(_t(a.get("rpn_post_nms_top_n", 300))),
                                    threshold=float(_t(a.get("threshold", 0.7))), rpn_min_size=int(_t(a.get("rpn_min_size", 16))),
                                    scales=tuple(float(v) for v in _t(a.get("scales", (4, 8, 16, 32)))),
                                    ratios=tuple(float(v) for v in _t(a.get("ratios", (0.5, 1, 2)))),
                                    feature_stride=int(_t(a.get("feature_stride", 16))), output_score=True,
                                    iou_loss=_b(a.get("iou_loss", False)), filter_scales=_b(a.get("filter_scales", False))))
        if op in ("ProposalTarget", "ProposalTarget_v2"):
            kw = dict(num_classes=int(_t(a["num_classes"])), batch_images=int(_t(a["batch_images"])),
                      image_rois=int(_t(a["image_rois"])), fg_thresh=float(_t(a["fg_thresh"])),
                      bg_thresh_hi=float(_t(a["bg_thresh_hi"])), bg_thresh_lo=float(_t(a["bg_thresh_lo"])),
                      proposal_without_gt=_b(a.get("proposal_without_gt", False)),
                      fg_fraction=float(_t(a.get("fg_fraction", 0.25))), class_agnostic=_b(a.get("class_agnostic", False)),
                      output_iou=_b(a.get("output_iou", False)),
                      bbox_mean=tuple(float(v) for v in _t(a.get("bbox_mean", (0, 0, 0, 0)))),
                      bbox_std=tuple(float(v) for v in _t(a.get("bbox_std", (0.1, 0.1, 0.2, 0.2)))),
                      bbox_weight=tuple(float(v) for v in _t(a.get("bbox_weight", (1, 1, 1, 1)))))
            with torch.no_grad():   # ProposalTargetProp: no gradient to either input
                if op == "ProposalTarget_v2":
                    return list(ops.OPS[op](arg.get("rois", x[0]).detach().contiguous(),
                                            arg.get("gt_boxes", x[1]).detach().contiguous(),
                                            arg.get("valid_ranges", x[2]).detach().contiguous(),
                                            filter_scales=_b(a.get("filter_scales", False)), **kw))
                return list(ops.OPS[op](arg.get("rois", x[0]).detach().contiguous(),
                                        arg.get("gt_boxes", x[1]).detach().contiguous(), **kw))
        if op == "ProposalMaskTarget":
            kw = dict(num_classes=int(_t(a["num_classes"])), batch_images=int(_t(a["batch_images"])),
                      image_rois=int(_t(a["image_rois"])), mask_size=int(_t(a["mask_size"])),
                      fg_thresh=float(_t(a["fg_thresh"])), bg_thresh_hi=float(_t(a["bg_thresh_hi"])),
                      bg_thresh_lo=float(_t(a["bg_thresh_lo"])), proposal_without_gt=_b(a.get("proposal_without_gt", False)),
                      fg_fraction=float(_t(a.get("fg_fraction", 0.25))), class_agnostic=_b(a.get("class_agnostic", False)),
                      output_iou=_b(a.get("output_iou", False)), output_ratio=_b(a.get("output_ratio", False)),
                      bbox_mean=tuple(float(v) for v in _t(a.get("bbox_mean", (0, 0, 0, 0)))),
                      bbox_std=tuple(float(v) for v in _t(a.get("bbox_std", (0.1, 0.1, 0.2, 0.2)))),
                      bbox_weight=tuple(float(v) for v in _t(a.get("bbox_weight", (1, 1, 1, 1)))))
            with torch.no_grad():
                return list(ops.OPS[op](arg.get("rois", x[0]).detach().contiguous(),
                                        arg.get("gt_boxes", x[1]).detach().contiguous(),
                                        arg.get("gt_polys", x[2]).detach().contiguous(), **kw))
        if op == "_contrib_SigmoidCrossEntropy":
            return [ops.OPS[op](x[0].contiguous(), x[1].contiguous(), grad_scale=float(_t(a.get("grad_scale", 1.0))))]
        if op == "_contrib_FocalLoss":
            return [ops.OPS[op](arg.get("data", x[0]).contiguous(), arg.get("label", x[1]).contiguous(),
                                alpha=float(_t(a.get("alpha", 0.25))), gamma=float(_t(a.get("gamma", 2.0))),
                                normalization=str(a.get("normalization", "null")),
                                grad_scale=float(_t(a.get("grad_scale", 1.0))))]
        if op == "_contrib_BBoxNorm":
            return [ops.OPS[op](arg.get("data", x[0]).contiguous(), arg.get("label", x[1]).contiguous(),
                                normalization=str(a.get("normalization", "valid")))]
        if op == "_contrib_ROIAlign_v2":
            out, ax, ay = ops.roi_align_v2_raw(arg["data"].contiguous(), arg["rois"].contiguous(), _tup(a["pooled_size"]),
                                               float(_t(a["spatial_scale"])), with_argmax=False)
            return [out, ax, ay]
        if op == "_contrib_DecodeBBox":
            return [ops.OPS[op](arg["rois"].contiguous(), arg["bbox_pred"].contiguous(), arg["im_info"].contiguous(),
                                tuple(_t(a.get("bbox_mean", (0, 0, 0, 0)))), tuple(_t(a.get("bbox_std", (0.1, 0.1, 0.2, 0.2)))),
                                class_agnostic=_b(a.get("class_agnostic", True)))]
        if op == "Custom":
            t = a.get("op_type")
            if t == "get_top_proposal":
                return list(ops.OPS[t](arg["bbox"].contiguous(), arg["score"].contiguous(), int(_t(a["top_n"]))))
            if t == "assign_layer_fpn":
                return list(ops.OPS[t](x[0].contiguous(), tuple(_t(a["rcnn_stride"])), int(_t(a["roi_canonical_scale"])),
                                       int(_t(a["roi_canonical_level"]))))
            if t == "maskiou_compute":
                with torch.no_grad():
                    return list(ops.OPS[t](*[v.detach().contiguous() for v in x[:4]]))
            if t == "bbox_target":
                return list(ops.OPS[t](x[0].contiguous(), x[1].contiguous(), int(_t(a["num_class"])),
                                       _b(a["add_gt_to_proposal"]), int(_t(a["image_rois"])), float(_t(a["fg_fraction"])),
                                       float(_t(a["fg_thresh"])), float(_t(a["bg_thresh_hi"])), float(_t(a["bg_thresh_lo"])),
                                       tuple(float(v) for v in _t(a["bbox_target_std"]))))
            if t == "BboxPostProcessing":
                return list(ops.OPS[t](x[0].contiguous(), x[1].contiguous(), int(_t(a["max_det_per_image"])),
                                       float(_t(a["min_det_score"])), str(a.get("nms_type", "nms")), float(_t(a["nms_thr"]))))
        raise NotImplementedError(f"operator {op!r} ({node.name}) is not wired into the façade executor")
