"""A stand-in for the third-party `mxnext` package (RogerChern/mxnext, un-pinned and not vendored in the reference:
SURVEY §8c(3)) - written from the reference's CALL SITES (symbol/builder.py, models/FPN/builder.py, config/*): the
`X.*` graph helpers, `normalizer_factory`, the ResNet-v1 FPN backbone builder and the `mxnext.tvm.*` operator
helpers.  Graph helpers only create façade symbols with the MXNet operator names the real package emits
(`X.conv` -> Convolution, `X.roi_align` -> _contrib_ROIAlign_v2, `X.proposal_target` -> ProposalTarget, ...:
SURVEY §8b's name <-> kwargs table)."""
from __future__ import annotations

from . import symbol as S

sym = S._OpNamespace("")
contrib = S._OpNamespace("_contrib_")


# ---- initialisers / variables ---------------------------------------------------------------------------------
class _Init:
    def __init__(self, kind, **kw):
        self.kind, self.kw = kind, kw

    def dumps(self):
        import json

        return json.dumps([self.kind, self.kw])


def gauss(std):
    return _Init("normal", sigma=std)


def zero_init():
    return _Init("zeros")


def one_init():
    return _Init("ones")


def constant(value):
    return _Init("constant", value=value)


def var(name, init=None, lr_mult=None, wd_mult=None, shape=None, dtype=None, **kw):
    return S.Variable(name, shape=shape, lr_mult=lr_mult, wd_mult=wd_mult, dtype=dtype, init=init, **kw)


_SHARED_VARS = {}


def shared_var(name, **kw):
    """One variable node per name, however often it is asked for (TridentNet's branches share weights this way:
    models/tridentnet/builder_v2.py:25-35)."""
    if name not in _SHARED_VARS:
        _SHARED_VARS[name] = var(name, **kw)
    return _SHARED_VARS[name]


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (int(v), int(v))


# ---- layers ----------------------------------------------------------------------------------------------------
def conv(data, name, filter, kernel=1, stride=1, pad=-1, dilate=1, num_group=1, no_bias=True, init=None, lr_mult=1.0,
         wd_mult=1.0, weight=None, bias=None):
    k, s, d = _pair(kernel), _pair(stride), _pair(dilate)
    p = _pair(pad) if pad != -1 else tuple(((kk - 1) * dd + 1) // 2 for kk, dd in zip(k, d))   # "same" for odd kernels
    weight = weight if weight is not None else var(name + "_weight", init=init, lr_mult=lr_mult, wd_mult=wd_mult)
    kw = dict(data=data, weight=weight, kernel=k, stride=s, pad=p, dilate=d, num_filter=int(filter), num_group=num_group,
              no_bias=bool(no_bias), name=name)
    if not no_bias:
        kw["bias"] = bias if bias is not None else var(name + "_bias", init=zero_init(), lr_mult=lr_mult, wd_mult=wd_mult)
    return sym.Convolution(**kw)


def dwconv(data, name, filter, kernel=3, stride=1, pad=-1, dilate=1, no_bias=True, init=None, **kw):
    """Depthwise convolution (models/efficientnet/builder.py:47): one group per channel."""
    return conv(data, name, filter, kernel, stride, pad, dilate, num_group=int(filter), no_bias=no_bias, init=init)


def gn(data, name=None, num_group=32, eps=1e-5, **kw):
    return contrib.GroupNorm(data=data, name=name, eps=eps, num_group=num_group)


def merge_sum(symbols, name=None):
    return add_n(*symbols, name=name)


def reluconvbn(data, filter=None, init=None, norm=None, name=None, prefix="", kernel=3, stride=1, pad=-1, filters=None,
               **kw):
    """NAS-FPN's / FPG's cell op (models/NASFPN/builder.py:63 positional, models/FPG/builder.py:94 by keyword with
    `filters`): relu -> conv -> norm, parameters under `prefix`."""
    r = relu(data, name=prefix + name + "_relu")
    c = conv(r, prefix + name + "_conv", filter if filter is not None else filters, kernel=kernel, stride=stride, pad=pad,
             no_bias=False, init=init)
    return norm(c, name=prefix + name + "_bn")


def relu6(data, name=None):
    return sym.clip(data=data, a_min=0.0, a_max=6.0, name=name)


def fc(data, name, filter, no_bias=False, flatten=True, init=None, weight=None, bias=None, lr_mult=1.0, wd_mult=1.0):
    weight = weight if weight is not None else var(name + "_weight", init=init, lr_mult=lr_mult, wd_mult=wd_mult)
    kw = dict(data=data, weight=weight, num_hidden=int(filter), no_bias=bool(no_bias), flatten=flatten, name=name)
    if not no_bias:
        kw["bias"] = bias if bias is not None else var(name + "_bias", init=zero_init(), lr_mult=lr_mult, wd_mult=wd_mult)
    return sym.FullyConnected(**kw)


def relu(data, name=None):
    return sym.Activation(data=data, act_type="relu", name=name)


def pool(data, name=None, kernel=3, stride=2, pad=-1, pool_type="max", pooling_convention="valid", global_pool=False):
    k, s = _pair(kernel), _pair(stride)
    p = _pair(pad) if pad != -1 else tuple(kk // 2 for kk in k)
    return sym.Pooling(data=data, kernel=k, stride=s, pad=p, pool_type=pool_type, pooling_convention=pooling_convention,
                       global_pool=global_pool, name=name)


def max_pool(data, name=None, kernel=3, stride=2, pad=-1, **kw):
    return pool(data, name, kernel, stride, pad, "max", **kw)


def avg_pool(data, name=None, kernel=3, stride=2, pad=-1, **kw):
    return pool(data, name, kernel, stride, pad, "avg", **kw)


def global_avg_pool(data, name=None):
    return sym.Pooling(data=data, kernel=(1, 1), pool_type="avg", global_pool=True, name=name)


_BN_PARAMS = ("gamma", "beta", "moving_mean", "moving_var")


def fixbn(data, name, eps=1e-5, **kw):
    shared = {k: kw[k] for k in _BN_PARAMS if kw.get(k) is not None}   # TridentNet shares BN parameters across branches
    return sym.BatchNorm(data=data, name=name, use_global_stats=True, fix_gamma=False, eps=eps, **shared)


def bn(data, name, eps=1e-5, mom=0.9, **kw):
    shared = {k: kw[k] for k in _BN_PARAMS if kw.get(k) is not None}
    return sym.BatchNorm(data=data, name=name, use_global_stats=False, fix_gamma=False, eps=eps, momentum=mom, **shared)


def convrelu(data, name, filter, kernel=1, stride=1, pad=-1, dilate=1, no_bias=False, init=None, **kw):
    return relu(conv(data, name, filter, kernel, stride, pad, dilate, no_bias=no_bias, init=init), name + "_relu")


def convnorm(normalizer, data, name, filter, kernel=1, stride=1, pad=-1, dilate=1, no_bias=True, init=None, **kw):
    return normalizer(conv(data, name, filter, kernel, stride, pad, dilate, no_bias=no_bias, init=init), name=name + "_bn")


def convnormrelu(normalizer, data, name, filter, kernel=1, stride=1, pad=-1, dilate=1, no_bias=True, init=None, **kw):
    return relu(convnorm(normalizer, data, name, filter, kernel, stride, pad, dilate, no_bias, init), name + "_relu")


def to_fp16(data, name=None):
    return sym.Cast(data=data, dtype="float16", name=name)


def to_fp32(data, name=None):
    return sym.Cast(data=data, dtype="float32", name=name)


def reshape(data, shape, name=None):
    return sym.Reshape(data=data, shape=tuple(shape), name=name)


def transpose(data, axes, name=None):
    return sym.transpose(data=data, axes=tuple(axes), name=name)


def flatten(data, name=None):
    return sym.Flatten(data=data, name=name)


def concat(data, axis=1, name=None):
    return sym.Concat(*data, dim=axis, num_args=len(data), name=name)


def add(lhs, rhs, name=None):
    return sym.elemwise_add(lhs, rhs, name=name)


def add_n(*args, name=None):
    return sym.add_n(*args, num_args=len(args), name=name)


def group(symbols):
    return S.Group(list(symbols))


def softmax(data, axis=-1, name=None):
    return sym.softmax(data=data, axis=axis, name=name)


def sigmoid(data, name=None):
    return sym.Activation(data=data, act_type="sigmoid", name=name)


def softmax_output(data, label, name=None, **kw):
    return sym.SoftmaxOutput(data=data, label=label, name=name, **kw)


def smooth_l1(data, scalar=1.0, name=None):
    return sym.smooth_l1(data=data, scalar=scalar, name=name)


def loss(data, grad_scale=1.0, name=None):
    return sym.MakeLoss(data=data, grad_scale=grad_scale, name=name)


make_loss = loss


def stop_grad(data, name=None):
    return sym.BlockGrad(data=data, name=name)


block_grad = stop_grad


def dropout(data, p=0.5, name=None):
    return sym.Dropout(data=data, p=p, name=name)


# ---- detection operators (names of SURVEY §8b) --------------------------------------------------------------------
def roi_align(feat, rois, out_size, stride, name=None):
    s = _pair(out_size)
    return contrib.ROIAlign_v2(data=feat, rois=rois, pooled_size=s, spatial_scale=1.0 / stride, name=name)


def proposal(cls_prob, bbox_pred, im_info, name=None, **kw):
    return contrib.Proposal(cls_prob=cls_prob, bbox_pred=bbox_pred, im_info=im_info, name=name, **kw)


def proposal_target(rois, gt_boxes, name=None, **kw):
    return sym.ProposalTarget(rois=rois, gt_boxes=gt_boxes, name=name, **kw)


def decode_bbox(rois, bbox_pred, im_info, name=None, **kw):
    return contrib.DecodeBBox(rois=rois, bbox_pred=bbox_pred, im_info=im_info, name=name, **kw)


def focal_loss(data, label, name=None, **kw):
    return contrib.FocalLoss(data=data, label=label, name=name, **kw)


def bbox_norm(data, label, name=None, **kw):
    return contrib.BBoxNorm(data=data, label=label, name=name, **kw)


# ---- mxnext.complicate -----------------------------------------------------------------------------------------------
def normalizer_factory(type="local", ndev=None, eps=1e-5, mom=0.9, wd_mult=1.0, lr_mult=1.0):
    if callable(type):   # models/tridentnet/resnet_v1.py:126 hands an already-built normalizer back to the factory
        return type

    def fix_bn(data, name=None, **kw):
        return fixbn(data, name, eps=eps, **kw)

    def local_bn(data, name=None, **kw):
        return bn(data, name, eps=eps, mom=mom, **kw)

    def dummy(data, name=None, **kw):
        return data

    def sync_bn(data, name=None, **kw):
        shared = {k: kw[k] for k in _BN_PARAMS if kw.get(k) is not None}
        return contrib.SyncBatchNorm(data=data, name=name, fix_gamma=False, eps=eps, momentum=mom, ndev=ndev, key=name,
                                     **shared)

    def gn(data, name=None, **kw):
        return contrib.GroupNorm(data=data, name=name, eps=eps, num_group=32)

    table = {"fixbn": fix_bn, "fix_bn": fix_bn, "fix": fix_bn, "local": local_bn, "localbn": local_bn,
             "local_bn": local_bn, "syncbn": sync_bn, "sync_bn": sync_bn, "gn": gn, "dummy": dummy}
    if type not in table:
        raise NotImplementedError(f"normalizer {type!r}: only fixbn / local / syncbn / gn / dummy are in the stand-in")
    return table[type]


# ---- mxnext.backbone.resnet_v1 ------------------------------------------------------------------------------------------
class ResNetV1Builder:
    """ResNet-v1 (MSRA layout: stride on the first 1x1 conv of a downsampling unit) with the parameter names of the
    reference's pretrained checkpoints (conv0, bn0, stage{s}_unit{u}_conv{1,2,3}, _bn{1,2,3}, _sc, _sc_bn)."""
    depth_config = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}

    @staticmethod
    def unit(data, name, filter, stride, dilate, proj, norm):
        c1 = convnormrelu(norm, data, name + "_conv1", filter // 4, kernel=1, stride=stride)
        c2 = convnormrelu(norm, c1, name + "_conv2", filter // 4, kernel=3, dilate=dilate)
        c3 = convnorm(norm, c2, name + "_conv3", filter, kernel=1)
        sc = convnorm(norm, data, name + "_sc", filter, kernel=1, stride=stride) if proj else data
        return relu(add(c3, sc, name=name + "_plus"), name=name + "_relu")

    @classmethod
    def stage(cls, data, name, num_unit, filter, stride, dilate, norm):
        x = cls.unit(data, f"{name}_unit1", filter, stride, dilate, True, norm)
        for i in range(2, num_unit + 1):
            x = cls.unit(x, f"{name}_unit{i}", filter, 1, dilate, False, norm)
        return x

    @classmethod
    def resnet_stage(cls, data, name, num_block, filter, stride, dilate, norm_type, norm_mom=0.9, ndev=None, **kw):
        """The C5 head of the C4 detectors (symbol/builder.py:624-634): one stage on top of the roi features."""
        return cls.stage(data, name, num_block, filter, stride, dilate, normalizer_factory(norm_type, ndev=ndev, mom=norm_mom))

    # the piecewise interface TridentNet's builder subclasses (models/tridentnet/resnet_v1.py:195-240)
    @classmethod
    def resnet_unit(cls, data, name, filter, stride, dilate, proj, norm_type, norm_mom=0.9, ndev=None):
        return cls.unit(data, name, filter, stride, dilate, proj, normalizer_factory(norm_type, ndev=ndev, mom=norm_mom))

    @classmethod
    def resnet_c1(cls, data, use_3x3_conv0, use_bn_preprocess, norm_type, norm_mom=0.9, ndev=None):
        if use_3x3_conv0 or use_bn_preprocess:
            raise NotImplementedError("3x3 conv0 / BN preprocessing stems are not in the stand-in")
        norm = normalizer_factory(norm_type, ndev=ndev, mom=norm_mom)
        return max_pool(convnormrelu(norm, data, "conv0", 64, kernel=7, stride=2), name="pool0", kernel=3, stride=2)

    @classmethod
    def resnet_c2(cls, data, num_block, stride, dilate, norm_type, norm_mom=0.9, ndev=None):
        return cls.resnet_stage(data, "stage1", num_block, 256, stride, dilate, norm_type, norm_mom, ndev)

    @classmethod
    def resnet_c3(cls, data, num_block, stride, dilate, norm_type, norm_mom=0.9, ndev=None):
        return cls.resnet_stage(data, "stage2", num_block, 512, stride, dilate, norm_type, norm_mom, ndev)

    @classmethod
    def resnet_c4(cls, data, num_block, stride, dilate, norm_type, norm_mom=0.9, ndev=None):
        return cls.resnet_stage(data, "stage3", num_block, 1024, stride, dilate, norm_type, norm_mom, ndev)

    @classmethod
    def resnet_c5(cls, data, num_block, stride, dilate, norm_type, norm_mom=0.9, ndev=None):
        return cls.resnet_stage(data, "stage4", num_block, 2048, stride, dilate, norm_type, norm_mom, ndev)

    def get_backbone(self, variant, depth, endpoint, normalizer, fp16):
        units = self.depth_config[depth]
        data = var("data")
        if fp16:
            data = to_fp16(data, "data_fp16")
        c1 = convnormrelu(normalizer, data, "conv0", 64, kernel=7, stride=2)
        c1 = max_pool(c1, name="pool0", kernel=3, stride=2)
        c2 = self.stage(c1, "stage1", units[0], 256, 1, 1, normalizer)
        c3 = self.stage(c2, "stage2", units[1], 512, 2, 1, normalizer)
        c4 = self.stage(c3, "stage3", units[2], 1024, 2, 1, normalizer)
        if endpoint == "c4":
            return c4
        c5 = self.stage(c4, "stage4", units[3], 2048, 2, 1, normalizer)
        if endpoint == "c5":
            return c5
        if endpoint == "fpn":
            return c2, c3, c4, c5
        raise NotImplementedError(endpoint)

    # the reference also calls these through thin wrappers
    def get_stage_endpoints(self, *a, **kw):
        return self.get_backbone(*a, **kw)


# ---- mxnext.backbone.resnet_v2 ---------------------------------------------------------------------------------------------
class ResNetV2Builder:
    """Pre-activation ResNet (He et al. 2016, the layout of MXNet's resnet-v2 checkpoints: bn -> relu -> conv three
    times, the projection shortcut taken from the first activation; parameter names stage{s}_unit{u}_bn{1,2,3},
    _conv{1,2,3}, _sc; stem bn_data / conv0 / bn0; the closing bn1 + relu1 belong to whoever consumes the last
    stage, symbol/builder.py:571-572).  Interface from the call sites: symbol/builder.py:564-574,655-660 and the
    piecewise methods TridentNet's builder overrides (models/tridentnet/resnet_v2.py:195-265)."""
    depth_config = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3), 200: (3, 24, 36, 3)}

    @classmethod
    def resnet_unit(cls, data, name, filter, stride, dilate, proj, norm_type, norm_mom=0.9, ndev=None):
        norm = normalizer_factory(norm_type, ndev=ndev, mom=norm_mom)
        a1 = relu(norm(data, name=name + "_bn1"), name=name + "_relu1")
        c1 = conv(a1, name=name + "_conv1", filter=filter // 4)
        a2 = relu(norm(c1, name=name + "_bn2"), name=name + "_relu2")
        c2 = conv(a2, name=name + "_conv2", filter=filter // 4, kernel=3, stride=stride, dilate=dilate)
        a3 = relu(norm(c2, name=name + "_bn3"), name=name + "_relu3")
        c3 = conv(a3, name=name + "_conv3", filter=filter)
        sc = conv(a1, name=name + "_sc", filter=filter, stride=stride) if proj else data
        return add(c3, sc, name=name + "_plus")

    @classmethod
    def resnet_stage(cls, data, name, num_block, filter, stride, dilate, norm_type, norm_mom=0.9, ndev=None, **kw):
        data = cls.resnet_unit(data, f"{name}_unit1", filter, stride, dilate, True, norm_type, norm_mom, ndev)
        for i in range(2, num_block + 1):
            data = cls.resnet_unit(data, f"{name}_unit{i}", filter, 1, dilate, False, norm_type, norm_mom, ndev)
        return data

    @classmethod
    def resnet_c1(cls, data, use_3x3_conv0, use_bn_preprocess, norm_type, norm_mom=0.9, ndev=None):
        if use_3x3_conv0:
            raise NotImplementedError("3x3 conv0 stem is not in the stand-in")
        norm = normalizer_factory(norm_type, ndev=ndev, mom=norm_mom)
        if use_bn_preprocess:
            data = sym.BatchNorm(data=data, name="bn_data", use_global_stats=True, fix_gamma=True, eps=2e-5)
        c = relu(norm(conv(data, name="conv0", filter=64, kernel=7, stride=2), name="bn0"), name="relu0")
        return pool(c, name="pool0", kernel=3, stride=2, pad=1, pool_type="max")

    @classmethod
    def resnet_c2(cls, data, num_block, stride, dilate, norm_type, norm_mom=0.9, ndev=None):
        return cls.resnet_stage(data, "stage1", num_block, 256, stride, dilate, norm_type, norm_mom, ndev)

    @classmethod
    def resnet_c3(cls, data, num_block, stride, dilate, norm_type, norm_mom=0.9, ndev=None):
        return cls.resnet_stage(data, "stage2", num_block, 512, stride, dilate, norm_type, norm_mom, ndev)

    @classmethod
    def resnet_c4(cls, data, num_block, stride, dilate, norm_type, norm_mom=0.9, ndev=None):
        return cls.resnet_stage(data, "stage3", num_block, 1024, stride, dilate, norm_type, norm_mom, ndev)

    @classmethod
    def resnet_c5(cls, data, num_block, stride, dilate, norm_type, norm_mom=0.9, ndev=None):
        return cls.resnet_stage(data, "stage4", num_block, 2048, stride, dilate, norm_type, norm_mom, ndev)

    def get_backbone(self, variant, depth, endpoint, normalizer, fp16):
        n2, n3, n4, n5 = self.depth_config[depth]
        data = var("data")
        if fp16:
            data = to_fp16(data, "data_fp16")
        c1 = self.resnet_c1(data, False, variant == "mxnet", normalizer)   # MXNet's checkpoints carry bn_data
        c2 = self.resnet_c2(c1, n2, 1, 1, normalizer)
        c3 = self.resnet_c3(c2, n3, 2, 1, normalizer)
        c4 = self.resnet_c4(c3, n4, 2, 1, normalizer)
        if endpoint == "c4":
            return c4
        c5 = self.resnet_c5(c4, n5, 2, 1, normalizer)
        if endpoint == "c5":
            return c5
        if endpoint == "c4c5":
            return c4, c5
        raise NotImplementedError(endpoint)


# ---- mxnext.backbone.resnet_v1b / resnet_v1b_helper ----------------------------------------------------------------------
class resnet_v1b_helper:
    """ResNet-v1b (stride on the 3x3 convolution).  Interface from the call sites in models/dcn/builder.py:56-110 and
    symbol/builder.py:745-775; unit structure and parameter names are those of `dcn_resnet_unit`
    (models/dcn/builder.py:8-36), which is this unit with conv2 swapped for a deformable convolution."""
    depth_config = {18: (2, 2, 2, 2), 34: (3, 4, 6, 3), 50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3),
                    200: (3, 24, 36, 3)}

    @staticmethod
    def resnet_unit(input, name, filter, stride, dilate, proj, norm, **kw):
        c1 = relu(norm(conv(input, name=name + "_conv1", filter=filter // 4), name=name + "_bn1"), name=name + "_relu1")
        c2 = relu(norm(conv(c1, name=name + "_conv2", filter=filter // 4, kernel=3, stride=stride, dilate=dilate),
                       name=name + "_bn2"), name=name + "_relu2")
        c3 = norm(conv(c2, name=name + "_conv3", filter=filter), name=name + "_bn3")
        sc = norm(conv(input, name=name + "_sc", filter=filter, stride=stride), name=name + "_sc_bn") if proj else input
        return relu(add(c3, sc, name=name + "_plus"), name=name + "_relu")

    @classmethod
    def resnet_stage(cls, data, name, num_block, filter, stride, dilate, norm, **kw):
        for i in range(1, num_block + 1):
            data = cls.resnet_unit(data, f"{name}_unit{i}", filter, stride if i == 1 else 1, dilate, i == 1, norm)
        return data

    @staticmethod
    def resnet_c1(data, norm):
        c = relu(norm(conv(data, name="conv0", filter=64, kernel=7, stride=2), name="bn0"), name="relu0")
        return pool(c, name="pool0", kernel=3, stride=2, pad=1, pool_type="max")

    @classmethod
    def resnet_c2(cls, data, num_block, stride, dilate, norm):
        return cls.resnet_stage(data, "stage1", num_block, 256, stride, dilate, norm)

    @classmethod
    def resnet_c3(cls, data, num_block, stride, dilate, norm):
        return cls.resnet_stage(data, "stage2", num_block, 512, stride, dilate, norm)

    @classmethod
    def resnet_c4(cls, data, num_block, stride, dilate, norm):
        return cls.resnet_stage(data, "stage3", num_block, 1024, stride, dilate, norm)

    @classmethod
    def resnet_c5(cls, data, num_block, stride, dilate, norm):
        return cls.resnet_stage(data, "stage4", num_block, 2048, stride, dilate, norm)


class resnet_v1_helper(resnet_v1b_helper):
    """The same helper interface for ResNet-v1 in the MSRA layout (stride on the first 1x1 convolution of a unit):
    `from mxnext.backbone import resnet_v1_helper` (models/tridentnet/builder_v2.py:4,183)."""

    @staticmethod
    def resnet_unit(input, name, filter, stride, dilate, proj, norm, **kw):
        c1 = relu(norm(conv(input, name=name + "_conv1", filter=filter // 4, stride=stride), name=name + "_bn1"),
                  name=name + "_relu1")
        c2 = relu(norm(conv(c1, name=name + "_conv2", filter=filter // 4, kernel=3, dilate=dilate), name=name + "_bn2"),
                  name=name + "_relu2")
        c3 = norm(conv(c2, name=name + "_conv3", filter=filter), name=name + "_bn3")
        sc = norm(conv(input, name=name + "_sc", filter=filter, stride=stride), name=name + "_sc_bn") if proj else input
        return relu(add(c3, sc, name=name + "_plus"), name=name + "_relu")


class ResNetV1bBuilder:
    def get_backbone(self, variant, depth, endpoint, normalizer, fp16):
        h = resnet_v1b_helper
        n2, n3, n4, n5 = h.depth_config[depth]
        data = var("data")
        if fp16:
            data = to_fp16(data, "data_fp16")
        c1 = h.resnet_c1(data, normalizer)
        c2 = h.resnet_c2(c1, n2, 1, 1, normalizer)
        c3 = h.resnet_c3(c2, n3, 2, 1, normalizer)
        c4 = h.resnet_c4(c3, n4, 2, 1, normalizer)
        if endpoint == "c4":
            return c4
        c5 = h.resnet_c5(c4, n5, 2, 1, normalizer)
        if endpoint == "c5":
            return c5
        if endpoint == "c4c5":
            return c4, c5
        if endpoint == "fpn":
            return c2, c3, c4, c5
        raise NotImplementedError(endpoint)

    resnet_stage = resnet_v1b_helper.resnet_stage


# ---- mxnext.tvm.* -----------------------------------------------------------------------------------------------------------
def tvm_proposal(cls_prob, bbox_pred, im_info, anchors=None, name="proposal", feature_stride=16, scales=(8,),
                 ratios=(0.5, 1.0, 2.0), rpn_pre_nms_top_n=1000, rpn_post_nms_top_n=1000, threshold=0.7, batch_size=1,
                 max_side=-1, output_score=True, variant="simpledet", rpn_min_size=0, **kw):
    """mxnext.tvm.proposal.proposal(variant="simpledet"): the TVM-built twin of _contrib_Proposal_v3 used for the
    large levels (models/FPN/builder.py:289-311).  The stand-in emits _contrib_Proposal_v3 itself - the anchors input
    carries no information beyond (stride, scales, ratios)."""
    return contrib.Proposal_v3(cls_prob=cls_prob, bbox_pred=bbox_pred, im_info=im_info,
                               rpn_pre_nms_top_n=rpn_pre_nms_top_n, rpn_post_nms_top_n=rpn_post_nms_top_n,
                               feature_stride=feature_stride, output_score=output_score, scales=tuple(scales),
                               ratios=tuple(ratios), rpn_min_size=rpn_min_size, threshold=threshold, iou_loss=False,
                               name=f"{name}_stride{feature_stride}")


def tvm_decode_bbox(F, rois, bbox_pred, im_info, bbox_mean=(0.0, 0.0, 0.0, 0.0), bbox_std=(0.1, 0.1, 0.2, 0.2),
                    class_agnostic=True, name=None, **kw):
    """mxnext.tvm.decode_bbox.decode_bbox (models/FreeAnchor/ops.py:166,267): the TVM-built twin of
    _contrib_DecodeBBox; the stand-in emits the operator itself."""
    return contrib.DecodeBBox(rois=rois, bbox_pred=bbox_pred, im_info=im_info, bbox_mean=tuple(bbox_mean),
                              bbox_std=tuple(bbox_std), class_agnostic=bool(class_agnostic), name=name)


def tvm_get_top_proposal(F, bbox, score, top_n, batch_size=1, **kw):
    return sym.Custom(bbox=bbox, score=score, op_type="get_top_proposal", top_n=top_n, name="get_top_proposal")


def tvm_fpn_roi_assign(F, rois, rcnn_stride, roi_canonical_scale, roi_canonical_level, **kw):
    return sym.Custom(rois=rois, op_type="assign_layer_fpn", rcnn_stride=tuple(rcnn_stride),
                      roi_canonical_scale=roi_canonical_scale, roi_canonical_level=roi_canonical_level,
                      name="assign_layer_fpn")
