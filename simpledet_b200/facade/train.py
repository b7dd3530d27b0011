"""Training side of the façade executor: the loss operators of the reference's graphs with MXNet's backward
semantics, and a `Trainer` that runs forward + backward of a TRAIN symbol the reference's builders produced, reduces
the gradients across ranks in ONE flat fp32 bucket (core/detection_module.py:680-690: kvstore push/pull per step;
north star: "a single NCCL allreduce over NVLink for gradients") and applies MXNet's SGD update
(detection_train.py:255-275).

MXNet's loss operators do not propagate the incoming gradient: SoftmaxOutput's backward is `(softmax - onehot) *
grad_scale / norm` and MakeLoss's is `grad_scale / norm`, whatever sits above them; `backward()` on the executor seeds
every head with ones.  They are restated here as torch.autograd Functions from upstream MXNet 1.6
(src/operator/softmax_output-inl.h SoftmaxOutputOp::Backward, make_loss-inl.h; the sources are NOT under
/root/reference - parity unpinned, restated from the published operator):
  SoftmaxOutput  multi_output: softmax over axis 1 of (n, k, rest); grad = softmax - onehot(label), zero where
                 label == ignore_label (use_ignore); scale = grad_scale / valid_cnt ('valid': labels != ignore_label,
                 at least 1), grad_scale / rest / n ('batch'), grad_scale / rest ('null').
                 otherwise (n, k): scale = grad_scale / n ('batch'), / #(label != ignore) ('valid'), / 1 ('null').
  MakeLoss       grad = grad_scale ('null'), / n ('batch'), / #(data > valid_thresh) ('valid').
  smooth_l1      sigma = scalar: 0.5 (sigma x)^2 if |x| < 1/sigma^2 else |x| - 0.5/sigma^2 (plain autograd).
  BlockGrad      detach.
"""
from __future__ import annotations

import torch


class _SoftmaxOutputFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, data, label, multi_output, normalization, use_ignore, ignore_label, grad_scale):
        if multi_output:
            n, k = data.shape[0], data.shape[1]
            d3 = data.reshape(n, k, -1)
            prob = torch.softmax(d3, 1)
        else:
            d3 = data.reshape(data.shape[0], -1)
            prob = torch.softmax(d3, 1)
        ctx.save_for_backward(prob, label)
        ctx.cfg = (bool(multi_output), str(normalization), bool(use_ignore), float(ignore_label), float(grad_scale),
                   tuple(data.shape))
        return prob.reshape(data.shape)

    @staticmethod
    def backward(ctx, _grad_out):  # out_grad=False: the head gradient is ignored
        prob, label = ctx.saved_tensors
        multi, norm, use_ignore, ignore, gscale, shape = ctx.cfg
        if multi:
            n, k, rest = prob.shape
            lab = label.reshape(n, rest).to(torch.long)
            keep = (lab != int(ignore)) if use_ignore else torch.ones_like(lab, dtype=torch.bool)
            onehot = torch.zeros_like(prob).scatter_(1, lab.clamp(0, k - 1).unsqueeze(1), 1.0)
            grad = (prob - onehot) * keep.unsqueeze(1).to(prob.dtype)
            if norm == "valid":
                # MXNet counts labels != ignore_label whether or not use_ignore is set
                cnt = (label.reshape(-1).to(torch.long) != int(ignore)).sum().clamp(min=1).to(prob.dtype)
                grad = grad * (gscale / cnt)
            elif norm == "batch":
                grad = grad * (gscale / rest / n)
            else:
                grad = grad * (gscale / rest)
        else:
            n, k = prob.shape
            lab = label.reshape(n).to(torch.long)
            keep = (lab != int(ignore)) if use_ignore else torch.ones_like(lab, dtype=torch.bool)
            onehot = torch.zeros_like(prob).scatter_(1, lab.clamp(0, k - 1).unsqueeze(1), 1.0)
            grad = (prob - onehot) * keep.unsqueeze(1).to(prob.dtype)
            if norm == "valid":
                cnt = (lab != int(ignore)).sum().clamp(min=1).to(prob.dtype)
                grad = grad * (gscale / cnt)
            elif norm == "batch":
                grad = grad * (gscale / n)
            else:
                grad = grad * gscale
        return grad.reshape(shape), None, None, None, None, None, None


def softmax_output(data, label, multi_output=False, normalization="null", use_ignore=False, ignore_label=-1.0,
                   grad_scale=1.0):
    return _SoftmaxOutputFn.apply(data, label.detach(), multi_output, normalization, use_ignore, ignore_label, grad_scale)


class _MakeLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, data, grad_scale, normalization, valid_thresh):
        ctx.save_for_backward(data)
        ctx.cfg = (float(grad_scale), str(normalization), float(valid_thresh))
        return data.view_as(data)

    @staticmethod
    def backward(ctx, _grad_out):
        (data,) = ctx.saved_tensors
        gscale, norm, thresh = ctx.cfg
        if norm == "valid":
            cnt = (data > thresh).sum().clamp(min=1).to(data.dtype)
            return torch.ones_like(data) * (gscale / cnt), None, None, None
        if norm == "batch":
            return torch.full_like(data, gscale / data.shape[0]), None, None, None
        return torch.full_like(data, gscale), None, None, None


def make_loss(data, grad_scale=1.0, normalization="null", valid_thresh=0.0):
    return _MakeLossFn.apply(data, grad_scale, normalization, valid_thresh)


def smooth_l1(x, scalar=1.0):
    s2 = float(scalar) * float(scalar)
    ax = x.abs()
    return torch.where(ax < 1.0 / s2, 0.5 * s2 * x * x, ax - 0.5 / s2)


class Trainer:
    """One data-parallel worker of the reference's training loop over a façade TRAIN symbol.

        tr = Trainer(train_sym, input_shapes, device, fixed_param=("conv0", "stage1", "gamma", "beta"),
                     label_names=("rpn_cls_label", "rpn_reg_target", "rpn_reg_weight"))
        outs = tr.forward_backward(data=..., im_info=..., gt_bbox=..., rpn_cls_label=..., ...)
        tr.allreduce_grads()          # one flat fp32 all-reduce when torch.distributed is initialised
        tr.update(lr, momentum=0.9, wd=1e-4, rescale_grad=1 / world)

    fixed_param follows core/detection_module.py:102-107 (substring match on argument names); auxiliary states
    (moving statistics) never receive gradients.  The update is MXNet's sgd_mom_update: g = clip(rescale_grad * grad);
    mom = momentum * mom - lr * lr_mult * (g + wd * wd_mult * w); w += mom (lr_mult / wd_mult from the Variable's
    attributes, as mx.optimizer reads them from the symbol)."""

    def __init__(self, sym, input_shapes, device="cuda:0", fixed_param=(), rng_std=0.01, arg_params=None,
                 aux_params=None, channels_last=False, label_names=()):
        from .executor import Executor

        self.ex = Executor(sym, device=device, channels_last=channels_last, fold_bn=False, is_train=True)
        self.ex.init_params(input_shapes, arg_params=arg_params, aux_params=aux_params, rng_std=rng_std)
        # data_names are the keys of input_shapes; label_names (the config's `label_name` list, e.g. rpn_cls_label,
        # rpn_reg_target, rpn_reg_weight) have their shapes inferred from the graph like MXNet's bind does, and are
        # inputs, not parameters: forward_backward() must be fed them
        self.input_names = set(input_shapes) | set(label_names)
        for n in label_names:
            if n not in self.ex.params:
                raise KeyError(f"label {n!r} is not an argument of the symbol")
            del self.ex.params[n]
        aux = set(sym.list_auxiliary_states())
        self.trainable = [n for n in sym.list_arguments()
                          if n in self.ex.params and n not in aux and not any(f in n for f in fixed_param)]
        for n in self.trainable:
            self.ex.params[n].requires_grad_(True)
        self._mult = {}
        for node in sym._topo():
            if node.op is None and node.name in self.ex.params:
                self._mult[node.name] = (float(node.attrs.get("__lr_mult__", 1.0)), float(node.attrs.get("__wd_mult__", 1.0)))
        self._mom = {}
        self._flat = None

    def forward_backward(self, **inputs):
        for n in self.trainable:
            self.ex.params[n].grad = None
        outs = self.ex.forward(**inputs)
        heads = [o for o in outs if o.requires_grad]
        torch.autograd.backward(heads, [torch.ones_like(h) for h in heads])
        return [o.detach() for o in outs]

    def grads(self):
        return {n: self.ex.params[n].grad for n in self.trainable if self.ex.params[n].grad is not None}

    def allreduce_grads(self, group=None, async_op=False):
        """Pack every gradient into one flat fp32 bucket, all-reduce (sum) it once, unpack.  Returns the work handle
        when async_op (call .wait() then `unpack()`), else None.  A no-op without an initialised process group."""
        import torch.distributed as dist

        g = self.grads()
        if not g or not (dist.is_available() and dist.is_initialized()):
            return None
        names = sorted(g)
        total = sum(g[n].numel() for n in names)
        if self._flat is None or self._flat.numel() != total:
            self._flat = torch.empty(total, device=g[names[0]].device, dtype=torch.float32)
        off = 0
        for n in names:
            k = g[n].numel()
            self._flat[off:off + k].copy_(g[n].reshape(-1))
            off += k
        self._bucket_names = names
        work = dist.all_reduce(self._flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            return work
        self.unpack()
        return None

    def unpack(self):
        off = 0
        for n in self._bucket_names:
            gr = self.ex.params[n].grad
            k = gr.numel()
            gr.copy_(self._flat[off:off + k].view_as(gr))
            off += k

    @torch.no_grad()
    def update(self, lr, momentum=0.9, wd=0.0, rescale_grad=1.0, clip_gradient=None):
        for n, gr in self.grads().items():
            w = self.ex.params[n]
            lr_mult, wd_mult = self._mult.get(n, (1.0, 1.0))
            g = gr * rescale_grad
            if clip_gradient is not None and clip_gradient >= 0:
                g = g.clamp(-clip_gradient, clip_gradient)
            step = g + (wd * wd_mult) * w
            if momentum:
                m = self._mom.get(n)
                if m is None:
                    m = self._mom[n] = torch.zeros_like(w)
                m.mul_(momentum).add_(step, alpha=-lr * lr_mult)
                w.add_(m)
            else:
                w.add_(step, alpha=-lr * lr_mult)
