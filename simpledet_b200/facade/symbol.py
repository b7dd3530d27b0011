"""A symbolic-graph stand-in for the slice of `mxnet.symbol` the reference's builders use (symbol/builder.py,
models/FPN/builder.py, ...): nodes record the operator NAME and its keyword attributes exactly as the builders pass
them; nothing is computed here.  `simpledet_b200.facade.executor` evaluates a graph with torch (library conv / GEMM
for the backbone and heads) and `simpledet_b200.ops.OPS` (the detection operators, by their registration strings)."""
from __future__ import annotations

import json
from typing import Any

_UID = [0]


def _auto_name(hint: str) -> str:
    _UID[0] += 1
    return f"{hint.lower()}{_UID[0]}"


class Node:
    __slots__ = ("op", "name", "inputs", "attrs", "num_outputs", "arg_names")

    def __init__(self, op, name, inputs, attrs, num_outputs=1, arg_names=None):
        self.op, self.name, self.inputs, self.attrs = op, name, list(inputs), dict(attrs)
        self.num_outputs, self.arg_names = num_outputs, list(arg_names or [])


# operators that create their parameter variables when the caller gives none (mx.sym.Convolution(data=..., name=...))
_AUTO_ARGS = {
    "Convolution": ("weight", "bias"), "FullyConnected": ("weight", "bias"), "Deconvolution": ("weight", "bias"),
    "BatchNorm": ("gamma", "beta", "moving_mean", "moving_var"),
    "_contrib_DeformableConvolution": ("weight", "bias"), "_contrib_ModulatedDeformableConvolution": ("weight", "bias"),
    "_contrib_SyncBatchNorm": ("gamma", "beta", "moving_mean", "moving_var"), "_contrib_GroupNorm": ("gamma", "beta"),
}
_AUX = {"moving_mean", "moving_var"}
# operators with several outputs: name -> callable(attrs) -> (total outputs, visible outputs)
_MULTI_OUT = {
    "_contrib_Proposal_v3": lambda a: (2, 2 if _truthy(a.get("output_score", False)) else 1),
    "_contrib_Proposal": lambda a: (2, 2 if _truthy(a.get("output_score", False)) else 1),
    "_contrib_Proposal_v2": lambda a: (2, 2 if _truthy(a.get("output_score", False)) else 1),
    "_contrib_NMS": lambda a: (2, 2 if _truthy(a.get("output_score", False)) else 1),
    "_contrib_GenProposalRetina": lambda a: (2, 2),          # generate_proposal_retina-inl.h: bbox, score
    "_contrib_ROIAlign_v2": lambda a: (3, 1),
    "ROIPooling_v1": lambda a: (2, 1),
    "ProposalTarget_v2": lambda a: (5, 5 if _truthy(a.get("output_iou", False)) else 4),
    "ProposalMaskTarget": lambda a: ((lambda n: (n, n))(5 + int(_truthy(a.get("output_iou", False)))
                                                        + int(_truthy(a.get("output_ratio", False))))),
    "ProposalTarget": lambda a: (5, 5 if _truthy(a.get("output_iou", False)) else 4),
    "BatchNorm": lambda a: (1, 1),
    "SliceChannel": lambda a: (int(a["num_outputs"]), int(a["num_outputs"])),
    "split": lambda a: (int(a["num_outputs"]), int(a["num_outputs"])),
}
# non-symbol positional arguments of the creation operators, in order
_POSITIONAL_ATTRS = {"full": ("shape", "value"), "zeros": ("shape",), "ones": ("shape",), "arange": ("start", "stop", "step"),
                     "Reshape": ("shape",), "reshape": ("shape",),
                     "maximum": ("scalar",), "minimum": ("scalar",),   # mx.sym.maximum(sym, 1.): models/msrcnn/builder.py:167
                     "repeat": ("repeats", "axis"), "tile": ("reps",)}   # mx.symbol.Reshape(data, (-3, -2)): models/tridentnet/resnet_v2.py:99
CUSTOM_OUTPUTS = {"get_top_proposal": 2, "assign_layer_fpn": None, "BboxPostProcessing": 3, "bbox_target": 4,
                  "decode_retina": 2}


def _truthy(v):
    return v in (True, 1, "True", "true", "1")


class Symbol:
    """A list of (node, output index) entries."""

    def __init__(self, entries):
        self.entries = list(entries)

    # ---- construction helpers
    @staticmethod
    def var(name, **attrs):
        return Symbol([(Node(None, name, [], attrs), 0)])

    @property
    def name(self):
        return self.entries[0][0].name if len(self.entries) == 1 else None

    def attr(self, key):
        v = self.entries[0][0].attrs.get(key)
        return None if v is None else str(v)

    def list_attr(self):
        return {k: str(v) for k, v in self.entries[0][0].attrs.items()}

    def __iter__(self):
        return (Symbol([e]) for e in self.entries)

    def __len__(self):
        return len(self.entries)

    def __getitem__(self, i):
        if isinstance(i, str):
            names = self.list_outputs()
            return Symbol([self.entries[names.index(i)]])
        if len(self.entries) == 1 and self.entries[0][0].num_outputs > 1:
            node = self.entries[0][0]
            return Symbol([(node, i)])
        return Symbol([self.entries[i]])

    # ---- graph queries
    def _topo(self):
        seen, order = set(), []
        stack = [(n, False) for n, _ in reversed(self.entries)]
        while stack:
            node, done = stack.pop()
            if done:
                order.append(node)
                continue
            if id(node) in seen:
                continue
            seen.add(id(node))
            stack.append((node, True))
            for s in reversed(node.inputs):
                for n, _ in reversed(s.entries):
                    if id(n) not in seen:
                        stack.append((n, False))
        return order

    def _variables(self):
        return [n for n in self._topo() if n.op is None]

    def list_arguments(self):
        return [n.name for n in self._variables() if not n.attrs.get("__aux__")]

    def list_auxiliary_states(self):
        return [n.name for n in self._variables() if n.attrs.get("__aux__")]

    def list_inputs(self):
        return [n.name for n in self._variables()]

    def list_outputs(self):
        out = []
        for node, idx in self.entries:
            if node.op is None:
                out.append(node.name)
            elif node.num_outputs == 1:
                out.append(node.name + "_output")
            else:
                out.append(f"{node.name}_output{idx}")
        return out

    def get_internals(self):
        ent = []
        for n in self._topo():
            for i in range(n.num_outputs):
                ent.append((n, i))
        return Symbol(ent)

    def get_children(self):
        node = self.entries[0][0]
        ent = [e for s in node.inputs for e in s.entries]
        return Symbol(ent) if ent else None

    def infer_shape(self, **kwargs):
        from .executor import infer_shapes

        return infer_shapes(self, kwargs)

    def infer_shape_partial(self, **kwargs):
        return self.infer_shape(**kwargs)

    def infer_type(self, **kwargs):
        import numpy as np

        return ([np.float32] * len(self.list_arguments()), [np.float32] * len(self.entries),
                [np.float32] * len(self.list_auxiliary_states()))

    def tojson(self):
        """MXNet-style graph JSON (nodes / heads); enough to rebuild the graph with `fromjson`."""
        nodes = []
        index = {}
        for n in self._topo():
            index[id(n)] = len(nodes)
            attrs = {}
            for k, v in n.attrs.items():
                if callable(v):
                    continue
                attrs[k] = v.dumps() if hasattr(v, "dumps") else (v if isinstance(v, (int, float, bool, str)) else str(v))
            nodes.append({"op": n.op or "null", "name": n.name, "attrs": attrs, "num_outputs": n.num_outputs,
                          "arg_names": n.arg_names,
                          "inputs": [[[index[id(m)], i] for m, i in s.entries] for s in n.inputs]})
        return json.dumps({"nodes": nodes, "heads": [[index[id(n)], i] for n, i in self.entries]})

    def save(self, fname):
        try:
            with open(fname, "w") as f:
                f.write(self.tojson())
        except OSError:
            pass  # the reference saves next to a checkpoint prefix that may not exist here

    # ---- arithmetic (elementwise with a symbol, scalar ops with a number)
    def _bin(self, other, op, sop, rev=False):
        if isinstance(other, Symbol):
            a, b = (other, self) if rev else (self, other)
            return make_op(op, [a, b], {})
        return make_op(sop, [self], {"scalar": float(other), "__rev__": rev})

    def __add__(self, o): return self._bin(o, "elemwise_add", "_plus_scalar")
    def __radd__(self, o): return self._bin(o, "elemwise_add", "_plus_scalar")
    def __sub__(self, o): return self._bin(o, "elemwise_sub", "_minus_scalar")
    def __rsub__(self, o): return self._bin(o, "elemwise_sub", "_minus_scalar", rev=True)
    def __mul__(self, o): return self._bin(o, "elemwise_mul", "_mul_scalar")
    def __rmul__(self, o): return self._bin(o, "elemwise_mul", "_mul_scalar")
    def __truediv__(self, o): return self._bin(o, "elemwise_div", "_div_scalar")
    def __rtruediv__(self, o): return self._bin(o, "elemwise_div", "_div_scalar", rev=True)
    def __neg__(self): return self._bin(-1.0, "elemwise_mul", "_mul_scalar")
    def __lt__(self, o): return self._bin(o, "_lesser", "_lesser_scalar")
    def __le__(self, o): return self._bin(o, "_lesser_equal", "_lesser_equal_scalar")
    def __gt__(self, o): return self._bin(o, "_greater", "_greater_scalar")
    def __ge__(self, o): return self._bin(o, "_greater_equal", "_greater_equal_scalar")
    def __pow__(self, o): return self._bin(o, "_power", "_power_scalar")
    def __rpow__(self, o): return self._bin(o, "_power", "_rpower_scalar")

    def reshape(self, *shape, **kw):
        """Symbol.reshape(shape) / reshape(d0, d1, ...) / reshape(-1) / reshape(shape=...), as mx.sym.Symbol takes them."""
        if not shape:
            shape = kw["shape"]
        elif len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = shape[0]
        return make_op("Reshape", [self], dict(shape=tuple(int(d) for d in shape)))

    def astype(self, dtype):
        return make_op("Cast", [self], dict(dtype=str(dtype)))


def make_op(op: str, inputs: list, attrs: dict[str, Any], name: str | None = None, arg_names=None) -> Symbol:
    name = name or _auto_name(op.replace("_contrib_", ""))
    total, visible = 1, 1
    if op in _MULTI_OUT:
        total, visible = _MULTI_OUT[op](attrs)
    elif op == "Custom":
        n = CUSTOM_OUTPUTS.get(attrs.get("op_type"))
        if attrs.get("op_type") == "assign_layer_fpn":
            import ast

            n = len(ast.literal_eval(str(attrs["rcnn_stride"])))
        if n is None:  # a CustomOpProp registered by the reference itself (mx.operator.register): ask it
            from . import ndarray as _nd

            prop_cls = _nd._CUSTOM.get(attrs.get("op_type"))
            if prop_cls is not None:
                try:
                    kw = {k: str(v) for k, v in attrs.items() if k != "op_type" and not k.startswith("__")}
                    n = len(prop_cls(**kw).list_outputs())
                except Exception:  # a prop that needs arguments the graph does not carry: single output
                    n = None
        total = visible = n or 1
    node = Node(op, name, inputs, attrs, total, arg_names)
    return Symbol([(node, i) for i in range(visible)]) if visible > 1 else Symbol([(node, 0)])


def Group(symbols):
    ent = []
    for s in symbols:
        ent.extend(s.entries)
    return Symbol(ent)


class _OpNamespace:
    """mx.sym.<Op>(*positional symbols, name=..., **attrs)  ->  Symbol; contrib ops get the `_contrib_` prefix."""

    def __init__(self, prefix=""):
        self._prefix = prefix

    def __getattr__(self, op):
        if op.startswith("__"):
            raise AttributeError(op)
        full = self._prefix + op

        def build(*args, **kwargs):
            name = kwargs.pop("name", None)
            kwargs.pop("attr", None)
            inputs, arg_names, attrs = [], [], {}
            for a in args:
                if isinstance(a, Symbol):
                    inputs.append(a)
                    arg_names.append(None)
                elif isinstance(a, (list, tuple)) and a and all(isinstance(x, Symbol) for x in a):
                    inputs.extend(a)
                    arg_names.extend([None] * len(a))
                elif op in _POSITIONAL_ATTRS:          # creation ops: mx.sym.full(shape, val), zeros(shape), ...
                    attrs[_POSITIONAL_ATTRS[op][len(attrs)]] = a
            for k, v in list(kwargs.items()):
                if isinstance(v, Symbol):
                    inputs.append(v)
                    arg_names.append(k)
                elif v is not None or k in ("weight", "bias"):
                    if v is not None:
                        attrs[k] = v
            name = name or _auto_name(op)
            if full in _AUTO_ARGS:  # missing parameter inputs become variables named <op name>_<arg>
                given = set(arg_names)
                # a variable handed to an auxiliary-state slot (TridentNet shares moving_mean / moving_var between
                # branches by passing them in) is an auxiliary state, as in MXNet: the slot decides, not the creator
                for sym_in, an in zip(inputs, arg_names):
                    if an in _AUX and len(sym_in.entries) == 1 and sym_in.entries[0][0].op is None:
                        sym_in.entries[0][0].attrs["__aux__"] = True
                for an in _AUTO_ARGS[full]:
                    if an in given:
                        continue
                    if an == "bias" and _truthy(attrs.get("no_bias", False)):
                        continue
                    v = Symbol.var(f"{name}_{an}", **({"__aux__": True} if an in _AUX else {}))
                    inputs.append(v)
                    arg_names.append(an)
            return make_op(full, inputs, attrs, name, arg_names)

        return build


def Variable(name, shape=None, lr_mult=None, wd_mult=None, dtype=None, init=None, **kw):
    attrs = {k: v for k, v in dict(__shape__=shape, __lr_mult__=lr_mult, __wd_mult__=wd_mult, __dtype__=dtype,
                                   __init__=init).items() if v is not None}
    attrs.update(kw)
    return Symbol.var(name, **attrs)


def fromjson(text: str) -> Symbol:
    """Rebuild a graph written by Symbol.tojson() (used to carry a graph the reference's builders produced to a
    machine that does not have the reference checkout)."""
    g = json.loads(text)
    nodes = []
    for d in g["nodes"]:
        ins = [Symbol([(nodes[i], o) for i, o in ent]) for ent in d["inputs"]]
        nodes.append(Node(None if d["op"] == "null" else d["op"], d["name"], ins, d["attrs"], d.get("num_outputs", 1),
                          d.get("arg_names")))
    return Symbol([(nodes[i], o) for i, o in g["heads"]])
