"""`mx` / `mxnext` façade: stand-in modules named `mxnet` and `mxnext`, large enough that the
reference's own `config/*.py`, `symbol/builder.py` and `models/*/builder.py` import and BUILD THEIR GRAPHS UNCHANGED,
and a torch executor that runs those graphs - library conv / GEMM (cuDNN / cuBLAS) for the backbone and heads,
`simpledet_b200.ops.OPS` (the C ABI) for every detection operator, looked up by the registration string the
builder put into the graph.

    from simpledet_b200 import facade
    facade.install(reference_root="/root/reference")      # puts the stand-ins into sys.modules
    import importlib; cfg = importlib.import_module("config.faster_r50v1_fpn_1x")
    sym = cfg.get_config(is_train=False)[6].test_symbol    # the reference's own builder code ran
    exe = facade.Executor(sym, device="cuda:0")
    outs = exe.forward(data=..., im_info=..., im_id=..., rec_id=...)

Nothing of MXNet's runtime is imitated beyond what graph construction and `DetModule`'s inference calls touch."""
from __future__ import annotations

import sys
import types

from . import mxnext_impl as X
from . import symbol as S
from .executor import Executor, infer_shapes  # noqa: F401


def _mod(modname, **attrs):
    name = modname
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Initializer:
    def __init__(self, *a, **kw):
        self.args, self.kw = a, kw

    def dumps(self):
        import json

        return json.dumps([type(self).__name__.lower(), self.kw])

    def __call__(self, desc, arr):
        pass


class Context:
    def __init__(self, device_type, device_id=0):
        self.device_type, self.device_id = device_type, device_id

    def __repr__(self):
        return f"{self.device_type}({self.device_id})"

    def __eq__(self, o):
        return isinstance(o, Context) and (o.device_type, o.device_id) == (self.device_type, self.device_id)

    def __hash__(self):
        return hash((self.device_type, self.device_id))


def install(reference_root: str | None = None):
    """Register the stand-in `mxnet` / `mxnext` packages (idempotent).  `reference_root` is appended to sys.path so the
    reference's `config`, `symbol`, `models`, `core`, `utils` packages import from where they lie."""
    if reference_root and reference_root not in sys.path:
        sys.path.append(reference_root)
    if getattr(sys.modules.get("mxnet"), "__simpledet_facade__", False):
        return sys.modules["mxnet"]
    from . import ndarray as nd_mod
    from . import module as module_mod

    sym_ns = S._OpNamespace("")

    class _SymModule(types.ModuleType):
        """mx.sym / mx.symbol: Variable, var, Group, contrib.*, and any operator by name."""

        def __getattr__(self, op):
            return getattr(sym_ns, op)

    symm = _SymModule("mxnet.symbol")
    symm.Variable = symm.var = S.Variable
    symm.Group = S.Group
    symm.Symbol = S.Symbol
    symm.contrib = S._OpNamespace("_contrib_")
    # builders probe the operator list (models/retinanet/builder.py:357 picks GenProposalRetina over the CustomOp)
    symm.contrib.__all__ = ["ROIAlign_v2", "DecodeBBox", "Proposal", "Proposal_v2", "Proposal_v3", "NMS", "GenAnchor",
                            "GenProposal", "GenProposalRetina", "FocalLoss", "BBoxNorm", "SigmoidCrossEntropy",
                            "DeformableConvolution", "ModulatedDeformableConvolution"]
    symm.load = lambda fname: (_ for _ in ()).throw(NotImplementedError("symbol.load"))
    sys.modules["mxnet.symbol"] = sys.modules["mxnet.sym"] = symm

    init = _mod("mxnet.initializer", Initializer=_Initializer, InitDesc=nd_mod.InitDesc,
                **{n: type(n, (_Initializer,), {}) for n in ("Xavier", "Normal", "Uniform", "Zero", "One", "Constant",
                                                             "MSRAPrelu", "Orthogonal", "Load", "Mixed")})
    init.register = lambda cls: cls
    io = _mod("mxnet.io", DataBatch=nd_mod.DataBatch, DataDesc=nd_mod.DataDesc, DataIter=object)
    ctx_mod = _mod("mxnet.context", Context=Context, cpu=lambda i=0: Context("cpu", i), gpu=lambda i=0: Context("gpu", i),
                   current_context=lambda: Context("cpu", 0))
    operator = _mod("mxnet.operator", CustomOp=nd_mod.CustomOp, CustomOpProp=nd_mod.CustomOpProp,
                    register=nd_mod.register_custom)
    metric = _mod("mxnet.metric", EvalMetric=type("EvalMetric", (), {"__init__": lambda self, *a, **k: None}),
                  CompositeEvalMetric=type("CompositeEvalMetric", (), {"__init__": lambda self, *a, **k: None}),
                  create=lambda *a, **k: None)
    optimizer = _mod("mxnet.optimizer", Optimizer=object, create=lambda *a, **k: None, get_updater=lambda *a, **k: None)
    base = _mod("mxnet.base", _as_list=lambda x: list(x) if isinstance(x, (list, tuple)) else [x], MXNetError=RuntimeError)
    model = _mod("mxnet.model", _create_kvstore=None, _initialize_kvstore=None, _update_params=None,
                 _update_params_on_kvstore=None, load_checkpoint=module_mod.load_checkpoint,
                 save_checkpoint=lambda *a, **k: None, BatchEndParam=None)
    callback = _mod("mxnet.callback", module_checkpoint=lambda *a, **k: None, Speedometer=lambda *a, **k: None)
    profiler = _mod("mxnet.profiler", set_state=lambda *a, **k: None, dump=lambda *a, **k: None,
                    set_config=lambda *a, **k: None)
    lr_sched = _mod("mxnet.lr_scheduler", LRScheduler=object)
    kv = _mod("mxnet.kvstore", KVStore=object, create=lambda *a, **k: None)
    ndm = _mod("mxnet.ndarray", **nd_mod.exports())
    ndm.contrib = types.SimpleNamespace()
    mod_pkg = _mod("mxnet.module")
    mod_pkg.__path__ = []
    base_module = _mod("mxnet.module.base_module", BaseModule=module_mod.BaseModule,
                       _check_input_names=module_mod._check_input_names, _parse_data_desc=module_mod._parse_data_desc)
    exec_group = _mod("mxnet.module.executor_group", DataParallelExecutorGroup=module_mod.DataParallelExecutorGroup)
    module_m = _mod("mxnet.module.module", Module=module_mod.BaseModule)
    mod_pkg.base_module, mod_pkg.executor_group, mod_pkg.module = base_module, exec_group, module_m
    mod_pkg.BaseModule, mod_pkg.Module = module_mod.BaseModule, module_mod.BaseModule

    mx = _mod("mxnet", sym=symm, symbol=symm, nd=ndm, ndarray=ndm, init=init, initializer=init, io=io, context=ctx_mod,
              Context=Context, cpu=ctx_mod.cpu, gpu=ctx_mod.gpu, operator=operator, metric=metric, optimizer=optimizer,
              base=base, model=model, callback=callback, profiler=profiler, lr_scheduler=lr_sched, kvstore=kv, kv=kv,
              module=mod_pkg, mod=mod_pkg, MXNetError=RuntimeError, __version__="1.6.0-simpledet_b200-facade",
              AttrScope=module_mod.AttrScope, name=types.SimpleNamespace(Prefix=module_mod.AttrScope))
    mx.__path__ = []
    mx.__simpledet_facade__ = True
    mx.contrib = types.SimpleNamespace(symbol=symm.contrib, sym=symm.contrib)   # mx.contrib.symbol.DeformableConvolution

    # ---- mxnext
    simple = {k: getattr(X, k) for k in dir(X) if not k.startswith("_") and callable(getattr(X, k))}
    mxn = _mod("mxnext", **simple)
    mxn.__path__ = []
    mxn.sym, mxn.contrib = X.sym, X.contrib
    _mod("mxnext.simple", **simple)
    _mod("mxnext.complicate", normalizer_factory=X.normalizer_factory, convrelu=X.convrelu, convnormrelu=X.convnormrelu,
         convnorm=X.convnorm)
    bb = _mod("mxnext.backbone")
    bb.__path__ = []
    _mod("mxnext.backbone.resnet_v1", Builder=X.ResNetV1Builder)
    _mod("mxnext.backbone.resnet_v1b", Builder=X.ResNetV1bBuilder)
    _mod("mxnext.backbone.resnet_v2", Builder=X.ResNetV2Builder)
    for hname, h in (("resnet_v1b_helper", X.resnet_v1b_helper), ("resnet_v1_helper", X.resnet_v1_helper)):
        setattr(bb, hname, _mod("mxnext.backbone." + hname, depth_config=h.depth_config, resnet_unit=h.resnet_unit,
                                resnet_stage=h.resnet_stage, resnet_c1=h.resnet_c1, resnet_c2=h.resnet_c2,
                                resnet_c3=h.resnet_c3, resnet_c4=h.resnet_c4, resnet_c5=h.resnet_c5))
    tvm = _mod("mxnext.tvm")
    tvm.__path__ = []
    _mod("mxnext.tvm.proposal", proposal=X.tvm_proposal)
    _mod("mxnext.tvm.decode_bbox", decode_bbox=X.tvm_decode_bbox)
    _mod("mxnext.tvm.get_top_proposal", get_top_proposal=X.tvm_get_top_proposal)
    _mod("mxnext.tvm.fpn_roi_assign", fpn_roi_assign=X.tvm_fpn_roi_assign)

    # ---- pycocotools: the mask detectors import it at module top (models/maskrcnn/process_output.py:2) for the
    # test-time RLE encoding.  Not in this image: an import-time stand-in whose functions say so when called.
    import importlib.util as _ilu

    if _ilu.find_spec("pycocotools") is None:
        def _absent(*a, **k):
            raise ImportError("pycocotools is not installed: RLE encoding / COCO evaluation are outside this package")

        pct = _mod("pycocotools")
        pct.__path__ = []
        pct.mask = _mod("pycocotools.mask", encode=_absent, decode=_absent, frPyObjects=_absent, merge=_absent, area=_absent)
        pct.coco = _mod("pycocotools.coco", COCO=_absent)
        pct.cocoeval = _mod("pycocotools.cocoeval", COCOeval=_absent)

    # ---- the reference's compiled helpers (operator_py/cython/*.pyx): numpy-facing drop-ins over the device ops
    def _bbox_overlaps_cython(boxes, query_boxes):
        import numpy as np
        import torch

        from .. import ops

        dev = torch.device("cuda", torch.cuda.current_device())
        o = ops.bbox_overlaps(torch.from_numpy(np.ascontiguousarray(boxes, np.float32)).to(dev),
                              torch.from_numpy(np.ascontiguousarray(query_boxes, np.float32)).to(dev))
        return o.cpu().numpy()

    def _bbox_selfoverlaps_cython(boxes, query_boxes):
        import numpy as np
        import torch

        from .. import ops

        dev = torch.device("cuda", torch.cuda.current_device())
        o = ops.bbox_overlaps(torch.from_numpy(np.ascontiguousarray(boxes, np.float32)).to(dev),
                              torch.from_numpy(np.ascontiguousarray(query_boxes, np.float32)).to(dev), mode="ioa")
        return o.cpu().numpy()

    def _soft_nms(boxes_in, sigma=0.5, Nt=0.3, threshold=0.001, method=0):
        import numpy as np
        import torch

        from .. import ops

        dev = torch.device("cuda", torch.cuda.current_device())
        b, i = ops.soft_nms(torch.from_numpy(np.ascontiguousarray(boxes_in, np.float32)).to(dev), sigma, Nt, threshold, method)
        return b.cpu().numpy(), i.cpu().numpy()

    def _greedy_nms(dets, thresh):
        import numpy as np
        import torch

        from .. import ops

        dev = torch.device("cuda", torch.cuda.current_device())
        return ops.greedy_nms(torch.from_numpy(np.ascontiguousarray(dets, np.float32)).to(dev), thresh).cpu().numpy()

    if "operator_py.cython.bbox" not in sys.modules:
        import importlib.util

        if importlib.util.find_spec("operator_py") is not None:
            import operator_py  # the reference's package (pure-Python modules import from where they lie)

            cy = _mod("operator_py.cython")
            cy.__path__ = []
            _mod("operator_py.cython.bbox", bbox_overlaps_cython=_bbox_overlaps_cython)
            _mod("operator_py.cython.bbox_self", bbox_selfoverlaps_cython=_bbox_selfoverlaps_cython)
            _mod("operator_py.cython.cpu_nms", soft_nms=_soft_nms, greedy_nms=_greedy_nms, cpu_nms=_greedy_nms)
            _mod("operator_py.cython.gpu_nms", gpu_nms=lambda dets, thresh, device_id=0: __import__(
                "simpledet_b200.ops", fromlist=["gpu_nms"]).gpu_nms(dets, thresh, device_id))
            operator_py.cython = cy
    return mx
