"""The slice of `mxnet.module` that core/detection_module.py's DetModule needs for INFERENCE: BaseModule state flags,
name / shape helpers, and a DataParallelExecutorGroup that owns one façade Executor on the first context."""
from __future__ import annotations

import logging

import torch

from . import ndarray as nd
from .executor import Executor, infer_shapes


class AttrScope:
    def __init__(self, *a, **kw):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def load_checkpoint(prefix, epoch):
    raise FileNotFoundError(f"{prefix}-{epoch:04d}.params: checkpoints are not part of the façade")


class BaseModule:
    def __init__(self, logger=logging):
        self.logger = logger
        self.binded = False
        self.for_training = False
        self.inputs_need_grad = False
        self.params_initialized = False
        self.optimizer_initialized = False
        self._symbol = None
        self._total_exec_bytes = 0

    @property
    def symbol(self):
        return self._symbol


def _check_input_names(symbol, names, typename, throw):
    args = symbol.list_arguments()
    for name in names:
        if name in args:
            continue
        msg = f"You created Module with Module(..., {typename}_names={names}) but input with name '{name}' is not found " \
              f"in symbol.list_arguments()."
        if throw:
            raise ValueError(msg)
        logging.warning(msg)


def _parse_data_desc(data_names, label_names, data_shapes, label_shapes):
    data_shapes = [x if isinstance(x, nd.DataDesc) else nd.DataDesc(*x) for x in data_shapes]
    if label_shapes is not None:
        label_shapes = [x if isinstance(x, nd.DataDesc) else nd.DataDesc(*x) for x in label_shapes]
    return data_shapes, label_shapes


class DataParallelExecutorGroup:
    """One executor, the first context.  (The reference's data parallelism is one process driving G GPUs; the
    B200-native layout is one process per GPU - bench.py / torchrun - so a group never holds more than one.)"""

    def __init__(self, symbol, contexts, workload, data_shapes, label_shapes, param_names, for_training,
                 inputs_need_grad, shared_group=None, logger=logging, fixed_param_names=None, grad_req="write",
                 group2ctxs=None, state_names=None):
        if for_training:
            raise NotImplementedError("the façade executor is inference-only")
        self.symbol, self.contexts, self.param_names = symbol, contexts, list(param_names)
        self.data_shapes, self.label_shapes = data_shapes, label_shapes
        self.for_training = False
        ctx = contexts[0]
        dev = torch.device("cuda", ctx.device_id) if ctx.device_type == "gpu" else torch.device("cpu")
        self.exe = Executor(symbol, dev)
        self.execs = [self.exe]
        self.batch_size = data_shapes[0].shape[0]
        self._bind_shapes()
        self._total_exec_bytes = sum(p.numel() * 4 for p in self.exe.params.values())
        self.outputs = None

    def _bind_shapes(self):
        shapes = {d.name: d.shape for d in self.data_shapes}
        self.exe.init_params(shapes)  # zeros, like MXNet's freshly bound executors
        self.param_arrays = [[nd.NDArray(self.exe.params[n])] for n in self.param_names]
        self.aux_names = self.symbol.list_auxiliary_states()
        self.aux_arrays = [[nd.NDArray(self.exe.params[n])] for n in self.aux_names]
        self.grad_arrays = [[None] for _ in self.param_names]
        self._out_shapes = infer_shapes(self.symbol, shapes)[1]

    def reshape(self, data_shapes, label_shapes):
        self.data_shapes, self.label_shapes = data_shapes, label_shapes
        self.batch_size = data_shapes[0].shape[0]
        self._run = None

    def set_params(self, arg_params, aux_params, allow_extra=False):
        for name, arr in list((arg_params or {}).items()) + list((aux_params or {}).items()):
            if name in self.exe.params:
                self.exe.params[name].copy_(arr.t if isinstance(arr, nd.NDArray) else torch.as_tensor(arr))
            elif not allow_extra:
                raise ValueError(f"Find name '{name}' that is not in the arguments")
        self.exe._folded = None  # folded BatchNorm weights and the recorded graph were built from the old values
        self._run = None

    def get_params(self, arg_params, aux_params):
        for n in self.param_names:
            arg_params[n] = nd.NDArray(self.exe.params[n].clone())
        for n in self.aux_names:
            aux_params[n] = nd.NDArray(self.exe.params[n].clone())

    def forward(self, data_batch, is_train=None):
        if isinstance(data_batch, list):
            data_batch = data_batch[0]
        feed = {d.name: v for d, v in zip(self.data_shapes, data_batch.data)}
        # inference with fixed shapes: record the pass once, replay it afterwards (the ~700 launches of a detection
        # graph are host-bound one by one); any change of parameters or shapes drops the recording
        if getattr(self, "_run", None) is None and getattr(self, "_graph_ok", True) and self.exe.device.type == "cuda":
            try:
                self._run = self.exe.capture(**feed)
            except Exception as exc:  # keep working eagerly, say why once
                self._graph_ok = False
                import warnings

                warnings.warn(f"CUDA-graph capture of the bound symbol failed, running eagerly: {exc}")
        with torch.no_grad():
            self.outputs = self._run(**feed) if getattr(self, "_run", None) is not None else self.exe.forward(**feed)

    def get_outputs(self, merge_multi_context=True, begin=0, end=None):
        outs = [nd.NDArray(o) for o in self.outputs]
        return outs if merge_multi_context else [[o] for o in outs]

    def get_output_shapes(self):
        return list(zip(self.symbol.list_outputs(), self._out_shapes))

    def install_monitor(self, mon):
        pass
