#!/usr/bin/env python
"""Writes tests/golden/faster_r50v1_fpn_1x_test_symbol.json: the inference graph that the REFERENCE'S OWN
config/faster_r50v1_fpn_1x.py + symbol/builder.py + models/FPN/builder.py build when they run, unchanged, on the
`mxnet` / `mxnext` façade (simpledet_b200.facade).  The GPU box has no reference checkout; the graph travels as this
fixture.  Run:  python tests/golden/make_golden_graph.py     (needs /root/reference)"""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from simpledet_b200 import facade  # noqa: E402

facade.install("/root/reference")
cfg = importlib.import_module("config.faster_r50v1_fpn_1x")
sym = cfg.get_config(is_train=False)[6].test_symbol
path = os.path.join(HERE, "faster_r50v1_fpn_1x_test_symbol.json")
open(path, "w").write(sym.tojson())
print("wrote", path, os.path.getsize(path), "bytes;", len(sym._topo()), "nodes; outputs", sym.list_outputs())
