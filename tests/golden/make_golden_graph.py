#!/usr/bin/env python
"""Writes tests/golden/<config>_test_symbol.json: the inference graphs that the REFERENCE'S OWN config/*.py +
symbol/builder.py + models/*/builder.py build when they run, unchanged, on the `mxnet` / `mxnext` façade
(simpledet_b200.facade).  The GPU box has no reference checkout; the graphs travel as these fixtures.  One interpreter
per config: the reference caches the RPN sub-graph in a class attribute (symbol/builder.py FasterRcnn._rpn_output).
Run:  python tests/golden/make_golden_graph.py     (needs /root/reference)"""
import importlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
# BASELINE.json configs 2-5: Faster R-CNN FPN, RetinaNet, Mask R-CNN, DCNv1 Faster R-CNN C4
CONFIGS = ["faster_r50v1_fpn_1x", "retina_r50v1_fpn_1x", "mask_r50v1_fpn_1x", "dcn.faster_dcn_r50v1bc4_c5_512roi_1x",
                 "crowdhuman.faster_r50v1b_fpn_1x", "cascade_r50v1_fpn_1x", "tridentnet_r50v1c4_c5_1x", "ms_r50v1_fpn_1x"]
# the TRAIN graphs the façade's Trainer runs (simpledet_b200/facade/train.py): <config>_train_symbol.json
TRAIN_CONFIGS = ["faster_r50v1_fpn_1x", "retina_r50v1_fpn_1x", "mask_r50v1_fpn_1x", "dcn.faster_dcn_r50v1bc4_c5_512roi_1x",
                 "crowdhuman.faster_r50v1b_fpn_1x", "cascade_r50v1_fpn_1x", "tridentnet_r50v1c4_c5_1x", "ms_r50v1_fpn_1x"]


def one(name, train=False):
    sys.path.insert(0, ROOT)
    from simpledet_b200 import facade

    facade.install("/root/reference")
    cfg = importlib.import_module("config." + name)
    sym = cfg.get_config(is_train=True)[6].train_symbol if train else cfg.get_config(is_train=False)[6].test_symbol
    stem = name.replace(".", "_") if name.startswith("crowdhuman.") else name.split(".")[-1]
    path = os.path.join(HERE, stem + ("_train_symbol.json" if train else "_test_symbol.json"))
    open(path, "w").write(sym.tojson())
    print("wrote", path, os.path.getsize(path), "bytes;", len(sym._topo()), "nodes; outputs", sym.list_outputs())


if __name__ == "__main__":
    if len(sys.argv) > 1:
        one(sys.argv[1], train=len(sys.argv) > 2 and sys.argv[2] == "train")
    else:
        for c in CONFIGS:
            subprocess.run([sys.executable, os.path.abspath(__file__), c], check=True)
        for c in TRAIN_CONFIGS:
            subprocess.run([sys.executable, os.path.abspath(__file__), c, "train"], check=True)
