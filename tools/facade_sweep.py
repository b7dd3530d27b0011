#!/usr/bin/env python
"""Builds the inference (and, with --train, the training) graph of EVERY config under /root/reference/config on the
`mx` / `mxnext` stand-ins, one interpreter per config (the reference caches sub-graphs in class attributes), and prints
one line per config plus a summary.  python tools/facade_sweep.py [--train]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROBE = r"""
import importlib, sys
sys.path.insert(0, {root!r})
from simpledet_b200 import facade
facade.install("/root/reference")
cfg = importlib.import_module("config." + {name!r})
out = cfg.get_config(is_train={train})[6]
sym = out.train_symbol if {train} else out.test_symbol
if sym is None:
    print("NONE the config defines no such symbol")
else:
    ops = sorted({{n.op for n in sym._topo() if n.op}})
    det = [o for o in ops if o.startswith("_contrib_") or o in ("Custom", "ProposalTarget", "ProposalTarget_v2", "ProposalMaskTarget", "ROIPooling_v1")]
    print("OK", len(sym._topo()), "nodes;", " ".join(det))
"""


def main():
    train = "--train" in sys.argv
    cdir = "/root/reference/config"
    names = []
    for d, _, files in os.walk(cdir):
        for f in sorted(files):
            if f.endswith(".py") and f != "__init__.py":
                names.append(os.path.relpath(os.path.join(d, f), cdir)[:-3].replace(os.sep, "."))
    ok = 0
    for name in sorted(names):
        r = subprocess.run([sys.executable, "-c", PROBE.format(root=ROOT, name=name, train=train)], capture_output=True,
                           text=True, timeout=300)
        line = (r.stdout.strip().splitlines() or [""])[-1]
        if r.returncode == 0 and line.startswith("OK"):
            ok += 1
        else:
            line = "FAIL " + (r.stderr.strip().splitlines() or [line or "?"])[-1][:160]
        print(f"{name:60s} {line}")
    print(f"{ok} of {len(names)} configs build their {'training' if train else 'inference'} graph")


if __name__ == "__main__":
    main()
