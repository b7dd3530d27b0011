"""Runs the reference's OWN detection_infer_speed.py, unchanged, on the façade (simpledet_b200.facade): the script's
`import mxnet`, the config, symbol/builder.py, models/FPN/builder.py and core/detection_module.py all come from
/root/reference; `mxnet` / `mxnext` are the stand-ins, the graph runs on torch + the C ABI.

  python benchmarks/infer_speed_facade.py [--config config/faster_r50v1_fpn_1x.py] [--shape 800 1333] [--count 100]

Prints the script's own output: milliseconds per iteration (all-zero weights, as the reference harness times it)."""
import argparse
import os
import runpy
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simpledet_b200 import facade  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--config", default="config/faster_r50v1_fpn_1x.py")
    ap.add_argument("--shape", nargs=2, default=["800", "1333"])
    ap.add_argument("--count", default="100")
    ap.add_argument("--gpu", default="0")
    a = ap.parse_args()
    facade.install(a.reference)
    os.chdir(a.reference)  # the script resolves the config path relative to the repository root
    sys.argv = ["detection_infer_speed.py", "--config", a.config, "--shape", *a.shape, "--count", a.count, "--gpu", a.gpu]
    runpy.run_path(os.path.join(a.reference, "detection_infer_speed.py"), run_name="__main__")


if __name__ == "__main__":
    main()
