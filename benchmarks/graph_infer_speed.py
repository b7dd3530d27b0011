"""detection_infer_speed.py's measurement (1 warm-up + `count` timed iterations with a device sync per iteration,
ms / iteration) for the faster_r50v1_fpn_1x inference graph the reference's builders produce (committed fixture
tests/golden/faster_r50v1_fpn_1x_test_symbol.json), run by the façade executor: cuDNN / cuBLAS for backbone and heads,
the C ABI for every detection operator.  `--weights zero` is the reference harness's setting, `random` the realistic
variant of SURVEY §8d.

  python benchmarks/graph_infer_speed.py [--count 100] [--weights zero|random] [--tf32 1] [--graph 1]
         [--config faster_r50v1_fpn_1x|retina_r50v1_fpn_1x|mask_r50v1_fpn_1x|faster_dcn_r50v1bc4_c5_512roi_1x]"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simpledet_b200 import _lib, facade  # noqa: E402
from simpledet_b200.facade import symbol as S  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--count", type=int, default=100)
    ap.add_argument("--weights", default="zero")
    ap.add_argument("--tf32", type=int, default=1)
    ap.add_argument("--channels-last", type=int, default=1)
    ap.add_argument("--graph", type=int, default=0, help="1: replay the forward pass from one CUDA graph")
    ap.add_argument("--config", default="faster_r50v1_fpn_1x", help="which committed graph fixture (tests/golden/*_test_symbol.json)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = bool(a.tf32)
    torch.backends.cuda.matmul.allow_tf32 = bool(a.tf32)
    torch.backends.cudnn.benchmark = True
    sym = S.fromjson(open(os.path.join(ROOT, "tests", "golden", a.config + "_test_symbol.json")).read())
    shapes = dict(data=(1, 3, 800, 1333), im_info=(1, 3), im_id=(1,), rec_id=(1,))
    ex = facade.Executor(sym, dev, channels_last=bool(a.channels_last)).init_params(
        shapes, rng_std=None if a.weights == "zero" else 0.02)
    feed = dict(data=torch.ones(shapes["data"], device=dev), im_info=torch.tensor([[400.0, 666.5, 2.0]], device=dev),
                im_id=torch.ones(1, device=dev), rec_id=torch.ones(1, device=dev))
    step = ex.capture(**feed) if a.graph else ex.forward
    with torch.no_grad():
        for _ in range(3):
            step(**feed)
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        tic = time.time()
        for _ in range(a.count):
            step(**feed)
            torch.cuda.synchronize()   # output.wait_to_read() per iteration, as the reference script does
        toc = time.time()
    ms = (toc - tic) / a.count * 1000
    npar = sum(v.numel() for v in ex.params.values())
    print(json.dumps({"graph": "%s test_symbol (%d nodes, %.1f M parameters)" % (a.config, len(sym._topo()), npar / 1e6),
                      "weights": a.weights,
                      "ms_per_iter": round(ms, 3), "images_per_s": round(1000 / ms, 2), "count": a.count,
                      "tf32_conv": bool(a.tf32), "channels_last": bool(a.channels_last), "cuda_graph": bool(a.graph),
                      "sdet_launches_per_iter": (_lib.launch_count() - n0) // a.count}))


if __name__ == "__main__":
    main()
