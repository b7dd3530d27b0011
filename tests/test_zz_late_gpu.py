"""GPU twins of the operators added after this round's 180 GPU-minutes were spent (round 2, last session): they
could not be run on a B200 by the builder, so the driver's round-end run is their FIRST run on a device.  Their host
logic and, for kernels, the kernel source itself run on the CPU in tests/test_bbox_target_host.py and
tests/test_mask_paste_host.py.  They are marked xfail(strict=False) for exactly that reason and nothing else: an
XPASS in the driver's record is the device confirmation, an XFAIL is a defect of these late additions that must not
mask the 200+ device-verified tests before them (the file sorts last for the same reason)."""
import os

import numpy as np
import pytest
import torch

from simpledet_b200 import ops

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="written after the round's GPU budget was spent; first device run "
                                                     "is the driver's (see the module docstring)")]
GOLD = os.path.join(os.path.dirname(__file__), "golden")


# ---- CustomOp 'bbox_target' --------------------------------------------------------------------------------------------
def test_bbox_target_against_the_reference_operator(cuda):
    from test_bbox_target_host import G, check_case, kwargs_of

    for name in (str(n) for n in G["names"]):
        np.random.seed(int(G[f"{name}_seed"]))                      # the operator draws from the global numpy RNG
        out = ops.OPS["bbox_target"](torch.from_numpy(G[f"{name}_prop"]).to(cuda),
                                     torch.from_numpy(G[f"{name}_gt"]).to(cuda), **kwargs_of(name))
        check_case(name, out)
