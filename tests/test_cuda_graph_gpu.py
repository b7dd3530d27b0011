"""The inference hot path of bench.py captured in ONE CUDA graph and replayed: every operator is
stream-ordered, allocation-free on the device (caller-owned workspaces from torch's graph-safe pool),
and never synchronises with the host — the property the reference's Proposal op lacks
(proposal_v3.cu:359-380 blocks on a D2H of the NMS mask)."""
import numpy as np
import pytest
import torch

import bench
from simpledet_b200 import _lib, ops

pytestmark = pytest.mark.gpu


def test_hot_path_replays_from_a_cuda_graph():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(5)
    sets = [{k: torch.from_numpy(v).to(dev) for k, v in bench.Infer.make_inputs(rng, 1).items()} for _ in range(2)]
    static = {k: v.clone() for k, v in sets[0].items()}
    def step(d):
        out = bench.Infer.step(ops, d)
        return [out["rois"], *out["result"]]

    eager = [[t.clone() for t in step(s)] for s in sets]   # also warms every lazy init
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        step(static)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    n0 = _lib.launch_count()
    with torch.cuda.graph(g):
        outs = step(static)
    launches = _lib.launch_count() - n0
    assert launches >= 10
    for i in (1, 0, 1):
        for k, v in sets[i].items():
            static[k].copy_(v)
        g.replay()
        torch.cuda.synchronize()
        for got, want in zip(outs, eager[i]):
            assert torch.equal(got, want)
