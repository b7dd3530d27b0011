"""world_size-2 gloo test of the N>1 path's host logic (no GPU): image sharding covers the global
batch exactly once and the timing reduction is the max over ranks."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from simpledet_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.image_shard(total, rank, world)
    owned = torch.zeros(total, dtype=torch.int64)
    owned[list(mine)] = 1
    dist.all_reduce(owned)  # test-only collective: every image owned exactly once
    ms = shard.max_over_ranks(10.0 + 5.0 * rank)
    thr = shard.whole_job_throughput(len(mine), 10.0 + 5.0 * rank)
    q.put((rank, list(mine), owned.tolist(), ms, thr))
    dist.destroy_process_group()


def test_shard_two_ranks_gloo():
    world, total = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 1, 2, 3] and res[1][1] == [4, 5, 6]
    assert res[0][2] == [1] * total
    assert res[0][3] == res[1][3] == 15.0            # max over ranks
    assert abs(res[0][4] - 2 * 4 / 0.015) < 1e-6     # world * units_per_rank / max time


def test_image_shard_properties():
    for total in (1, 2, 5, 16, 17):
        for world in (1, 2, 3, 8):
            parts = [list(shard.image_shard(total, r, world)) for r in range(world)]
            assert sum(parts, []) == list(range(total))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1
    assert shard.max_over_ranks(3.5) == 3.5  # no process group: identity


# ---- the gradient all-reduce of the façade Trainer (simpledet_b200/facade/train.py): one flat bucket, two ranks ------
def _train_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from simpledet_b200.facade import mxnext_impl as X
    from simpledet_b200.facade import symbol as S
    from simpledet_b200.facade import train as T

    data = S.Variable("data")
    h = X.relu(X.fc(data, "fc1", 6), name="r")
    out = X.loss(X.fc(h, "fc2", 3), grad_scale=0.5, name="l")
    tr = T.Trainer(out, dict(data=(4, 5)), device="cpu", rng_std=0.3)          # same seed: same initial weights
    x = torch.arange(20, dtype=torch.float32).reshape(4, 5) * (0.1 + rank)     # each rank its own images
    tr.forward_backward(data=x)
    local = {k: v.clone() for k, v in tr.grads().items()}
    tr.allreduce_grads()
    summed = {k: v.clone() for k, v in tr.grads().items()}
    tr.update(lr=0.05, momentum=0.9, wd=1e-4, rescale_grad=1.0 / world)        # detection_train.py:266
    q.put((rank, {k: v.tolist() for k, v in local.items()}, {k: v.tolist() for k, v in summed.items()},
           {k: tr.ex.params[k].detach().tolist() for k in tr.trainable}))
    dist.destroy_process_group()


def test_trainer_flat_allreduce_two_ranks_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_train_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=180) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, l0, s0, w0), (_, l1, s1, w1) = res
    assert set(l0) == {"fc1_weight", "fc1_bias", "fc2_weight", "fc2_bias"}
    for k in l0:
        want = torch.tensor(l0[k]) + torch.tensor(l1[k])
        assert torch.allclose(torch.tensor(s0[k]), want, rtol=1e-6, atol=1e-7) and s0[k] == s1[k], k
        assert w0[k] == w1[k], k                                               # replicas stay identical
    assert l0["fc1_weight"] != l1["fc1_weight"]
