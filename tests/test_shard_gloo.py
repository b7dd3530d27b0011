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
