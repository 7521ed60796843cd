"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: env sharding and the single
flat gradient all-reduce the trainer issues per iteration."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    from warp_drive_b200.training.utils.distributed import (
        flat_allreduce_mean_, init_process_group, max_over_ranks, shard_of)

    r, w = init_process_group("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    x = torch.randn(16, 7, generator=torch.Generator().manual_seed(100 + rank))
    model(x).square().mean().backward()
    local = [p.grad.clone() for p in model.parameters()]
    buf = flat_allreduce_mean_([p.grad for p in model.parameters()], world)
    buf2 = flat_allreduce_mean_([p.grad for p in model.parameters()], world, buf)
    assert buf2 is buf                                    # the bucket is reused
    gathered = [None] * world
    dist.all_gather_object(gathered, [g.tolist() for g in local])
    mean = [sum(torch.tensor(gathered[k][i]) for k in range(world)) / world
            for i in range(len(local))]
    # two all-reduces of already-averaged grads leave the mean unchanged
    ok = all(torch.allclose(p.grad, m, atol=1e-6) for p, m in zip(model.parameters(), mean))
    slow = max_over_ranks(1.0 + rank)
    out.put((rank, ok, slow, shard_of(16001, rank, world)))
    dist.destroy_process_group()


def test_flat_gradient_allreduce_and_sharding_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res)
    assert all(r[2] == 2.0 for r in res)                  # max over ranks
    (s0, n0), (s1, n1) = res[0][3], res[1][3]
    assert s0 == 0 and s1 == n0 and n0 + n1 == 16001 and abs(n0 - n1) <= 1


def test_shard_of_covers_everything():
    from warp_drive_b200.training.utils.distributed import shard_of

    for total, world in [(16000, 8), (7, 3), (5, 8)]:
        spans = [shard_of(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and sum(n for _, n in spans) == total
        for (s, n), (s2, _) in zip(spans[:-1], spans[1:]):
            assert s + n == s2
