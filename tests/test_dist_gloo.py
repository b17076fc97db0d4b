"""world_size-2 gloo test of the data-parallel host logic (flat gradient buffer, single all-reduce,
batch sharding).  The CUDA forward/backward itself is covered by the -m gpu tests."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pose2mesh_release_b200.dist import DataParallelStep, shard_batch

    torch.manual_seed(100 + rank)  # different init per rank: broadcast must fix it
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.BatchNorm1d(5), torch.nn.Linear(5, 2))
    step = DataParallelStep(net)
    w0 = step.flat.data.clone()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(8, 6, generator=g)
    a, b = shard_batch(8, rank, world)
    step.zero_grad()
    net(x[a:b]).sum().backward()
    local = step.flat.grad.clone()
    step.reduce_gradients()
    q.put((rank, w0, local, step.flat.grad.clone(), (a, b)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_flat_allreduce_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=100) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(30)
    (_, w0a, la, ga, sa), (_, w0b, lb, gb, sb) = res
    assert torch.equal(w0a, w0b), "parameters must be identical after the initial broadcast"
    assert sa == (0, 4) and sb == (4, 8)
    assert torch.allclose(ga, (la + lb) / 2) and torch.equal(ga, gb)


def test_shard_batch_covers_everything():
    from pose2mesh_release_b200.dist import shard_batch

    for n in (1, 7, 256, 2048):
        for w in (1, 2, 3, 8):
            spans = [shard_batch(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
