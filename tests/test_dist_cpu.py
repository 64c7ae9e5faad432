"""Host-side multi-GPU logic on CPU with the gloo backend, world_size 2: image sharding (d2 InferenceSampler
semantics) and the gradient-bucket all-reduce + SGD update giving identical replicas."""
import os
import socket

import torch
import torch.multiprocessing as mp

from ttdg_b200 import dist as tdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [tdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    assert tdist.batches(0, 8, 5) == [(0, 5), (5, 8)]          # TEST.BATCH 5 on a shard of 8: problems of 5 + 3 graphs


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    p = torch.randn(1000, generator=g)                              # identical replicas
    m = torch.zeros(1000)
    lr, mu, wd = 0.005, 0.9, 1e-4
    for step in range(3):
        grad = torch.randn(1000, generator=torch.Generator().manual_seed(100 * step + rank))      # per-rank gradient
        flat = tdist.allreduce_mean_(grad.clone(), world)
        d = flat + wd * p
        m = d if step == 0 else mu * m + d
        p = p - lr * m
    out[rank] = p
    torch.distributed.destroy_process_group()


def test_gradient_allreduce_keeps_replicas_identical():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        p0, p1 = out[0], out[1]
    assert torch.equal(p0, p1)
    # and equal to one process seeing the averaged gradients
    g = torch.Generator().manual_seed(0)
    p = torch.randn(1000, generator=g)
    m = torch.zeros(1000)
    for step in range(3):
        gs = [torch.randn(1000, generator=torch.Generator().manual_seed(100 * step + r)) for r in range(world)]
        d = (gs[0] + gs[1]) * 0.5 + 1e-4 * p
        m = d if step == 0 else 0.9 * m + d
        p = p - 0.005 * m
    torch.testing.assert_close(p0, p, rtol=1e-6, atol=1e-7)
