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


# ---- the sharded adaptation pass of BaselineTrainer.test: ragged shards and ranks without a loss must not dead-lock
class _StubTTT(torch.nn.Module):
    """Meta-arch stand-in on CPU: loss = w * sum(image ids); a batch with a single image yields no loss (mgm:489-490)."""

    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.ones(1))

    def forward(self, inputs, branch="supervised"):
        if self.training:
            if len(inputs) < 2:
                return None, [], [], []
            return self.w * float(sum(d["image_id"] for d in inputs)), [], [], []
        return [{"image_id": d["image_id"]} for d in inputs]


class _StubOpt:
    """FlatSGD's interface (zero_grad / step(world) with the all-reduce inside) on a CPU parameter."""

    def __init__(self, p, lr):
        self.p, self.lr, self.flat_g, self.steps = p, lr, None, 0

    def zero_grad(self):
        self.p.grad = torch.zeros_like(self.p)

    def step(self, world):
        tdist.allreduce_mean_(self.p.grad, world)
        with torch.no_grad():
            self.p -= self.lr * self.p.grad
        self.steps += 1


class _Eval:
    def __init__(self):
        self.seen = []

    def reset(self):
        self.seen = []

    def process(self, inputs, outputs):
        self.seen += [o["image_id"] for o in outputs]

    def evaluate(self):
        return {"Dice Coefficient": float(len(self.seen))}


def _ttt_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    from adapteacher.config import add_ateacher_config
    from adapteacher.data import build_detection_test_loader
    from adapteacher.engine.trainer import BaselineTrainer
    cfg = add_ateacher_config()
    cfg.TEST.BATCH = 2
    cfg.DATASETS.TEST = ("synthetic_polyp_7_32",)
    # 7 images over 2 ranks: rank 0 gets 4 (batches 2 + 2), rank 1 gets 3 (batches 2 + 1: the second yields NO loss)
    loader = build_detection_test_loader(cfg, "synthetic_polyp_7_32")
    model = _StubTTT()
    opt = _StubOpt(model.w, 0.01)
    ev = _Eval()
    BaselineTrainer.test(cfg, model, opt, evaluators=[ev], data_loaders={"synthetic_polyp_7_32": loader}, world_size=world)
    out[rank] = (float(model.w), opt.steps, sorted(ev.seen))
    torch.distributed.destroy_process_group()


def test_sharded_ttt_pass_is_collective_safe():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_ttt_worker, args=(world, port, out), nprocs=world, join=True)
        r0, r1 = out[0], out[1]
    assert r0[0] == r1[0] and r0[1] == r1[1] == 2             # identical replicas, both ranks took both steps
    # step 1: grads (0 + 1) and (4 + 5) averaged; step 2: rank 0 has (2 + 3), rank 1 has no loss -> contributes zero
    expect = 1.0 - 0.01 * (1 + 9) / 2 - 0.01 * (5 + 0) / 2
    assert abs(r0[0] - expect) < 1e-6
    assert r0[2] == [0, 1, 2, 3] and r1[2] == [4, 5, 6]        # evaluation pass: every image once, on its own rank
