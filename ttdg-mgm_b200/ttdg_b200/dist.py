"""Multi-GPU plumbing of the hot path (one process per GPU, torch.distributed).

The path shards over IMAGES: every rank adapts on its own contiguous slice of the test set (the reference's
``InferenceSampler``, adapteacher/data/build.py:139-146) and forms its own matching problem of ``TEST.BATCH`` graphs;
the only exchange is one all-reduce (sum) of the flat gradient bucket per step, averaged inside the fused SGD kernel
(``optim.FlatSGD.step(world_size)``).  The reference's eval-only path has no gradient sync at all; the
average mirrors what its DDP training does (engine/trainer.py:210-213)."""
import torch


def shard_range(n_items, rank, world_size):
    """Contiguous shard [begin, end) of ``n_items`` for ``rank`` - d2 InferenceSampler: shard sizes differ by at most one,
    the first ``n_items % world_size`` ranks get the extra item."""
    base, extra = divmod(n_items, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def batches(begin, end, batch_size):
    """``BatchSampler(..., batch_size, drop_last=False)`` over one shard (data/build.py:141-146)."""
    return [(i, min(i + batch_size, end)) for i in range(begin, end, batch_size)]


def allreduce_mean_(flat, world_size, group=None):
    """Reference semantics of the gradient exchange on any backend (used by the CPU tests; on the GPU the 1 / world_size
    is folded into ttdg_sgd_step instead of a separate pass)."""
    if world_size > 1:
        torch.distributed.all_reduce(flat, group=group)
        flat.mul_(1.0 / world_size)
    return flat
