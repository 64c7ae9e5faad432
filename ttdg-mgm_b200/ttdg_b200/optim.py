"""Flat-bucket SGD for the test-time-adaptation step (reference: the ``optimizer`` the caller steps at
adapteacher/engine/trainer.py:480-482, built by Detectron2's ``build_optimizer`` - SGD, momentum 0.9, weight
decay 1e-4, lr ``SOLVER.BASE_LR`` = 0.005 from configs/test_segment.yaml:28, no scheduler during TTT).

All adapted parameters live in ONE flat fp32 buffer (parameters become views of it), their gradients in a second
one (``p.grad`` are views, autograd accumulates in place), so a step is: [one NCCL all-reduce of the gradient
bucket when world_size > 1] + one fused kernel (``ttdg_sgd_step``, 20 bytes per parameter of HBM traffic).  The
1 / world_size of the all-reduce average is folded into the kernel."""
import ctypes

import torch

from . import _C
from ._C import check


class FlatSGD:
    """``buckets``: optional list of parameter counts that cut ``params`` (in order) into gradient buckets for the overlapped
    all-reduce: bucket k is reduced on NCCL's stream as soon as ``grad_ready(k)`` is called from the backward pass (tensor
    hooks placed by the detector where a stage's gradients are complete), while the rest of the backward still runs."""

    def __init__(self, params, lr=0.005, momentum=0.9, weight_decay=1e-4, buckets=None, on_step=None):
        self.on_step = on_step        # called after every update (the detector refreshes its tensor-core weight copies)
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatSGD: no trainable parameters")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise _C.TTDGError("FlatSGD needs CUDA parameters (there is no CPU fallback)")
        self.lr, self.momentum, self.weight_decay = float(lr), float(momentum), float(weight_decay)
        sizes = [(p.numel() + 3) // 4 * 4 for p in self.params]          # keep every view 16-byte aligned
        self.numel = sum(sizes)
        self.flat_p = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        off, starts = 0, []
        with torch.no_grad():
            for p, n in zip(self.params, sizes):
                starts.append(off)
                view = self.flat_p[off:off + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view
                p.grad = self.flat_g[off:off + p.numel()].view_as(p)
                off += n
        self.steps = 0
        # gradient buckets (element ranges of the flat buffer)
        self.bucket_ranges = [(0, self.numel)]
        if buckets:
            assert sum(buckets) == len(self.params), (sum(buckets), len(self.params))
            self.bucket_ranges, i = [], 0
            for cnt in buckets:
                lo = starts[i] if i < len(starts) else self.numel
                i += cnt
                self.bucket_ranges.append((lo, starts[i] if i < len(starts) else self.numel))
        self._overlap = None          # (world_size, process_group) while the overlapped all-reduce is enabled
        self._pending, self._launched = [], 0

    def zero_grad(self, set_to_none=False):
        self.flat_g.zero_()
        self._pending, self._launched = [], 0

    # ---- overlapped, bucketed gradient all-reduce
    def enable_overlap(self, world_size, process_group=None):
        """From now on ``grad_ready(k)`` (called by the backward pass) starts the all-reduce of buckets 0..k right away."""
        self._overlap = (world_size, process_group) if world_size > 1 else None

    def grad_ready(self, k):
        if self._overlap is None:
            return
        while self._launched <= min(k, len(self.bucket_ranges) - 1):
            lo, hi = self.bucket_ranges[self._launched]
            if hi > lo:                 # async: NCCL's stream waits for the gradient kernels issued so far, the backward goes on
                self._pending.append(torch.distributed.all_reduce(self.flat_g[lo:hi], group=self._overlap[1], async_op=True))
            self._launched += 1

    def allreduce(self, world_size=1, process_group=None):
        """The path's one exchange: all-reduce (sum) of the flat gradient buffer over NVLink - in one piece, or the
        buckets the backward pass has not started yet followed by a wait for all of them."""
        if world_size <= 1:
            return
        if self._overlap is None:
            torch.distributed.all_reduce(self.flat_g, group=process_group)
            return
        self.grad_ready(len(self.bucket_ranges) - 1)          # every rank issues the same collectives in the same order
        for w in self._pending:
            w.wait()
        self._pending = []

    def step(self, world_size=1, process_group=None, reduce=True):
        """All-reduce (sum) the gradient bucket when world_size > 1 (unless the caller already did), then one fused update."""
        if reduce:
            self.allreduce(world_size, process_group)
        s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(_C.lib().ttdg_sgd_step(ctypes.c_void_p(self.flat_p.data_ptr()), ctypes.c_void_p(self.flat_g.data_ptr()),
                                     ctypes.c_void_p(self.flat_m.data_ptr()), self.numel, self.lr, self.momentum,
                                     self.weight_decay, 1.0 / world_size, int(self.steps == 0), s), "sgd_step")
        self.steps += 1
        from . import detector
        detector.PARAM_EPOCH[0] += 1          # K-major tensor-core copies of the weights are stale now
        if self.on_step is not None:
            self.on_step()
