"""Sinkhorn layer; mirrors reference adapteacher/modeling/GModule/utils/sinkhorn.py:7-87 (``Sinkhorn`` ->
``pygmtools.sinkhorn(backend='pytorch')``), computed by the sm_100a kernels behind ``ttdg_b200.ops.sinkhorn``."""
import torch.nn as nn

from ttdg_b200 import ops


class Sinkhorn(nn.Module):
    def __init__(self, max_iter=10, tau=1., epsilon=1e-4, log_forward=True, batched_operation=False):
        super().__init__()
        self.max_iter = max_iter
        self.tau = tau
        self.epsilon = epsilon                      # ignored by the reference's log path too (sinkhorn.py:85-87)
        self.log_forward = log_forward
        if not log_forward:
            raise NotImplementedError("only the log-space forward (the reference default) is provided")
        self.batched_operation = batched_operation  # per-item and batched paths are one kernel here

    def forward(self, s, nrows=None, ncols=None, dummy_row=False):
        return ops.sinkhorn(s, nrows, ncols, dummy_row=dummy_row, max_iter=self.max_iter, tau=self.tau)
