"""PermutationLoss / BCEFocalLoss; mirrors reference adapteacher/modeling/GModule/utils/losses.py:72-103,
400-455 (the only losses ``perm_loss='perm'`` selects, mgm:459-461)."""
import torch
import torch.nn as nn

from ttdg_b200 import ops


class BCEFocalLoss(nn.Module):
    def __init__(self, gamma=2, alpha=0.25, reduction='elementwise_mean'):
        super().__init__()
        if gamma != 2 or alpha != 0.25 or reduction != 'elementwise_mean':
            raise NotImplementedError("the matching head uses BCEFocalLoss() defaults (losses.py:77, :417)")
        self.gamma, self.alpha, self.reduction = gamma, alpha, reduction

    def forward(self, _input, target):
        return ops.focal_bce(_input, target)


class PermutationLoss(nn.Module):
    def __init__(self):
        super().__init__()
        self.loss = BCEFocalLoss()

    def forward(self, pred_dsmat, gt_perm, src_ns=None, tgt_ns=None):
        pred = pred_dsmat.to(dtype=torch.float32)
        if pred.dim() == 3:
            pred = pred.squeeze()
        gt = gt_perm.squeeze() if gt_perm.dim() == 3 else gt_perm
        # losses.py:437-442
        assert torch.all((pred >= 0) * (pred <= 1))
        assert torch.all((gt >= 0) * (gt <= 1))
        return self.loss(pred.contiguous(), gt.to(torch.float32).contiguous())
