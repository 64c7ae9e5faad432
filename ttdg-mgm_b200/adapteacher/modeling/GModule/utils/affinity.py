"""Learned node affinity; mirrors reference adapteacher/modeling/GModule/utils/affinity.py:9-57 (same
parameters / state-dict keys), evaluated in separable form without the N1 x N2 x 512 tensor."""
import torch
import torch.nn as nn

from ttdg_b200 import ops


class Affinity(nn.Module):
    def __init__(self, d=256):
        super().__init__()
        self.d = d
        self.fc_M = nn.Sequential(nn.Linear(512, 512), nn.ReLU(), nn.Linear(512, 1))
        self.project_sr = nn.Linear(256, 256, bias=False)
        self.project_tg = nn.Linear(256, 256, bias=False)
        self.reset_parameters()

    def reset_parameters(self):
        for m in self.fc_M:
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, std=0.01)
                nn.init.constant_(m.bias, 0)
        nn.init.normal_(self.project_sr.weight, std=0.01)
        nn.init.normal_(self.project_tg.weight, std=0.01)

    def params(self):
        return (self.project_sr.weight, self.project_tg.weight, self.fc_M[0].weight, self.fc_M[0].bias,
                self.fc_M[2].weight, self.fc_M[2].bias)

    def forward_pairs(self, X, sizes, pairs):
        """All listed (src, tgt) graph pairs of the stacked node matrix X in one pass (flat pair blocks)."""
        return ops.affinity_pairs(X, *self.params(), sizes, pairs)

    def forward(self, X, Y):
        n1, n2 = X.shape[0], Y.shape[0]
        out = ops.affinity_pairs(torch.cat([X, Y], 0), *self.params(), [n1, n2], [(0, 1)])
        return out.reshape(n1, n2).squeeze()            # the reference ends in .squeeze() (affinity.py:55)
