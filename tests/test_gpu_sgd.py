"""Fused flat-bucket SGD against torch.optim.SGD (the optimizer the reference's caller steps, trainer.py:480-482)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_flat_sgd_matches_torch_sgd():
    from ttdg_b200.optim import FlatSGD
    gen = torch.Generator().manual_seed(0)
    shapes = [(512, 512), (512,), (1, 512), (1,), (256, 256), (7, 3)]
    ref = [torch.randn(*s, generator=gen).requires_grad_(True) for s in shapes]
    ours = [p.detach().clone().cuda().requires_grad_(True) for p in ref]
    opt_ref = torch.optim.SGD(ref, lr=0.005, momentum=0.9, weight_decay=1e-4)
    opt = FlatSGD(ours, lr=0.005, momentum=0.9, weight_decay=1e-4)
    for step in range(4):
        grads = [torch.randn(*s, generator=gen) for s in shapes]
        opt_ref.zero_grad()
        opt.zero_grad()
        for p, q, g in zip(ref, ours, grads):
            p.grad = g.clone()
            (q * g.cuda()).sum().backward()          # autograd accumulates into the flat gradient views
        opt_ref.step()
        opt.step()
        for p, q in zip(ref, ours):
            np.testing.assert_allclose(q.detach().cpu().numpy(), p.detach().numpy(), rtol=2e-6, atol=1e-7)
    assert all(q.data_ptr() >= opt.flat_p.data_ptr() for q in ours)
