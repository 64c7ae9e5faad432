"""GPU parity of the GA-GM solver and of MGM3_unsup end to end."""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import mgm_port  # noqa: E402  (checker only)
from ttdg_b200 import ops, synth  # noqa: E402

T = torch.from_numpy
MGM_FILES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "mgm_*.npz")))
IDS = [os.path.basename(p)[4:-4] for p in MGM_FILES]
TRACE_FILES = [p for p in MGM_FILES if "trace_iter" in np.load(p).files]


@pytest.mark.parametrize("path", TRACE_FILES, ids=[os.path.basename(p)[4:-4] for p in TRACE_FILES])
def test_gagm_single_steps_match_reference_trace(path):
    """Teacher-forced on the reference's own fp32 trajectory: (U_t, projector, tau) -> U_{t+1}."""
    g = np.load(path)
    ms = [int(x) for x in g["sizes"]]
    A, W = T(g["A"]).cuda(), T(g["Wds"]).cuda()
    for k in range(len(g["trace_iter"])):
        proj, tau = int(g["trace_proj"][k]), float(g["trace_tau"][k])
        Uin, Uout = T(g[f"trace_Uin_{k}"]), g[f"trace_Uout_{k}"]
        U1 = ops.gagm_solve(A, W, Uin.cuda(), ms, init_tau=tau if proj == 0 else 1.0, mode=1, step_projector=proj).cpu().numpy()
        ref64, V = mgm_port.gagm_step(T(g["A"]).double(), T(g["Wds"]).double(), Uin.double(), ms, 32,
                                      "hungarian" if proj else "sinkhorn", tau, return_V=True)
        if proj == 1:
            assert np.array_equal(U1, ref64.float().numpy()), (k, "vs float64 oracle")
            if bool(((Uin == 0) | (Uin == 1)).all()):
                assert np.array_equal(U1, Uout), (k, "vs reference fp32 trace")
            else:       # near-ties of V that fp32 noise breaks arbitrarily: both optimal up to that noise
                o1, o2 = float((V.numpy() * U1).sum()), float((V.numpy() * Uout).sum())
                assert abs(o1 - o2) <= 1e-5 * abs(o2) and U1.sum() == Uout.sum()
        else:
            np.testing.assert_allclose(U1, ref64.float().numpy(), atol=1e-6)
            np.testing.assert_allclose(U1, Uout, atol=5e-4 if tau < 0.02 else 5e-5)


@pytest.mark.parametrize("path", MGM_FILES, ids=IDS)
def test_gagm_full_solve_bit_exact_vs_float64_oracle(path):
    g = np.load(path)
    if "A" not in g.files:
        pytest.skip("inputs not stored for the large case (covered end to end below)")
    ms = [int(x) for x in g["sizes"]]
    U, info = ops.gagm_solve(T(g["A"]).cuda(), T(g["Wds"]).cuda(), T(g["U0"]).cuda(), ms, return_info=True)
    U = U.cpu().numpy()
    trace = []
    ref = mgm_port.gagm(T(g["A"]), T(g["Wds"]), T(g["U0"]), ms, 32, precise=True, trace=trace).numpy()
    assert np.array_equal(U, ref)
    info = info.cpu().tolist()
    assert info[0] == len(trace)
    assert info[1] == sum(1 for t in trace if t[0] == "sinkhorn") and info[2] == info[0] - info[1]
    if len(ms) == 2:        # the reference's own fp32 result is reproducible only here (mgm:358-359 pins graph 0)
        assert np.array_equal(U.astype(np.uint8), g["U"])
        assert info[0] == int(g["gagm_iters"])


def _module(variant):
    from adapteacher.modeling.GModule.multi_graph_matching import MGM3_unsup
    sd = synth.mgm_unsup_state(0)
    if variant == "pert":
        sd = synth.perturb_affinity_state(sd, 0)
    m = MGM3_unsup(2, 32).cuda()
    m.load_state_dict(sd, strict=True)
    m.train()
    return m, sd


@pytest.mark.parametrize("path", MGM_FILES, ids=IDS)
def test_mgm3_unsup_teacher_forced_loss_and_grads_vs_reference(path):
    """Everything differentiable (affinity -> Sinkhorn -> focal loss and its backward), with the matching
    result U forced to the reference's, against the reference's own loss and gradients."""
    g = np.load(path)
    sizes, seed, variant = tuple(int(x) for x in g["sizes"]), int(g["seed"]), str(g["variant"])
    m, sd = _module(variant)
    nodes, labels, masks = synth.mgm_inputs(sizes, seed)
    nodes = [n.cuda().requires_grad_(True) for n in nodes]
    m.debug_keep_masks = [k.cuda() for k in masks]
    m.debug_U_override = T(g["U"].astype(np.float32)).cuda()
    loss = m(nodes, [l.cuda() for l in labels], synth.universe(0).cuda())
    loss.backward()
    m.check_flags()
    np.testing.assert_allclose(loss.item(), float(g["loss"]), rtol=2e-5)
    aux = m.last_aux
    if "Wds" in g.files:
        np.testing.assert_allclose(aux["Wds"].cpu().numpy(), g["Wds"], atol=3e-6)
        np.testing.assert_allclose(aux["A"].cpu().numpy(), g["A"], atol=2e-7)
        np.testing.assert_allclose(aux["U0"].cpu().numpy(), g["U0"], atol=2e-5, rtol=1e-5)
    for i, n in enumerate(nodes):
        np.testing.assert_allclose(n.grad.cpu().numpy(), g[f"grad_nodes_{i}"], atol=2e-7, rtol=2e-3, err_msg=f"nodes {i}")
    gw = m.node_affinity.fc_M[2].weight.grad.cpu().numpy()
    np.testing.assert_allclose(gw, g["grad_aff_fc_M.2.weight"], atol=2e-7, rtol=2e-3)
    for k, p in m.node_affinity.named_parameters():
        if "grad_aff_" + k + "_sum" in g.files:
            gr = p.grad.cpu()
            sub = gr[::16, ::16] if gr.dim() == 2 else gr[::16]
            np.testing.assert_allclose(sub.numpy(), g["grad_aff_" + k + "_sub16"], atol=2e-7, rtol=2e-3, err_msg=k)
            np.testing.assert_allclose(float((gr.double() ** 2).sum()), float(g["grad_aff_" + k + "_sumsq"]), rtol=5e-3)
    # attention and universe receive no gradient at test time
    assert all(p.grad is None for p in m.intra_domain_graph.parameters())


@pytest.mark.parametrize("path", MGM_FILES, ids=IDS)
def test_mgm3_unsup_end_to_end(path):
    """Free-running: U must equal the float64 oracle solver run on the CUDA path's own (A, Wds, U0); the loss must
    equal the oracle loss for that U.  G == 2 cases also reproduce the reference's fp32 result bit for bit."""
    g = np.load(path)
    sizes, seed, variant = tuple(int(x) for x in g["sizes"]), int(g["seed"]), str(g["variant"])
    m, sd = _module(variant)
    nodes, labels, masks = synth.mgm_inputs(sizes, seed)
    m.debug_keep_masks = [k.cuda() for k in masks]
    loss = m([n.cuda().requires_grad_(True) for n in nodes], [l.cuda() for l in labels], synth.universe(0).cuda())
    aux = m.last_aux
    U = aux["U"].cpu().numpy()
    ref = mgm_port.gagm(aux["A"].cpu(), aux["Wds"].cpu(), aux["U0"].cpu(), list(sizes), 32, precise=True).numpy()
    assert np.array_equal(U, ref)
    assert set(np.unique(U)) <= {0.0, 1.0}
    for gi, n in enumerate(sizes):
        o = sum(sizes[:gi])
        assert U[o:o + n].sum() == min(n, 32) and (U[o:o + n].sum(0) <= 1).all() and (U[o:o + n].sum(1) <= 1).all()
    # loss for this U from the oracle's formula
    Wds = aux["Wds"].cpu()
    offs = np.concatenate([[0], np.cumsum(sizes)])
    tot, cnt = 0.0, 0
    for i1 in range(len(sizes)):
        for i2 in range(i1 + 1, len(sizes)):
            s = Wds[offs[i2]:offs[i2 + 1], offs[i1]:offs[i1 + 1]].t()
            y = T(U[offs[i1]:offs[i1 + 1]] @ U[offs[i2]:offs[i2 + 1]].T)
            tot += float(mgm_port.focal_bce(s, y))
            cnt += 1
    np.testing.assert_allclose(loss.item(), tot / cnt, rtol=2e-5)
    if len(sizes) == 2:
        assert np.array_equal(U.astype(np.uint8), g["U"])
        np.testing.assert_allclose(loss.item(), float(g["loss"]), rtol=2e-5)


def test_mgm3_unsup_single_graph_returns_none():
    m, _ = _module("init")
    assert m([torch.randn(5, 256, device="cuda")], [torch.ones(5, device="cuda")], synth.universe(0).cuda()) is None
    assert m(None, None, synth.universe(0).cuda()) is None


def test_gagm_many_graphs_and_extreme_sizes():
    """More graphs than the cluster has CTAs (G = 11 > 8), sizes 1 and 96."""
    gen = torch.Generator().manual_seed(3)
    for ms in ([96, 1, 33, 32, 31, 64, 5, 17, 40, 2, 50], [96, 96, 96]):
        M = sum(ms)
        A = torch.rand(M, M, generator=gen) * 0.05
        W = torch.rand(M, M, generator=gen)
        offs = np.concatenate([[0], np.cumsum(ms)])
        mask = torch.zeros(M, M)
        for a, b in zip(offs[:-1], offs[1:]):
            mask[a:b, a:b] = 1
        A = A * mask
        A.fill_diagonal_(0)
        U0 = torch.randn(M, 32, generator=gen)
        U = ops.gagm_solve(A.cuda(), W.cuda(), U0.cuda(), ms, max_iter=30).cpu().numpy()
        ref = mgm_port.gagm(A, W, U0, ms, 32, max_iter=30, precise=True).numpy()
        assert np.array_equal(U, ref)
