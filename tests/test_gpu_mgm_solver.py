"""GPU parity of the GA-GM solver and of MGM3_unsup end to end."""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import mgm_port  # noqa: E402  (checker only)
from ttdg_b200 import ops, synth  # noqa: E402
from _traj import verify_trajectory  # noqa: E402

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
            binary_in = bool(((Uin == 0) | (Uin == 1)).all())
            if binary_in:
                assert np.array_equal(U1, ref64.float().numpy()), (k, "vs float64 oracle")
            if binary_in:
                assert np.array_equal(U1, Uout), (k, "vs reference fp32 trace")
            else:       # near-ties of V that fp32 noise breaks arbitrarily: both optimal up to that noise
                o1, o2 = float((V.numpy() * U1).sum()), float((V.numpy() * Uout).sum())
                assert abs(o1 - o2) <= 1e-5 * abs(o2) and U1.sum() == Uout.sum()
        else:
            np.testing.assert_allclose(U1, ref64.float().numpy(), atol=1e-6)
            np.testing.assert_allclose(U1, Uout, atol=5e-4 if tau < 0.02 else 5e-5)


@pytest.mark.parametrize("path", MGM_FILES, ids=IDS)
def test_gagm_full_solve_every_iteration_vs_float64_oracle(path):
    """Full solve from the reference's (A, Wds, U0): every iteration of the CUDA trajectory equals the float64
    oracle step applied to the previous CUDA state (bit for bit on Hungarian steps unless the LAP has an exact
    tie, in which case both answers must be optimal), the stage schedule and stopping rule are the reference's;
    (the oracle's FREE-running solve can still end elsewhere: its Sinkhorn-stage states differ from ours by ~1e-13,
    which the first Hungarian step may amplify - that is the chaos, not an error).  G == 2 problems are stable
    and must reproduce the reference's own fp32 result bit for bit."""
    g = np.load(path)
    if "A" not in g.files:
        pytest.skip("inputs not stored for the large case (covered end to end below)")
    ms = [int(x) for x in g["sizes"]]
    A, W, U0 = T(g["A"]), T(g["Wds"]), T(g["U0"])
    U, info, trace, meta = ops.gagm_solve(A.cuda(), W.cuda(), U0.cuda(), ms, trace_cap=1300)
    info = info.cpu().tolist()
    assert info[0] <= 1300
    verify_trajectory(A, W, U0, ms, trace, meta, info)
    U = U.cpu().numpy()
    assert np.array_equal(U, trace[info[0]].float().cpu().numpy())
    assert info[2] == 0 or set(np.unique(U)) <= {0.0, 1.0}
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
        np.testing.assert_allclose(aux["Wds"].cpu().numpy(), g["Wds"], atol=1e-5, rtol=1e-4)   # fp32 x/tau ulp is 8e-6
        np.testing.assert_allclose(aux["A"].cpu().numpy(), g["A"], atol=2e-7)
        np.testing.assert_allclose(aux["U0"].cpu().numpy(), g["U0"], atol=2e-5, rtol=1e-5)
    # (1) against the reference's own fp32 gradients: they carry fp32 noise from 20 exp/log Sinkhorn steps, so the
    #     tolerance is relative to each tensor's scale
    def close(a, ref, what, rel=2e-2):
        np.testing.assert_allclose(a, ref, atol=rel * np.abs(ref).max(), rtol=2e-3, err_msg=what)
    for i, n in enumerate(nodes):
        close(n.grad.cpu().numpy(), g[f"grad_nodes_{i}"], f"nodes {i}")
    close(m.node_affinity.fc_M[2].weight.grad.cpu().numpy(), g["grad_aff_fc_M.2.weight"], "fc_M.2.weight")
    for k, p in m.node_affinity.named_parameters():
        if "grad_aff_" + k + "_sum" in g.files:
            gr = p.grad.cpu()
            sub = gr[::16, ::16] if gr.dim() == 2 else gr[::16]
            close(sub.numpy(), g["grad_aff_" + k + "_sub16"], k)
            np.testing.assert_allclose(float((gr.double() ** 2).sum()), float(g["grad_aff_" + k + "_sumsq"]), rtol=1e-2)
    # (2) against the oracle port run in float64 on the same inputs and the same forced U: tight
    sd64 = {k: v.double().requires_grad_(k.startswith("node_affinity.")) for k, v in sd.items()}
    n64 = [n.double().requires_grad_(True) for n in synth.mgm_inputs(sizes, seed)[0]]
    l64 = mgm_port.mgm3_unsup_forward(sd64, n64, labels, synth.universe(0).double(), [k.double() for k in masks],
                                      U_override=T(g["U"].astype(np.float64)))
    l64.backward()
    #     (the affinity matrix crosses to the Sinkhorn kernel as fp32, like the reference's tensor: its ulp / tau is
    #     ~1e-5 in the log domain, and 20 Sinkhorn steps amplify it - that bounds the agreement with an all-float64
    #     pipeline; the Sinkhorn backward alone is checked to 1e-6 in test_sinkhorn_small_bwd_vs_float64_autograd)
    np.testing.assert_allclose(loss.item(), l64.item(), rtol=2e-5)
    for i, n in enumerate(nodes):
        ref = n64[i].grad.float().numpy()
        np.testing.assert_allclose(n.grad.cpu().numpy(), ref, atol=5e-3 * np.abs(ref).max(), rtol=5e-4, err_msg=f"nodes {i} vs f64")
    for k, p in m.node_affinity.named_parameters():
        ref = sd64["node_affinity." + k].grad.float().numpy()
        if k == "fc_M.2.bias":      # exactly zero in exact arithmetic (Sinkhorn is shift invariant)
            assert abs(float(p.grad)) < 1e-7
            continue
        np.testing.assert_allclose(p.grad.cpu().numpy(), ref, atol=5e-3 * np.abs(ref).max(), rtol=5e-4, err_msg=k + " vs f64")
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
    # same solver call with the trajectory recorded: identical result, every iteration verified
    U2, info, trace, meta = ops.gagm_solve(aux["A"], aux["Wds"], aux["U0"], list(sizes), trace_cap=1300)
    assert np.array_equal(U, U2.cpu().numpy()) and torch.equal(info[:8], aux["info"][:8])     # [8:] = cycle counters
    verify_trajectory(aux["A"].cpu(), aux["Wds"].cpu(), aux["U0"].cpu(), list(sizes), trace, meta, info.cpu().tolist())
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
    # (the third case: more graphs than CTAs AND small enough for the tensor-core products of the Sinkhorn stage - two graphs per CTA)
    for ms in ([96, 1, 33, 32, 31, 64, 5, 17, 40, 2, 50], [96, 96, 96], [12, 20, 9, 15, 30, 7, 18, 25, 11, 16, 1, 33],
               [1, 5, 33, 2, 57]):                          # one graph per CTA (shared-memory Hungarian path) with 1- and 2-node graphs
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
        U, info, trace, meta = ops.gagm_solve(A.cuda(), W.cuda(), U0.cuda(), ms, max_iter=30, trace_cap=300)
        verify_trajectory(A, W, U0, ms, trace, meta, info.cpu().tolist(), max_iter=30)
        assert np.array_equal(U.cpu().numpy(), trace[int(info[0])].float().cpu().numpy())


def test_hippi_matches_reference(golden_dir):
    """HiPPI (mgm:392-449) through the device operators against the reference's own class (oracle/gen_golden_hippi.py):
    Hungarian projections bit for bit, Sinkhorn projections (tau = 1/200, 20 iterations, dummy rows) to fp32 noise."""
    from adapteacher.modeling.GModule.multi_graph_matching import HiPPI
    g = np.load(f"{golden_dir}/hippi.npz")
    W, U0 = torch.from_numpy(g["W"]).cuda(), torch.from_numpy(g["U0"]).cuda()
    ms, d = torch.from_numpy(g["ms"]), int(g["d"])
    for iters in (1, 3):
        h = HiPPI(max_iter=iters)
        U = h(W, U0, ms, d, projector="hungarian").cpu().numpy()
        assert np.array_equal(U, g[f"U_hungarian_{iters}_f32"]) and np.array_equal(U, g[f"U_hungarian_{iters}_f64"])
        U = h(W, U0, ms, d, projector="sinkhorn").cpu().numpy()
        np.testing.assert_allclose(U, g[f"U_sinkhorn_{iters}_f64"], rtol=2e-4, atol=1e-7)
        np.testing.assert_allclose(U, g[f"U_sinkhorn_{iters}_f32"], rtol=2e-4, atol=1e-7)
        assert h.last_iterations == iters
    with pytest.raises(NameError):
        HiPPI()(W, U0, ms, d, projector="softmax")
    # runs to convergence (or max_iter) without a host-side projector
    h = HiPPI(max_iter=50)
    U = h(W, U0, ms, d, projector="hungarian")
    assert h.last_iterations <= 50 and float(U.sum()) == float(sum(ms.tolist()))


@pytest.mark.parametrize("mode", [1, 2, 3, 4])
@pytest.mark.parametrize("sizes,seed", [((33, 34, 33, 33, 46, 30, 34, 38), 77), ((46, 33, 22, 41, 33, 32, 33, 34), 21), ((60,) * 8, 16), ((23, 40, 31, 57), 3), ((30, 30, 30, 30, 30), 5), ((12, 20), 9)])
def test_certified_fast_lap_gives_the_same_solution(sizes, seed, mode):
    """Certified fast paths of the Hungarian projections (1: row-reduction start, 2: Jacobi-auction start, 3 - the default -
    the lean solve: auction rounds + Dijkstra without SciPy's bookkeeping, 4: the same with label-correcting rounds instead of
    Dijkstra; each followed by the optimality / uniqueness
    certificate and the SciPy-order solve as the fall-back, lap.cuh): the whole GA-GM solve -
    every projection of ~200 iterations - must give the same U and the same iteration counts as the SciPy-order solver."""
    from ttdg_b200 import _C
    from adapteacher.modeling.GModule.multi_graph_matching import MGM3_unsup
    m = MGM3_unsup(2, 32).cuda()
    m.load_state_dict(synth.mgm_unsup_state(0))
    nodes, labels, masks = synth.mgm_inputs(sizes, seed)
    m.debug_keep_masks = [k.cuda() for k in masks]
    with torch.no_grad():
        m([n.cuda() for n in nodes], [l.cuda() for l in labels], synth.universe(0).cuda())
    aux = m.last_aux
    L = _C.lib()
    prev = L.ttdg_gagm_set_lap_fast(0)
    try:
        U0, i0 = ops.gagm_solve(aux["A"], aux["Wds"], aux["U0"], list(sizes), return_info=True)
        L.ttdg_gagm_set_lap_fast(mode)
        U1, i1 = ops.gagm_solve(aux["A"], aux["Wds"], aux["U0"], list(sizes), return_info=True)
    finally:
        L.ttdg_gagm_set_lap_fast(prev)
    assert torch.equal(U0, U1)
    a, b = i0.tolist(), i1.tolist()
    assert a[:5] == b[:5]                                     # same iteration schedule (stages, Sinkhorn / Hungarian counts, LAP calls)
