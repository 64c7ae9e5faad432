"""Trajectory verifier for the GA-GM solver (test helper).

The solver's end state is the result of ~200 discrete iterations and is chaotic (oracle/mgm_port.gagm docstring),
so instead of only comparing end states, EVERY iteration of the CUDA trajectory is checked to be one application
of the oracle's single step (float64) to the CUDA path's own previous state, and the stage / stopping control
flow is re-derived from the trajectory with the reference's rules (multi_graph_matching.py:309-383)."""
import numpy as np
import torch

from oracle import mgm_port


def verify_trajectory(A, W, U0, ms, trace, meta, info, init_tau=0.1, min_tau=1e-2, sk_gamma=0.5, max_iter=200,
                      converge_tol=1e-3, quad_weight=0.5):
    A64, W64 = A.double().cpu(), W.double().cpu()
    trace, meta = trace.cpu(), meta.cpu()
    n_it = int(info[0])
    assert torch.equal(trace[0], U0.double().cpu())
    offs = np.concatenate([[0], np.cumsum(ms)])
    tau, proj, t = init_tau, "sinkhorn", 0
    U = trace[0]
    lastU = torch.zeros_like(U)
    ambiguous = 0
    while True:
        for _ in range(max_iter):
            lastU2, lastU = lastU, U
            assert t < n_it, "CUDA solver stopped early"
            assert int(meta[t, 0]) == (1 if proj == "hungarian" else 0) and float(meta[t, 1]) == tau, (t, proj, tau, meta[t])
            ref, V = mgm_port.gagm_step(A64, W64, U, ms, 32, proj, tau, quad_weight=quad_weight, return_V=True)
            got = trace[t + 1]
            if proj == "hungarian":
                if not torch.equal(got, ref):
                    # exact ties / sub-1e-12 margins (e.g. universe columns no node uses): both must be optimal
                    ambiguous += 1
                    assert bool(((got == 0) | (got == 1)).all())
                    for g, n in enumerate(ms):
                        blk, vb = got[offs[g]:offs[g + 1]], V[offs[g]:offs[g + 1]]
                        if len(ms) == 2 and g == 0:
                            continue
                        assert blk.sum() == min(n, 32) and (blk.sum(0) <= 1).all() and (blk.sum(1) <= 1).all()
                        o_got, o_ref = float((vb * blk).sum()), float((vb * ref[offs[g]:offs[g + 1]]).sum())
                        assert abs(o_got - o_ref) <= 1e-10 * max(abs(o_ref), 1e-30), (t, g, o_got, o_ref)
            else:
                np.testing.assert_allclose(got.numpy(), ref.numpy(), atol=1e-11, rtol=1e-9, err_msg=f"iteration {t}")
            U = got
            t += 1
            if torch.norm(U - lastU) < converge_tol or torch.norm(U - lastU2) == 0:
                break
        if proj == "hungarian":
            break
        elif tau > min_tau:
            tau *= sk_gamma
        else:
            proj = "hungarian"
    assert t == n_it, (t, n_it)
    return ambiguous
