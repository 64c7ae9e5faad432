"""Golden vectors for HiPPI (multi_graph_matching.py:392-449) from the reference's OWN class, imported through the shim.
TEST INFRASTRUCTURE, build container only:   python -m oracle.gen_golden_hippi   -> tests/golden/hippi.npz
The reference runs in fp32; the same module is also run on float64 copies of the inputs (the noise-free limit the
fp64-accumulating device operators are compared with tightly)."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def inputs(seed=0, ms=(20, 25, 18), d=32):
    g = torch.Generator().manual_seed(seed)
    M = sum(ms)
    B = torch.rand(M, M, generator=g)
    W = (B + B.t()) * 0.5 / M                                    # symmetric, non-negative multi-graph similarity
    U0 = torch.softmax(torch.randn(M, d, generator=g) * 2.0, dim=1)
    return W, U0, torch.tensor(ms, dtype=torch.int32), d


def main():
    from oracle import ref_shim
    ref = ref_shim.load()
    torch.set_num_threads(1)
    W, U0, ms, d = inputs()
    out = {"W": W.numpy(), "U0": U0.numpy(), "ms": ms.numpy(), "d": np.array(d)}
    for iters in (1, 3):
        for proj in ("sinkhorn", "hungarian"):
            h = ref.HiPPI(max_iter=iters)
            with torch.no_grad():
                out[f"U_{proj}_{iters}_f32"] = h(W, U0, ms, d, projector=proj).numpy()
                out[f"U_{proj}_{iters}_f64"] = h(W.double(), U0.double(), ms, d, projector=proj).numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "hippi.npz"), **out)
    for k, v in out.items():
        if k.startswith("U_"):
            print(k, v.shape, float(np.abs(v).max()), float(v.sum()))
    print("f32 vs f64, sinkhorn 3 iterations:", float(np.abs(out["U_sinkhorn_3_f32"] - out["U_sinkhorn_3_f64"]).max()))


if __name__ == "__main__":
    main()
