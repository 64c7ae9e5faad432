"""BASELINE.json configs[4]: Sinkhorn / affinity micro-benchmark, N = 256 / 512 / 1024, 50 iterations, 1 GPU.
Prints one JSON line per size: algorithmic GB/s (batch * N^2 * 4 * 2 * iters / time) against the measured HBM peak,
for a batch beyond L2 (>= 512 MiB) and for batch 1 ("L2-resident")."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ttdg-mgm_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from ttdg_b200 import ops  # noqa: E402

peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
dev = torch.device("cuda", 0)


def timeit(fn, reps=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs) / reps


for n in (256, 512, 1024):
    for batch in ((512 << 20) // (n * n * 4), 1):
        s = torch.randn(batch, n, n, device=dev)
        out = torch.empty_like(s)
        ms = timeit(lambda: ops.sinkhorn_stream(s, tau=0.05, max_iter=50, out=out))
        alg = batch * n * n * 4 * 2 * 50
        gbs = alg / (ms * 1e-3) / 1e9
        print(json.dumps({"op": "sinkhorn", "N": n, "batch": batch, "iters": 50, "ms": round(ms, 4), "algorithmic_GBps": round(gbs, 1),
                          "frac_of_measured_hbm": round(gbs / peak, 3)}))
# affinity pair kernel (separable form), one N x N pair
from adapteacher.modeling.GModule.utils.affinity import Affinity  # noqa: E402
aff = Affinity(256).to(dev)
for n in (64, 96):
    X = torch.randn(2 * n, 256, device=dev)
    with torch.no_grad():
        ms = timeit(lambda: aff.forward_pairs(X, [n, n], [(0, 1)]))
    flops_ref = 2 * n * n * (512 * 512 + 512) + 2 * 2 * n * 256 * 256
    print(json.dumps({"op": "affinity_fwd", "N": n, "ms": round(ms, 4), "reference_form_GFLOPs": round(flops_ref / (ms * 1e-3) / 1e9, 1)}))
