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
# affinity (utils/affinity.py:44-57), one N x N pair, forward and forward + backward, separable form
# (sum_k w1_k relu(a_ik + c_jk): no N x N x 512 tensor; the reference materialises 512 N^2 floats = 2.1 GB at N = 1024)
from adapteacher.modeling.GModule.utils.affinity import Affinity  # noqa: E402
aff = Affinity(256).to(dev)
for n in (64, 96, 256, 512, 1024):
    X = torch.randn(n, 256, device=dev, requires_grad=True)
    Y = torch.randn(n, 256, device=dev, requires_grad=True)
    up = torch.randn(n, n, device=dev)
    with torch.no_grad():
        ms_f = timeit(lambda: aff(X, Y))

    def fb():
        for p in list(aff.parameters()) + [X, Y]:
            p.grad = None
        (aff(X, Y) * up).sum().backward()
    ms_fb = timeit(fb)
    flops_ref = 2 * n * n * (512 * 512 + 512) + 2 * 2 * n * 256 * 256
    flops_sep = 2 * (2 * n) * 256 * (256 + 512) + 3 * n * n * 512
    bytes_alg = 4 * (256 * 2 * n + n * n) + 4 * (2 * 256 * 256 + 512 * 512 + 1025)
    print(json.dumps({"op": "affinity", "N": n, "fwd_ms": round(ms_f, 4), "fwd_bwd_ms": round(ms_fb, 4),
                      "fwd_reference_form_GFLOPs": round(flops_ref / (ms_f * 1e-3) / 1e9, 1),
                      "fwd_separable_form_GFLOPs": round(flops_sep / (ms_f * 1e-3) / 1e9, 1),
                      "fwd_algorithmic_GBps": round(bytes_alg / (ms_f * 1e-3) / 1e9, 2),
                      "bound": "fp64 CUDA-core ALU (3 * 512 * N^2 flops between the two contractions; no tensor cores: the ReLU sits between them)"}))
