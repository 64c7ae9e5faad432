"""One convolution layer through the tensor-core kernel a few times - the target command for ncu captures of a single geometry.
    python tools/conv_layer.py N H W Cin Cout k pad stride res relu [reps]      (mode from TTDG_CONV: tf32x3 | tf32 | bf16)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ttdg-mgm_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from ttdg_b200 import detector as det  # noqa: E402

N, H, W, cin, cout, k, pad, stride, res, relu = [int(a) for a in sys.argv[1:11]]
reps = int(sys.argv[11]) if len(sys.argv) > 11 else 5
mode = os.environ.get("TTDG_CONV", "tf32x3")
det.set_conv_mode(mode)
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(1)
layer = det.Conv2d(cin, cout, k, stride, pad, bias=True).to(dev)
layer.load_state_dict({"weight": torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5, "bias": torch.randn(cout, generator=g)})
dt = torch.bfloat16 if mode == "bf16" else torch.float32
x = torch.randn(N, H, W, cin, generator=g).to(dev).to(dt)
Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
r = torch.randn(N, Ho, Wo, cout, generator=g).to(dev).to(dt) if res else None
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
with torch.no_grad():
    for i in range(reps):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        y = layer(x, relu=bool(relu), residual=r, res_mode=int(res))
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
byts = (x.numel() + y.numel() + (r.numel() if res else 0)) * x.element_size()
print("us per call", ["%.1f" % t for t in ts], "min %.1f us  %.2f TB/s  %.1f TFLOP/s" % (min(ts), byts / min(ts) / 1e6, 2.0 * N * Ho * Wo * cin * cout * k * k / min(ts) / 1e6))
if os.environ.get("TTDG_TRACE"):
    from ttdg_b200 import _C
    items = 16
    buf = torch.zeros(items * 8, dtype=torch.int64, device=dev)
    _C.lib().ttdg_conv_tc_set_trace(buf.data_ptr(), items)
    with torch.no_grad():
        flush.zero_()
        y = layer(x, relu=bool(relu), residual=r, res_mode=int(res))
    torch.cuda.synchronize()
    _C.lib().ttdg_conv_tc_set_trace(None, 0)
    t = buf.cpu().view(items, 8)
    t0 = int(t[t > 0].min())
    names = ["tma_first", "tma_last", "mma_own_acc", "mma_ops_ready", "mma_commit", "epi_drained", "epi_done", "epi_start"]
    print("cycles since the first event (CTA 0):", " ".join(names))
    for i in range(items):
        if int(t[i].max()) == 0:
            continue
        print("item %2d " % i + " ".join("%9d" % (int(v) - t0 if int(v) > 0 else -1) for v in t[i]))
