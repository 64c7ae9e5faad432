"""BASELINE.json configs[2]: the bf16 backbone (ttdg_conv_tc_bf16: tcgen05.mma.kind::f16, bf16 activations and weight copies in
HBM, fp32 TMEM accumulation, fp32 master weights) with the fp32 pyramid / heads / matching stage around it.

Layer level: against torch on the SAME bf16-rounded operands with fp32 accumulation (what the kernel computes: tight).
Network level: against the fp32 oracle - bf16 has 8 mantissa bits, so this is a REPORT of what the precision costs (features,
detections, mask mIoU), with loose sanity bounds; the numbers are written to gpurun_out/parity_bf16.json."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import _parity  # noqa: E402
from oracle import detector_port as dp  # noqa: E402  (checker only)
from ttdg_b200 import synth  # noqa: E402
from ttdg_b200 import detector as det  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def bf(x):
    return x.to(torch.bfloat16).float()


@pytest.fixture
def bf16_mode():
    old = det.CONV_MODE[0]
    det.set_conv_mode("bf16")
    yield
    det.set_conv_mode(old)


@pytest.mark.parametrize("cin,cout,k,stride,pad,H,W,N,relu,res,out_bf16", [
    (64, 256, 1, 1, 0, 16, 20, 2, True, 1, True),          # res2 conv3 + identity shortcut
    (128, 128, 3, 1, 1, 12, 10, 2, True, 0, True),         # conv2
    (256, 512, 1, 2, 0, 16, 16, 2, False, 0, True),        # strided projection shortcut (TMA element strides)
    (512, 256, 1, 1, 0, 8, 8, 2, False, 2, True),          # FPN lateral + nearest-upsampled top-down sum
    (256, 256, 3, 1, 1, 16, 16, 2, False, 0, False),       # FPN output conv: bf16 in, fp32 pyramid out
    (2048, 512, 1, 1, 0, 4, 4, 3, True, 0, True),          # long K (32 k-blocks: 4 accumulation chunks)
    (256, 64, 1, 1, 0, 14, 14, 5, True, 0, True),          # 64-wide N tile, map width not a power of two
])
def test_bf16_conv_layer_vs_torch(bf16_mode, cin, cout, k, stride, pad, H, W, N, relu, res, out_bf16):
    g = torch.Generator().manual_seed(cin + 3 * cout + k)
    norm = relu
    layer = det.Conv2d(cin, cout, k, stride, pad, bias=not norm, norm=norm).cuda()
    w = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    sd = {"weight": w}
    if norm:
        sd.update({"norm.weight": torch.rand(cout, generator=g) + 0.5, "norm.bias": torch.randn(cout, generator=g) * 0.1,
                   "norm.running_mean": torch.randn(cout, generator=g) * 0.1, "norm.running_var": torch.rand(cout, generator=g) + 0.5})
    else:
        sd["bias"] = torch.randn(cout, generator=g)
    layer.load_state_dict(sd)
    x = bf(torch.randn(N, cin, H, W, generator=g))
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    r = None
    if res == 1:
        r = bf(torch.randn(N, cout, Ho, Wo, generator=g))
    elif res == 2:
        r = bf(torch.randn(N, cout, Ho // 2, Wo // 2, generator=g))
    with torch.no_grad():
        y = layer(nhwc(x).cuda().to(torch.bfloat16), relu=relu, residual=None if r is None else nhwc(r).cuda().to(torch.bfloat16),
                  res_mode=res, out_bf16=out_bf16)
    assert y.dtype == (torch.bfloat16 if out_bf16 else torch.float32)
    ref = F.conv2d(x.double(), bf(w).double(), None, stride, pad)
    if norm:
        ref = dp.frozen_bn(ref.float(), {"n." + kk[5:]: v for kk, v in sd.items() if kk.startswith("norm.")}, "n").double()
    else:
        ref = ref + sd["bias"].double().view(1, -1, 1, 1)
    if res == 1:
        ref = ref + r.double()
    elif res == 2:
        ref = ref + F.interpolate(r.double(), scale_factor=2.0, mode="nearest")
    if relu:
        ref = F.relu(ref)
    got = nchw(y.float().cpu()).double()
    scale = float(ref.abs().max())
    if out_bf16:                                         # one bf16 rounding of the result (2^-9 relative) on top of fp32 accumulation
        np.testing.assert_allclose(got.numpy(), ref.numpy(), atol=2e-5 * scale, rtol=2 ** -8)
    else:
        np.testing.assert_allclose(got.numpy(), ref.numpy(), atol=2e-5 * scale, rtol=1e-5)


def test_bf16_backbone_and_ttt_step_vs_fp32_oracle(bf16_mode):
    """Whole network in the configs[2] mode on 4 x 256 x 256: pyramid, eval-pass detections / masks and one adaptation step."""
    cfg = dict(size=256, batch=4, num_classes=2, polyp=False, first=100)
    m, sd_det, sd_mgm, U = _parity.build_model(2)
    ims, images = _parity.images_of(cfg)
    with torch.no_grad():
        feats = m._det[0].features(images)
        ref = dp.backbone(sd_det, dp.preprocess(images))
    assert all(f.dtype == torch.float32 for f in feats)                      # the pyramid the heads and the matching stage see
    rel = [float((nchw(f.cpu()) - r).abs().max() / r.abs().max()) for f, r in zip(feats, ref)]
    rel2 = [float((nchw(f.cpu()) - r).norm() / r.norm()) for f, r in zip(feats, ref)]
    ev = _parity.eval_parity(m, sd_det, cfg, with_f64=False, log=lambda *a: None)
    tt = _parity.ttt_parity(m, sd_det, sd_mgm, U, cfg, with_f64=False, log=lambda *a: None)
    rep = {"pyramid_rel_max": rel, "pyramid_rel_l2": rel2, "eval": ev, "ttt": {k: v for k, v in tt.items() if k != "per_tensor"}}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_bf16.json"), "w") as f:
        json.dump(rep, f, indent=1)
    print(json.dumps(rep, indent=1)[:3000])
    # Measured (r02): this RANDOM-INIT network amplifies a per-layer rounding error ~75x over its depth (fp32-grade convs: 2e-6
    # per layer -> 1.5e-4 at the pyramid, tests/test_gpu_parity_configs.py), so 2^-9 per bf16 rounding arrives as ~0.2 relative
    # L2 at the pyramid: individual detections no longer pair up with the fp32 oracle's, the segmentation STATISTIC (mIoU of
    # all detections against the ground truth) moves by 4e-3.  A trained network is far better conditioned; what is asserted
    # here is the kernel-level exactness above plus these sanity bounds.
    assert max(rel2) < 0.35, (rel, rel2)
    fr = ev["free_running_gpu_vs_fp32"]
    assert fr["miou_delta_all"] < 2e-2, fr
    assert ev["mask_branch_forced_detections"]["miou_delta"] < 2e-2, ev["mask_branch_forced_detections"]
    assert np.isfinite(tt["loss_gpu"]) and tt["loss_rel_gpu_vs_fp32"] < 5e-2, tt["loss_rel_gpu_vs_fp32"]
    assert tt["weight_max_abs_gpu"] < 1e-4
