"""tcgen05 / TMA implicit-GEMM convolution (csrc/conv_tc.cu) against torch.nn.functional, in both math modes:
'tf32x3' (hi/lo operand split, fp32-grade - the parity configuration) and 'tf32' (single pass)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from ttdg_b200 import detector as det  # noqa: E402


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


@pytest.fixture(autouse=True)
def _restore_mode():
    old = det.CONV_MODE[0]
    yield
    det.set_conv_mode(old)


CASES = [
    # cin, cout, k, pad, H, W, N
    (64, 64, 1, 0, 16, 16, 2),            # BN_TILE 64
    (64, 256, 1, 0, 8, 24, 1),
    (128, 128, 3, 1, 12, 12, 2),          # 3x3: taps are TMA coordinate shifts, padding = OOB fill
    (256, 256, 3, 1, 14, 14, 5),          # mask-head shape: 14 is not a power of two (partial boxes)
    (256, 128, 3, 1, 14, 14, 40),         # ... with enough rois that the 14 x 1 x 9-image box (126 of 128 rows) is chosen
    (64, 64, 3, 1, 6, 10, 23),            # odd everything: 10 x 6 x 2 = 120 rows
    (256, 256, 3, 1, 128, 128, 1),        # BW = 128
    (256, 64, 3, 1, 8, 8, 3),             # several images per tile
    (12544, 1024, 1, 0, 1, 1, 300),       # box-head fc1 as a 1x1 conv on N = rows
    (2048, 256, 1, 0, 4, 4, 2),
]


@pytest.mark.parametrize("mode,rtol", [("tf32x3", 2e-5), ("tf32", 4e-3)])
@pytest.mark.parametrize("cin,cout,k,pad,H,W,N", CASES)
def test_conv_tc_forward(mode, rtol, cin, cout, k, pad, H, W, N):
    det.set_conv_mode(mode)
    g = torch.Generator().manual_seed(cin + cout + k + H)
    layer = det.Conv2d(cin, cout, k, 1, pad, bias=True).cuda()
    w = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    b = torch.randn(cout, generator=g)
    layer.load_state_dict({"weight": w, "bias": b})
    x = torch.randn(N, cin, H, W, generator=g)
    res = torch.randn(N, cout, H, W, generator=g)
    with torch.no_grad():
        y = layer(nhwc(x).cuda(), relu=True, residual=nhwc(res).cuda(), res_mode=1)
    ref = F.relu(F.conv2d(x, w, b, 1, pad) + res)
    err = float((nchw(y.cpu()) - ref).abs().max() / ref.abs().max())
    rtol = rtol * max(1.0, (cin * k * k / 1024.0) ** 0.5)          # fp32 accumulation error grows ~ sqrt(K)
    assert err < rtol, err
    det.set_conv_mode("simt")
    with torch.no_grad():
        y2 = layer(nhwc(x).cuda(), relu=True, residual=nhwc(res).cuda(), res_mode=1)
    assert float((y2 - y).abs().max() / ref.abs().max()) < rtol


@pytest.mark.parametrize("mode,rtol", [("tf32x3", 3e-5), ("tf32", 6e-3)])
@pytest.mark.parametrize("cin,cout,k,pad,H,W,N", [(128, 128, 3, 1, 10, 12, 2), (512, 128, 1, 0, 8, 8, 2), (256, 256, 3, 1, 16, 16, 1),
                                                 (128, 128, 3, 1, 14, 14, 30)])
def test_conv_tc_dgrad(mode, rtol, cin, cout, k, pad, H, W, N):
    det.set_conv_mode(mode)
    g = torch.Generator().manual_seed(cin + cout + k)
    layer = det.Conv2d(cin, cout, k, 1, pad, bias=False).cuda()
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).requires_grad_(True)
    layer.load_state_dict({"weight": w.detach()})
    x = torch.randn(N, cin, H, W, generator=g).requires_grad_(True)
    up = torch.randn(N, cout, H, W, generator=g)
    (F.conv2d(x, w, None, 1, pad) * up).sum().backward()
    xc = nhwc(x.detach()).cuda().requires_grad_(True)
    (layer(xc) * nhwc(up).cuda()).sum().backward()
    err = float((nchw(xc.grad.cpu()) - x.grad).abs().max() / x.grad.abs().max())
    assert err < rtol, err
    gw = layer.weight.grad[:, :, :cin, :cout].permute(3, 2, 0, 1).cpu()
    assert float((gw - w.grad).abs().max() / w.grad.abs().max()) < rtol           # weight gradient: tensor cores too (below)


WGRAD_CASES = [
    # cin, cout, k, pad, H, W, N
    (128, 64, 1, 0, 8, 8, 2),              # BN_TILE 64, one 32-pixel patch = 4 x 8 pixels
    (128, 128, 3, 1, 10, 12, 2),           # partial patches in both directions + padding taps
    (256, 256, 3, 1, 14, 14, 5),           # mask-head shape
    (256, 256, 3, 1, 64, 64, 2),           # many pixel blocks: split over CTAs, fp32 atomics
    (512, 512, 3, 1, 4, 4, 3),             # res5 shape: patch spans images
    (256, 1024, 1, 0, 16, 16, 1),
    (1024, 256, 1, 0, 16, 16, 2),
    (12544, 1024, 1, 0, 1, 1, 70),         # box-head fc1: pixels = rows
    (1024, 1024, 1, 0, 1, 1, 33),
]


@pytest.mark.parametrize("mode,rtol", [("tf32x3", 2e-5), ("tf32", 5e-3)])
@pytest.mark.parametrize("cin,cout,k,pad,H,W,N", WGRAD_CASES)
def test_wgrad_tc(mode, rtol, cin, cout, k, pad, H, W, N):
    """Weight gradient with MN-major tcgen05 operands (pixels = GEMM k) against torch autograd and against the
    CUDA-core fp32 kernel; the gradient is accumulated INTO dw (checked with a non-zero start)."""
    L = det._C.lib()
    assert L.ttdg_wgrad_tc_supported(cin, cout, 1) == 1 and L.ttdg_wgrad_tc_supported(cin, cout, 2) == 0
    g = torch.Generator().manual_seed(cin * 3 + cout + k + H)
    x = torch.randn(N, cin, H, W, generator=g)
    w = torch.zeros(cout, cin, k, k, requires_grad=True)
    dy = torch.randn(N, cout, H + 2 * pad - k + 1, W + 2 * pad - k + 1, generator=g)
    F.conv2d(x, w, None, 1, pad).backward(dy)
    ref = w.grad.permute(2, 3, 1, 0).contiguous()                                # [R][S][Cin][Cout]
    xc, dc = nhwc(x).cuda(), nhwc(dy).cuda()
    start = torch.randn(k, k, cin, cout, generator=g)
    dw = start.cuda()
    p = lambda t: None if t is None else t.data_ptr()
    rc = L.ttdg_wgrad_tc(p(xc), p(dc), int(mode == "tf32x3"), N, H, W, cin, cout, k, k, 1, pad, p(dw), None)
    torch.cuda.synchronize()
    assert rc == 0
    got = dw.cpu() - start
    scale = float(ref.abs().max())
    assert float((got - ref).abs().max()) / scale < rtol * max(1.0, (N * H * W / 1024.0) ** 0.5)
    dw2 = torch.zeros(k, k, cin, cout, device="cuda")
    assert L.ttdg_conv_wgrad(p(xc), p(dc), N, H, W, cin, cout, k, k, 1, pad, p(dw2), None) == 0
    assert float((dw2.cpu() - got).abs().max()) / scale < rtol * max(1.0, (N * H * W / 1024.0) ** 0.5)


def test_wgrad_tc_rejects_unsupported():
    L = det._C.lib()
    x = torch.zeros(1, 4, 4, 64, device="cuda")
    assert L.ttdg_wgrad_tc(x.data_ptr(), x.data_ptr(), 1, 1, 4, 4, 64, 64, 1, 1, 1, 0, x.data_ptr(), None) != 0
    x = torch.zeros(1, 4, 4, 128, device="cuda")                       # stride 2 is for 1x1 convs only
    assert L.ttdg_wgrad_tc(x.data_ptr(), x.data_ptr(), 1, 1, 4, 4, 128, 128, 3, 3, 2, 1, x.data_ptr(), None) != 0


def test_fpn_upsample_add_epilogue_tc():
    det.set_conv_mode("tf32x3")
    g = torch.Generator().manual_seed(7)
    lat = det.Conv2d(512, 256, 1, 1, 0, bias=True).cuda()
    lat.load_state_dict({"weight": torch.randn(256, 512, 1, 1, generator=g) * 0.05, "bias": torch.randn(256, generator=g)})
    x, coarse = torch.randn(2, 512, 16, 16, generator=g), torch.randn(2, 256, 8, 8, generator=g)
    with torch.no_grad():
        y = lat(nhwc(x).cuda(), residual=nhwc(coarse).cuda(), res_mode=2)
    sd = lat.state_dict()
    ref = F.conv2d(x, sd["weight"].cpu(), sd["bias"].cpu()) + F.interpolate(coarse, scale_factor=2.0, mode="nearest")
    assert float((nchw(y.cpu()) - ref).abs().max() / ref.abs().max()) < 2e-5


@pytest.mark.parametrize("cin,cout,k,pad,H,W,N,norm,relu", [
    (128, 128, 3, 1, 4, 4, 2, False, False), (128, 128, 3, 1, 4, 4, 2, True, True), (512, 128, 1, 0, 4, 4, 2, True, True),
    (128, 512, 1, 0, 4, 4, 2, True, False), (128, 128, 3, 1, 10, 12, 2, True, True), (2048, 512, 1, 0, 2, 2, 2, True, True),
    (512, 512, 3, 1, 2, 2, 2, True, True)])
def test_conv_tc_layer_vs_torch(cin, cout, k, pad, H, W, N, norm, relu):
    """One layer (conv + FrozenBN + ReLU), forward and input gradient, 3xTF32 with chunked accumulation: fp32-grade
    (measured ~2e-6 relative, the CUDA-core fp32 kernel gives ~1e-6)."""
    from oracle import detector_port as dp
    det.set_conv_mode("tf32x3")
    g = torch.Generator().manual_seed(cin + cout + k)
    layer = det.Conv2d(cin, cout, k, 1, pad, bias=False, norm=norm).cuda()
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).requires_grad_(True)
    sd = {"weight": w.detach()}
    if norm:
        sd.update({"norm.weight": torch.rand(cout, generator=g) + 0.5, "norm.bias": torch.randn(cout, generator=g) * 0.1,
                   "norm.running_mean": torch.randn(cout, generator=g) * 0.1, "norm.running_var": torch.rand(cout, generator=g) + 0.5})
    layer.load_state_dict(sd)
    x = torch.randn(N, cin, H, W, generator=g).requires_grad_(True)
    up = torch.randn(N, cout, H, W, generator=g)
    y = F.conv2d(x, w, None, 1, pad)
    if norm:
        y = dp.frozen_bn(y, {"n." + kk[5:]: v for kk, v in sd.items() if kk.startswith("norm.")}, "n")
    if relu:
        y = F.relu(y)
    up = up * (y.detach().abs() > 1e-3)            # keep the functional away from the ReLU kink
    (y * up).sum().backward()
    xc = nhwc(x.detach()).cuda().requires_grad_(True)
    yc = layer(xc, relu=relu)
    (yc * nhwc(up).cuda()).sum().backward()
    rel = lambda a, r: float((a - r).abs().max() / r.abs().max())
    assert rel(nchw(yc.detach().cpu()), y.detach()) < 1e-5
    assert rel(nchw(xc.grad.cpu()), x.grad) < 1e-5


@pytest.mark.parametrize("mode,rtol", [("tf32x3", 1e-5), ("tf32", 6e-3)])
@pytest.mark.parametrize("cin,cout,H,W,N", [(256, 128, 16, 16, 2), (256, 512, 16, 16, 2), (512, 1024, 9, 11, 3), (1024, 2048, 6, 6, 2),
                                           (128, 64, 130, 130, 1)])
def test_conv_tc_stride2_1x1(mode, rtol, cin, cout, H, W, N):
    """The strided 1x1 convs of res3/res4/res5's first block (STRIDE_IN_1X1, projection shortcut): forward through TMA
    element strides, data gradient scattered to the even pixels, weight gradient from the strided X - against torch."""
    det.set_conv_mode(mode)
    g = torch.Generator().manual_seed(cin + cout + H)
    layer = det.Conv2d(cin, cout, 1, 2, 0, bias=False).cuda()
    w = (torch.randn(cout, cin, 1, 1, generator=g) / cin ** 0.5).requires_grad_(True)
    layer.load_state_dict({"weight": w.detach()})
    x = torch.randn(N, cin, H, W, generator=g).requires_grad_(True)
    y = F.conv2d(x, w, None, 2, 0)
    up = torch.randn(y.shape, generator=g)
    (y * up).sum().backward()
    xc = nhwc(x.detach()).cuda().requires_grad_(True)
    yc = layer(xc)
    (yc * nhwc(up).cuda()).sum().backward()
    rel = lambda a, r: float((a - r).abs().max() / r.abs().max())
    assert rel(nchw(yc.detach().cpu()), y.detach()) < rtol
    assert rel(nchw(xc.grad.cpu()), x.grad) < rtol
    assert rel(layer.weight.grad[:, :, :cin, :cout].permute(3, 2, 0, 1).cpu(), w.grad) < rtol * max(1.0, (N * H * W / 4096.0) ** 0.5)
    det.set_conv_mode("simt")                                  # and the CUDA-core kernels agree
    xs = nhwc(x.detach()).cuda().requires_grad_(True)
    layer.weight.grad = None
    ys = layer(xs)
    (ys * nhwc(up).cuda()).sum().backward()
    assert rel(ys.detach(), yc.detach()) < rtol and rel(xs.grad, xc.grad) < rtol


@pytest.mark.parametrize("mode,rtol", [("tf32x3", 1e-5), ("tf32", 6e-3)])
@pytest.mark.parametrize("N,H,W", [(2, 64, 96), (1, 512, 512), (3, 40, 24)])
def test_stem_tc(mode, rtol, N, H, W):
    """d2 BasicStem (7x7 stride 2 pad 3 + FrozenBN + ReLU) on tensor cores from the padded image (overlapping 8-pixel TMA
    windows) against torch and against the CUDA-core path; H, W not multiples of 32 exercise size_divisibility padding."""
    det.set_conv_mode(mode)
    g = torch.Generator().manual_seed(H + W)
    stem = det.Stem().cuda()
    w = torch.randn(64, 3, 7, 7, generator=g) / 147 ** 0.5
    sd = {"conv1.weight": w, "conv1.norm.weight": torch.rand(64, generator=g) + 0.5, "conv1.norm.bias": torch.randn(64, generator=g) * 0.1,
          "conv1.norm.running_mean": torch.randn(64, generator=g) * 0.1, "conv1.norm.running_var": torch.rand(64, generator=g) + 0.5}
    stem.load_state_dict(sd)
    ims = [torch.randint(0, 256, (3, H, W), generator=g, dtype=torch.uint8) for _ in range(N)]
    xp, width = det.preprocess(ims, torch.device("cuda"), stem_padded=True)
    H32, W32 = (H + 31) // 32 * 32, (W + 31) // 32 * 32
    assert width == W32 and tuple(xp.shape) == (N, H32, W32 + 8, 4)
    y = stem(xp, width)
    x = torch.stack(ims).float() - torch.tensor(det.PIXEL_MEAN).view(1, 3, 1, 1)
    x = F.pad(x, (0, W32 - W, 0, H32 - H))
    scale = sd["conv1.norm.weight"] * (sd["conv1.norm.running_var"] + 1e-5).rsqrt()
    bias = sd["conv1.norm.bias"] - sd["conv1.norm.running_mean"] * scale
    ref = F.relu(F.conv2d(x, w, None, 2, 3) * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1))
    assert tuple(y.shape) == (N, H32 // 2, W32 // 2, 64)
    assert float((nchw(y.cpu()) - ref).abs().max() / ref.abs().max()) < rtol
    det.set_conv_mode("simt")
    y2 = stem(det.preprocess(ims, torch.device("cuda")))
    assert float((y2 - y).abs().max() / ref.abs().max()) < rtol


@pytest.mark.parametrize("cl", [2, 4])
@pytest.mark.parametrize("cin,cout,k,pad,H,W,N", [(256, 256, 3, 1, 32, 32, 2), (256, 256, 3, 1, 14, 14, 7), (128, 512, 1, 0, 24, 40, 1),
                                                 (1024, 64, 1, 0, 16, 16, 3)])
def test_conv_tc_cluster_multicast(cl, cin, cout, k, pad, H, W, N):
    """Thread-block clusters along the pixel tiles with the weight tile delivered by TMA multicast (each CTA loads 1 / cl of
    it): the same MMAs in the same order, so the result must be BIT-IDENTICAL to the cluster-less launch - including grids
    that are padded to whole clusters (7 x 2 tiles at 14 x 14) and the 64-wide N tile."""
    from ttdg_b200 import _C
    det.set_conv_mode("tf32x3")
    g = torch.Generator().manual_seed(cl + cin + H)
    layer = det.Conv2d(cin, cout, k, 1, pad, bias=True).cuda()
    layer.load_state_dict({"weight": torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5, "bias": torch.randn(cout, generator=g)})
    x = nhwc(torch.randn(N, cin, H, W, generator=g)).cuda()
    L = _C.lib()
    prev = L.ttdg_conv_tc_set_cluster(1)
    try:
        with torch.no_grad():
            y1 = layer(x, relu=True)
            assert L.ttdg_conv_tc_set_cluster(cl) == 1
            y2 = layer(x, relu=True)
        torch.cuda.synchronize()
    finally:
        L.ttdg_conv_tc_set_cluster(prev)
    assert torch.equal(y1, y2)
    assert L.ttdg_conv_tc_set_cluster(3) == -1


@pytest.mark.parametrize("mode", ["tf32x3", "tf32", "bf16"])
@pytest.mark.parametrize("cin,cout,k,pad,stride,H,W,N,res", [
    (64, 256, 1, 0, 1, 40, 24, 2, 1),        # short-K 1x1 with residual (the layers the transposed epilogue is for)
    (256, 64, 1, 0, 1, 32, 32, 2, 0),        # 64-wide N tile
    (256, 256, 3, 1, 1, 14, 14, 7, 0),       # 14 x 14 boxes: 126 of the 128 tile rows, partial last image group
    (256, 512, 1, 0, 2, 26, 38, 1, 0),       # strided 1x1 (TMA element strides), ragged tiles
    (128, 128, 3, 1, 1, 19, 23, 3, 1),       # ragged tiles with residual
    (64, 256, 1, 0, 1, 64, 64, 9, 1),        # many tiles per SM: the residual / store pipeline of the TMA epilogue wraps around
    (256, 128, 1, 0, 1, 128, 128, 3, 0),     # the same without a residual (stores only)
])
def test_conv_tc_transposed_epilogue_is_bit_identical(mode, cin, cout, k, pad, stride, H, W, N, res):
    """The warp-transposed (coalesced) epilogue performs the same operations per element as the row-per-thread one: both
    must give the same bits for fp32 and bf16 outputs / residuals, every box shape and ragged tiles."""
    from ttdg_b200 import _C
    det.set_conv_mode(mode)
    try:
        g = torch.Generator().manual_seed(cin + cout + H)
        layer = det.Conv2d(cin, cout, k, stride, pad, bias=True).cuda()
        layer.load_state_dict({"weight": torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5, "bias": torch.randn(cout, generator=g)})
        x = nhwc(torch.randn(N, cin, H, W, generator=g)).cuda()
        Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        r = torch.randn(N, Ho, Wo, cout, generator=g).cuda() if res else None
        if mode == "bf16":
            x = x.to(torch.bfloat16)
            r = r.to(torch.bfloat16) if res else None
        L = _C.lib()
        prev = L.ttdg_conv_tc_set_epilogue(0)
        try:
            with torch.no_grad():
                y0 = layer(x, relu=True, residual=r, res_mode=int(res))
                assert L.ttdg_conv_tc_set_epilogue(1) == 0
                y1 = layer(x, relu=True, residual=r, res_mode=int(res))
                assert L.ttdg_conv_tc_set_epilogue(3) == 1              # TMA store / TMA residual load where the layer allows it
                y3 = layer(x, relu=True, residual=r, res_mode=int(res))
                y3b = layer(x, relu=False, residual=r, res_mode=int(res))
                L.ttdg_conv_tc_set_epilogue(0)
                y0b = layer(x, relu=False, residual=r, res_mode=int(res))
            torch.cuda.synchronize()
        finally:
            L.ttdg_conv_tc_set_epilogue(prev)
        assert y0.dtype == y1.dtype and torch.equal(y0, y1)
        assert y0.dtype == y3.dtype and torch.equal(y0, y3) and torch.equal(y0b, y3b)
        assert float(y0.float().abs().sum()) > 0
        assert L.ttdg_conv_tc_set_epilogue(4) == -1
    finally:
        det.set_conv_mode("tf32x3")
