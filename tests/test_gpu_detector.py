"""GPU parity of the detector half (csrc/conv.cu, csrc/detect.cu, ttdg_b200/detector.py) against the CPU oracle
restatement of Detectron2's Mask R-CNN (oracle/detector_port.py, parity unpinned - d2 is not installable) and against
torch.nn.functional for the convolution family."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import detector_port as dp  # noqa: E402  (checker only)
from ttdg_b200 import synth  # noqa: E402
from ttdg_b200 import detector as det  # noqa: E402


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


CONV_CASES = [
    # cin, cout, k, stride, pad, H, W, N
    (64, 64, 1, 1, 0, 16, 20, 2),
    (64, 256, 1, 1, 0, 9, 7, 1),
    (256, 128, 1, 2, 0, 16, 16, 2),
    (128, 128, 3, 1, 1, 12, 10, 2),
    (256, 256, 3, 1, 1, 8, 8, 1),
    (3, 64, 7, 2, 3, 32, 32, 2),
    (256, 15, 1, 1, 0, 8, 8, 2),
    (256, 60, 1, 1, 0, 8, 8, 2),
    (1024, 3, 1, 1, 0, 1, 1, 37),
]


@pytest.mark.parametrize("cin,cout,k,stride,pad,H,W,N", CONV_CASES)
def test_conv_forward_vs_torch(cin, cout, k, stride, pad, H, W, N):
    g = torch.Generator().manual_seed(cin * 7 + cout + k)
    layer = det.Conv2d(cin, cout, k, stride, pad, bias=True).cuda()
    w = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    b = torch.randn(cout, generator=g)
    layer.load_state_dict({"weight": w, "bias": b})
    x = torch.randn(N, cin, H, W, generator=g)
    xin = torch.nn.functional.pad(nhwc(x), (0, layer.cin_p - cin)).cuda()
    with torch.no_grad():
        y = layer(xin, relu=True)
    ref = F.relu(F.conv2d(x, w, b, stride, pad))
    np.testing.assert_allclose(nchw(y.cpu())[:, :cout].numpy(), ref.numpy(), atol=3e-5 * float(ref.abs().max()), rtol=1e-4)
    assert (y[..., cout:] == 0).all() or layer.cout_p == cout


def test_conv_frozen_bn_residual_and_upsample_add():
    g = torch.Generator().manual_seed(1)
    layer = det.Conv2d(128, 256, 3, 1, 1, bias=False, norm=True).cuda()
    sd = {"weight": torch.randn(256, 128, 3, 3, generator=g) * 0.03, "norm.weight": torch.rand(256, generator=g) + 0.5,
          "norm.bias": torch.randn(256, generator=g), "norm.running_mean": torch.randn(256, generator=g),
          "norm.running_var": torch.rand(256, generator=g) + 0.5}
    layer.load_state_dict(sd)
    x, res = torch.randn(2, 128, 12, 16, generator=g), torch.randn(2, 256, 12, 16, generator=g)
    with torch.no_grad():
        y = layer(nhwc(x).cuda(), relu=True, residual=nhwc(res).cuda(), res_mode=1)
    ref = F.relu(dp.frozen_bn(F.conv2d(x, sd["weight"], None, 1, 1), {"n." + k[5:]: v for k, v in sd.items() if k.startswith("norm.")}, "n") + res)
    np.testing.assert_allclose(nchw(y.cpu()).numpy(), ref.numpy(), atol=3e-5, rtol=1e-4)
    lat = det.Conv2d(128, 256, 1, 1, 0, bias=True).cuda()
    lat.load_state_dict({"weight": torch.randn(256, 128, 1, 1, generator=g) * 0.1, "bias": torch.randn(256, generator=g)})
    coarse = torch.randn(2, 256, 6, 8, generator=g)
    with torch.no_grad():
        y = lat(nhwc(x).cuda(), residual=nhwc(coarse).cuda(), res_mode=2)
    sdl = lat.state_dict()
    ref = F.conv2d(x, sdl["weight"].cpu(), sdl["bias"].cpu()) + F.interpolate(coarse, scale_factor=2.0, mode="nearest")
    np.testing.assert_allclose(nchw(y.cpu()).numpy(), ref.numpy(), atol=3e-5, rtol=1e-4)


@pytest.mark.parametrize("cin,cout,k,stride,pad,H,W,N,relu,res", [
    (128, 128, 3, 1, 1, 10, 12, 2, True, False), (256, 128, 1, 2, 0, 12, 12, 2, True, False),
    (128, 512, 1, 1, 0, 9, 9, 2, True, True), (256, 256, 3, 1, 1, 8, 8, 1, False, False), (2048, 256, 1, 1, 0, 4, 4, 2, False, False)])
def test_conv_autograd_vs_torch(cin, cout, k, stride, pad, H, W, N, relu, res):
    g = torch.Generator().manual_seed(cin + cout + k + stride)
    norm = relu
    layer = det.Conv2d(cin, cout, k, stride, pad, bias=not norm, norm=norm).cuda()
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).requires_grad_(True)
    sd = {"weight": w.detach()}
    if norm:
        sd.update({"norm.weight": torch.rand(cout, generator=g) + 0.5, "norm.bias": torch.randn(cout, generator=g) * 0.1,
                   "norm.running_mean": torch.randn(cout, generator=g) * 0.1, "norm.running_var": torch.rand(cout, generator=g) + 0.5})
    else:
        b = torch.randn(cout, generator=g).requires_grad_(True)
        sd["bias"] = b.detach()
    layer.load_state_dict(sd)
    x = torch.randn(N, cin, H, W, generator=g).requires_grad_(True)
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    r = torch.randn(N, cout, Ho, Wo, generator=g).requires_grad_(True) if res else None
    up = torch.randn(N, cout, Ho, Wo, generator=g)
    # reference
    y = F.conv2d(x, w, None if norm else b, stride, pad)
    if norm:
        y = dp.frozen_bn(y, {"n." + kk[5:]: v for kk, v in sd.items() if kk.startswith("norm.")}, "n")
    if res:
        y = y + r
    if relu:
        y = F.relu(y)
    (y * up).sum().backward()
    # ours
    xc = nhwc(x.detach()).cuda().requires_grad_(True)
    rc = nhwc(r.detach()).cuda().requires_grad_(True) if res else None
    yc = layer(xc, relu=relu, residual=rc, res_mode=1 if res else 0)
    (yc * nhwc(up).cuda()).sum().backward()
    np.testing.assert_allclose(nchw(yc.detach().cpu()).numpy(), y.detach().numpy(), atol=3e-5 * float(y.detach().abs().max()), rtol=1e-4)
    scale = lambda t: 2e-5 * float(t.abs().max())
    np.testing.assert_allclose(nchw(xc.grad.cpu()).numpy(), x.grad.numpy(), atol=scale(x.grad), rtol=1e-4)
    gw = layer.weight.grad[:, :, :cin, :cout].permute(3, 2, 0, 1).cpu()
    np.testing.assert_allclose(gw.numpy(), w.grad.numpy(), atol=scale(w.grad), rtol=1e-4)
    if res:
        np.testing.assert_allclose(nchw(rc.grad.cpu()).numpy(), r.grad.numpy(), atol=scale(r.grad), rtol=1e-4)
    if not norm:
        np.testing.assert_allclose(layer.bias.grad[:cout].cpu().numpy(), b.grad.numpy(), atol=scale(b.grad), rtol=1e-4)


@pytest.fixture(scope="module")
def model_and_sd():
    sd = synth.detector_state_calibrated(0)
    m = det.MaskRCNN(2).cuda()
    missing, unexpected = m.load_state_dict(sd, strict=True)
    return m, sd


def _images(n, size, polyp=False):
    return [synth.fundus_like_image(100 + i, size, polyp)["image"] for i in range(n)]


def test_backbone_features_vs_oracle(model_and_sd):
    m, sd = model_and_sd
    ims = _images(2, 128)
    with torch.no_grad():
        feats = m.features(ims)
        ref = dp.backbone(sd, dp.preprocess(ims))
    for l, (f, r) in enumerate(zip(feats, ref)):
        assert tuple(f.shape) == (2, r.shape[2], r.shape[3], 256)
        np.testing.assert_allclose(nchw(f.cpu()).numpy(), r.numpy(), atol=2e-4 * float(r.abs().max()), rtol=2e-3, err_msg=f"p{l + 2}")


@pytest.fixture(params=["simt", "tf32x3"])
def conv_mode(request):
    old = det.CONV_MODE[0]
    det.set_conv_mode(request.param)
    yield request.param
    det.set_conv_mode(old)


@pytest.mark.parametrize("stage,blocks,cin,H", [("res3", 4, 256, 8), ("res4", 6, 512, 8), ("res5", 3, 1024, 4)])
def test_residual_stage_backward_vs_oracle(model_and_sd, conv_mode, stage, blocks, cin, H):
    """A whole residual stage (stride-2 first block with projection shortcut + identity blocks), forward and every
    gradient, on random inputs - tight, because random inputs keep pre-activations away from the ReLU kink."""
    m, sd = model_and_sd
    q = f"backbone.bottom_up.{stage}."
    seq = getattr(m.backbone.bottom_up, stage)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, cin, H, H, generator=g).requires_grad_(True)
    sdg = {k: v.clone().requires_grad_(k.startswith(q) and k.endswith("weight") and "norm" not in k) for k, v in sd.items()}
    y = x
    for b in range(blocks):
        y = dp.bottleneck(y, sdg, f"{q}{b}.", 2 if b == 0 else 1, b == 0)
    # outputs within 1e-3 of the ReLU kink are left out of the functional: one flipped mask there is ~1 % of the L2
    # norm of these tiny gradients (measured), and 3xTF32 forward noise (~5e-6) does flip one now and then
    w = torch.randn(y.shape, generator=g) * (y.detach().abs() > 1e-3)
    (y * w).sum().backward()
    for p in m.parameters():
        p.grad = None
    xc = nhwc(x.detach()).cuda().requires_grad_(True)
    yc = seq(xc)
    (yc * nhwc(w).cuda()).sum().backward()
    rel = lambda a, r: float((a - r).abs().max() / r.abs().max())
    rel2 = lambda a, r: float((a - r).norm() / r.norm())
    assert rel(nchw(yc.detach().cpu()), y.detach()) < 5e-5          # 3xTF32 products + fp32 accumulation over 3-6 blocks
    # gradients in L2: a forward difference of 1e-5 can flip a single ReLU mask out of ~1e5 activations, which moves a
    # few gradient entries by ~1e-2 of the maximum (same effect as in test_backbone_backward_vs_oracle) but not the norm
    tol = 2e-5 if det.CONV_MODE[0] == "simt" else 2e-2       # interior masks can still flip under 3xTF32 noise
    assert rel2(nchw(xc.grad.cpu()), x.grad) < tol
    for b in range(blocks):
        for name in ("conv1", "conv2", "conv3", "shortcut"):
            layer = getattr(seq[b], name)
            if layer is not None:
                gr = layer.weight.grad[:, :, :layer.cin, :layer.cout].permute(3, 2, 0, 1).cpu()
                assert rel2(gr, sdg[f"{q}{b}.{name}.weight"].grad) < tol, (b, name)


def test_backbone_backward_vs_oracle(model_and_sd, conv_mode):
    """Gradients of a random linear functional of the pyramid w.r.t. the adapted parameters (res3-res5, FPN) on real
    synthetic images.  Loose by necessity: with FrozenBN-centred pre-activations a ~1e-5 forward difference flips a few
    ReLU masks, and the ORACLE's own gradients move by up to 30 % (max norm) / 5 % (L2) under a 1e-5 perturbation of
    res2 (measured); the tight check is test_residual_stage_backward_vs_oracle.  In the 3xTF32 mode the frozen stem and
    res2 run on tensor cores too (2e-6 per layer instead of 1e-6), which is that perturbation: the bound is the oracle's
    own 5 % sensitivity there."""
    m, sd = model_and_sd
    ims = _images(2, 128)
    g = torch.Generator().manual_seed(9)
    names = [k for k in sd if (k.startswith("backbone.fpn") or any(f"bottom_up.res{i}" in k for i in (3, 4, 5)))
             and (k.endswith(".weight") or k.endswith(".bias")) and ".norm." not in k]
    sdg = {k: v.clone().requires_grad_(k in names) for k, v in sd.items()}
    ref = dp.backbone(sdg, dp.preprocess(ims))
    ws = [torch.randn(f.shape, generator=g) for f in ref]
    sum((f * w).sum() for f, w in zip(ref, ws)).backward()
    for p in m.parameters():
        p.grad = None
    feats = m.features(ims)
    sum((f * nhwc(w).cuda()).sum() for f, w in zip(feats, ws)).backward()
    mods = dict(m.named_modules())
    errs = []
    for name in names:
        mod_name, kind = name.rsplit(".", 1)
        layer = mods[mod_name]
        gr = getattr(layer, kind).grad
        gr = gr[:, :, :layer.cin, :layer.cout].permute(3, 2, 0, 1) if kind == "weight" else gr[:layer.cout]
        r = sdg[name].grad
        errs.append(float((gr.cpu() - r).norm() / r.norm()))
    assert len(errs) == 58 and max(errs) < 0.15 and float(np.median(errs)) < (0.03 if conv_mode == "simt" else 0.06), \
        (max(errs), float(np.median(errs)))
    # the FPN parameters sit above every ReLU of the loss graph: tight
    for name in ("backbone.fpn_output2.weight", "backbone.fpn_output5.bias"):
        mod_name, kind = name.rsplit(".", 1)
        layer = mods[mod_name]
        gr = getattr(layer, kind).grad
        gr = gr[:, :, :layer.cin, :layer.cout].permute(3, 2, 0, 1) if kind == "weight" else gr[:layer.cout]
        np.testing.assert_allclose(gr.cpu().numpy(), sdg[name].grad.numpy(), atol=1e-3 * float(sdg[name].grad.abs().max()), rtol=2e-3)
    assert m.backbone.bottom_up.res2[0].conv1.weight.grad is None    # FREEZE_AT = 2


def _iou(a, b):
    inter = (a & b).flatten(1).sum(1).float()
    union = (a | b).flatten(1).sum(1).float()
    return torch.where(union > 0, inter / union, torch.ones_like(union))


def _match(a, b, tol):
    """For every box of a: index of the nearest box of b (L-inf) and whether it is within tol."""
    d = (a[:, None, :] - b[None, :, :]).abs().max(-1)[0]
    mn, idx = d.min(1)
    return idx, mn < tol


@pytest.mark.parametrize("size,polyp", [(128, False), (256, True)])
def test_inference_vs_oracle(model_and_sd, conv_mode, size, polyp):
    """Eval pass end to end.  Scores of different boxes can be closer than the fp32 noise of two different convolution
    summation orders, so orderings / NMS survivors may differ in a few places: boxes are compared as SETS, masks on
    the matched instances."""
    m, sd = model_and_sd
    ims = _images(2, size, polyp)
    res, feats, props, dets = m.inference(ims)
    ref, rfeats, rprops, rdets, _ = dp.inference(sd, ims)
    for n in range(2):
        pb, rb = props[n][0].cpu(), rprops[n][0]
        assert abs(len(pb) - len(rb)) <= 5
        _, ok = _match(pb, rb, 0.05)
        # 3xTF32 objectness logits differ by ~2e-6 from the oracle's: a few more top-k / NMS flips than exact fp32 products
        assert ok.float().mean() > (0.97 if conv_mode == "simt" else 0.93), ok.float().mean()
        a, b = res[n], ref[n]
        assert len(a["scores"]) == len(b["scores"]) == 100
        idx, ok = _match(a["pred_boxes"].cpu(), b["pred_boxes"], 0.1)
        ok = ok & (a["pred_classes"].cpu() == b["pred_classes"][idx])
        assert ok.float().mean() >= 0.95, ok.float().mean()
        np.testing.assert_allclose(a["scores"].cpu()[ok].numpy(), b["scores"][idx][ok].numpy(), atol=5e-4 if conv_mode == "simt" else 1e-3)
        iou = _iou(a["pred_masks"].cpu()[ok], b["pred_masks"][idx][ok])
        # matched boxes differ by up to ~0.05 px (box deltas come out of a K = 12544 FC), which moves a few boundary pixels
        if conv_mode == "simt":      # exact-fp32 products: the pipeline semantics, tight
            assert iou.mean() > 0.998 and (iou > 0.95).float().mean() > 0.98, (iou.mean(), iou.min())
        else:                        # 3xTF32 (per-layer error 2e-6, test_conv_tc_layer_vs_torch): the random-weight mask head
            assert iou.mean() > 0.97, (iou.mean(), iou.min())       # amplifies it through near-zero logits
        gt = synth.fundus_like_image(100 + n, size, polyp)["gt_masks"]

        def miou(pred):
            best = torch.stack([_iou(pred, gt[j:j + 1].expand_as(pred)) for j in range(len(gt))]).max(0)[0]
            return float(best.mean())
        # matched instances only (what "same inputs, same detections" means): within 1e-4
        assert abs(miou(a["pred_masks"].cpu()[ok]) - miou(b["pred_masks"][idx][ok])) < (1e-4 if conv_mode == "simt" else 1e-3)
        assert abs(miou(a["pred_masks"].cpu()) - miou(b["pred_masks"])) < 3e-3


def test_ttt_detections_vs_oracle(model_and_sd, conv_mode):
    m, sd = model_and_sd
    ims = _images(2, 128)
    with torch.no_grad():
        feats, props, dets = m.detect_ttt(ims)
    rfeats, rprops, rdets = dp.forward_ttt(sd, ims)
    for n in range(2):
        assert abs(len(props[n][0]) - len(rprops[n][0])) <= 5          # train mode: 2000 per level before NMS
        assert len(dets[n][0]) == len(rdets[n][0])
        idx, ok = _match(dets[n][0].cpu(), rdets[n][0], 0.1)
        ok = ok & (dets[n][2].cpu() == rdets[n][2][idx])
        assert ok.float().mean() >= 0.95
        np.testing.assert_allclose(dets[n][1].cpu()[ok].numpy(), rdets[n][1][idx][ok].numpy(), atol=5e-4)


def test_roi_align_vs_torchvision():
    import torchvision.ops as tvo
    g = torch.Generator().manual_seed(3)
    feats = [torch.randn(2, 256, 64 // (2 ** l), 64 // (2 ** l), generator=g) for l in range(4)]
    boxes = [torch.tensor([[10., 12., 60., 80.], [0., 0., 255., 255.], [100., 30., 140., 200.], [5., 5., 9., 8.], [30., 30., 30., 30.]]),
             torch.tensor([[20., 40., 220., 230.], [128., 128., 129.5, 131.]])]
    for pooled in (7, 14):
        ref = dp.roi_pool(feats, boxes, pooled)
        rois = det._rois([b.cuda() for b in boxes])
        out = det.roi_align([nhwc(f).cuda() for f in feats], rois, pooled)
        np.testing.assert_allclose(out.permute(0, 3, 1, 2).cpu().numpy(), ref.numpy(), atol=2e-5, rtol=1e-4)


def test_nms_vs_torchvision():
    import torchvision.ops as tvo
    g = torch.Generator().manual_seed(4)
    for n in (1, 63, 64, 65, 1000, 3000):
        xy = torch.rand(n, 2, generator=g) * 200
        wh = torch.rand(n, 2, generator=g) * 60 + 1
        boxes = torch.cat([xy, xy + wh], 1)
        scores = torch.rand(n, generator=g)
        cats = torch.randint(0, 3, (n,), generator=g)
        ref = tvo.batched_nms(boxes, scores, cats, 0.5)
        order = torch.argsort(scores, descending=True, stable=True)
        keep, nk = det.nms_sorted(boxes[order].cuda().contiguous(), cats[order].int().cuda().contiguous(), 0.5, n)
        got = order[keep[:int(nk[0])].long().cpu()]
        assert torch.equal(got, ref)


def _rpn_select_torch(rpn, logits_l, deltas_l, sizes, pre_topk):
    """The round-1 torch formulation (topk / argsort / gather) of find_top_rpn_proposals, as the reference for csrc/select.cu."""
    import ctypes
    from ttdg_b200 import _C
    from ttdg_b200.ops import _p, _stream
    L = _C.lib()
    A = rpn._cell.shape[0]
    N = logits_l[0].shape[0]
    dev = logits_l[0].device
    cell_h = (ctypes.c_float * (A * 4))(*rpn._cell.reshape(-1).tolist())
    boxes_l, scores_l, valid_l, lvl_l = [], [], [], []
    for l, (logits, deltas) in enumerate(zip(logits_l, deltas_l)):
        _, H, W, _ = logits.shape
        flat = logits[..., :A].reshape(N, H * W * A)
        k = min(flat.shape[1], pre_topk)
        sc, idx = torch.topk(flat, k, dim=1, sorted=True)
        boxes = torch.empty(N, k, 4, dtype=torch.float32, device=dev)
        valid = torch.empty(N, k, dtype=torch.uint8, device=dev)
        for n in range(N):
            det.check(L.ttdg_rpn_decode(_p(deltas[n]), deltas.shape[-1], _p(idx[n].contiguous()), 1, k, H, W, A, det.STRIDES[l],
                                        ctypes.cast(cell_h, ctypes.c_void_p), float(sizes[n][0]), float(sizes[n][1]), _p(boxes[n]),
                                        _p(valid[n]), _stream()), "rpn_decode")
        boxes_l.append(boxes); scores_l.append(sc); valid_l.append(valid)
        lvl_l.append(torch.full((k,), l, dtype=torch.int32, device=dev))
    boxes, scores, valid, lvl = torch.cat(boxes_l, 1), torch.cat(scores_l, 1), torch.cat(valid_l, 1).bool(), torch.cat(lvl_l)
    key = torch.where(valid & torch.isfinite(scores), scores, torch.full_like(scores, -float("inf")))
    order = torch.argsort(key, dim=1, descending=True, stable=True)
    n_valid = (key > -float("inf")).sum(1)
    out = []
    for n in range(N):
        o = order[n, :int(n_valid[n])]
        b, s_, lv = boxes[n][o], scores[n][o], lvl[o]
        keep = tvo_batched_nms(b, s_, lv, rpn.nms_thresh)[:rpn.post_topk]
        out.append((b[keep], s_[keep]))
    return out


def tvo_batched_nms(b, s, c, t):
    import torchvision.ops as tvo
    return tvo.batched_nms(b.cpu(), s.cpu(), c.cpu().long(), t).to(b.device)


@pytest.mark.parametrize("training", [True, False])
def test_rpn_device_side_selection_vs_torch(model_and_sd, training):
    """csrc/select.cu (radix-select top-k + sort + decode, cross-level ordering, compaction; no host round trip) against the
    torch.topk / argsort / torchvision batched_nms formulation on random head outputs, mixed image sizes included."""
    m, _ = model_and_sd
    rpn = m.proposal_generator
    g = torch.Generator().manual_seed(11)
    N, S = 3, 256
    sizes = [(256, 256), (200, 256), (256, 180)]
    # distinct logits over ALL levels (one permutation): the order of EQUAL scores is unspecified in torch.topk / batched_nms,
    # csrc/select.cu takes the lower index (see test_rpn_topk_ties_take_the_lowest_indices)
    shapes = [(N, S // s, S // s, 64) for s in det.STRIDES]
    tot = sum(a * b * c * d for a, b, c, d in shapes)
    perm = (torch.randperm(tot, generator=g).float() / tot - 0.5) * 8
    logits_l, o = [], 0
    for sh in shapes:
        n_el = sh[0] * sh[1] * sh[2] * sh[3]
        logits_l.append(perm[o:o + n_el].reshape(sh).cuda())
        o += n_el
    deltas_l = [(torch.randn(N, S // s, S // s, 64, generator=g) * 0.5).cuda() for s in det.STRIDES]
    deltas_l[1][0, 3, 3, 0] = float("nan")                                 # a non-finite box must be dropped, not crash
    ref = _rpn_select_torch(rpn, logits_l, deltas_l, sizes, 2000 if training else 1000)
    # feed the same head outputs through the device-side path (bypassing the head convolutions)
    import ctypes
    from ttdg_b200 import _C
    from ttdg_b200.ops import _p, _stream
    A = rpn._cell.shape[0]
    ks = [min(t.shape[1] * t.shape[2] * A, 2000 if training else 1000) for t in logits_l]
    Kt = sum(ks)
    boxes = torch.empty(N, Kt, 4, device="cuda"); scores = torch.empty(N, Kt, device="cuda"); valid = torch.empty(N, Kt, dtype=torch.uint8, device="cuda")
    vp = lambda ts: ctypes.cast((ctypes.c_void_p * 5)(*[t.data_ptr() for t in ts]), ctypes.c_void_p)
    ip = lambda vs: ctypes.cast((ctypes.c_int32 * len(vs))(*vs), ctypes.c_void_p)
    fp = lambda vs: ctypes.cast((ctypes.c_float * len(vs))(*vs), ctypes.c_void_p)
    det.check(_C.lib().ttdg_rpn_select(vp(logits_l), vp(deltas_l), ip([v for t in logits_l for v in t.shape[1:3]]), ip(list(det.STRIDES)), ip(ks),
                                       5, 64, 64, A, fp(rpn._cell.reshape(-1).tolist()), N, fp([float(v) for s_ in sizes for v in s_]),
                                       _p(boxes), _p(scores), _p(valid), _stream()), "rpn_select")
    lvl = torch.cat([torch.full((k,), l, dtype=torch.int32) for l, k in enumerate(ks)]).cuda()
    bs, ss, cats, nv = det.sort_candidates(boxes, scores, valid, lvl, 0, -float("inf"))
    keep, nk = det.nms_sorted(bs, cats, rpn.nms_thresh, rpn.post_topk)
    ob, os_, _, counts = det.gather_kept(bs, ss, cats, keep, nk, nv, rpn.post_topk, -float("inf"), False)
    # the production path: one shared-memory NMS per (image, level) + one ordering pass (levels never interact in batched_nms)
    kept = torch.empty(N, Kt, dtype=torch.uint8, device="cuda")
    det.check(_C.lib().ttdg_rpn_nms_levels(_p(boxes), _p(valid), ip(ks), 5, N, float(rpn.nms_thresh), int(rpn.post_topk), _p(kept), _stream()),
              "rpn_nms_levels")
    ob2 = torch.empty(N, rpn.post_topk, 4, device="cuda"); os2 = torch.empty(N, rpn.post_topk, device="cuda")
    counts2 = torch.empty(N, dtype=torch.int32, device="cuda")
    det.check(_C.lib().ttdg_top_candidates(_p(boxes), _p(scores), _p(kept), N, Kt, int(rpn.post_topk), -float("inf"), _p(ob2), _p(os2),
                                           _p(counts2), _stream()), "top_candidates")
    for n in range(N):
        c = int(counts[n])
        assert c == len(ref[n][0])
        assert torch.equal(os_[n, :c], ref[n][1]) and torch.equal(ob[n, :c], ref[n][0])
        assert bool(torch.isinf(os_[n, c:]).all())
        assert int(counts2[n]) == c
        assert torch.equal(os2[n], os_[n]) and torch.equal(ob2[n], ob[n])


def test_rpn_topk_ties_take_the_lowest_indices():
    """More candidates tied at the k-th value than slots left: the selection is deterministic (lowest anchor index first)."""
    import ctypes
    from ttdg_b200 import _C
    from ttdg_b200.ops import _p, _stream
    H = W = 16
    A = 15
    logits = torch.zeros(1, H, W, 64).cuda()
    logits[0, 2, 3, 4] = 5.0                                               # one clear winner, the other 3839 tie at 0
    deltas = torch.zeros(1, H, W, 64).cuda()
    k = 100
    boxes = torch.empty(1, k, 4, device="cuda"); scores = torch.empty(1, k, device="cuda"); valid = torch.empty(1, k, dtype=torch.uint8, device="cuda")
    cell = det.cell_anchors()
    one = lambda t: ctypes.cast((ctypes.c_void_p * 1)(t.data_ptr()), ctypes.c_void_p)
    ip = lambda vs: ctypes.cast((ctypes.c_int32 * len(vs))(*vs), ctypes.c_void_p)
    fp = lambda vs: ctypes.cast((ctypes.c_float * len(vs))(*vs), ctypes.c_void_p)
    det.check(_C.lib().ttdg_rpn_select(one(logits), one(deltas), ip([H, W]), ip([4]), ip([k]), 1, 64, 64, A, fp(cell.reshape(-1).tolist()), 1,
                                       fp([64.0, 64.0]), _p(boxes), _p(scores), _p(valid), _stream()), "rpn_select")
    assert float(scores[0, 0]) == 5.0 and bool((scores[0, 1:] == 0).all())
    # anchor index -> decoded box with zero deltas = the anchor itself (clipped): entries 1.. are anchors 0, 1, 2, ... in order
    flat_idx = [(2 * W + 3) * A + 4] + [i for i in range(k + 1) if i != (2 * W + 3) * A + 4][:k - 1]
    for j, idx in enumerate(flat_idx[:20]):
        pix, a = divmod(idx, A)
        y, x = divmod(pix, W)
        anc = cell[a] + torch.tensor([x * 4.0, y * 4.0, x * 4.0, y * 4.0])
        want = torch.stack((anc[0].clamp(0, 64), anc[1].clamp(0, 64), anc[2].clamp(0, 64), anc[3].clamp(0, 64)))
        assert torch.allclose(boxes[0, j].cpu(), want, atol=1e-5), (j, idx)
