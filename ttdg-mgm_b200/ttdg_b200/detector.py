"""Detector half of the hot path: Detectron2 0.5 Mask R-CNN R50-FPN (configs/Base-RCNN-FPN.yaml +
configs/test_segment.yaml) as driven by adapteacher/modeling/meta_arch/rcnn.py:154-357, on the sm_100a kernels of
libttdg_sm100.so (csrc/conv.cu, csrc/detect.cu).

Activations are NHWC fp32 on the device.  Parameters keep Detectron2's state-dict names and shapes on load / save
(``backbone.bottom_up.res3.0.conv1.weight`` ...), but live in the kernels' layout ([R][S][Cin][Cout], channels padded
to a multiple of 4) so that the fused SGD step and the weight-gradient kernel work on them in place.  FrozenBN is
folded into a per-channel (scale, bias) epilogue.  torch ops appear only as plumbing (allocation, slicing, sort /
top-k for ordering); there is no eager or CPU fallback.
"""
import ctypes
import math
import os

import torch
import torch.nn as nn

from . import _C
from ._C import check
from .ops import _p, _stream, _need_cuda

PIXEL_MEAN = (103.530, 116.280, 123.675)
STRIDES = (4, 8, 16, 32, 64)
ANCHOR_SIZES = (32, 64, 128, 256, 512)
ANCHOR_RATIOS = (0.5, 1.0, 2.0)
RES_STAGES = (("res2", 3, 64, 256, 1), ("res3", 4, 128, 512, 2), ("res4", 6, 256, 1024, 2), ("res5", 3, 512, 2048, 2))


def _pad4(c):
    return (c + 3) // 4 * 4


# Convolution math: "tf32x3" = tcgen05 tensor cores with the hi/lo operand split (fp32-grade, the parity config),
# "tf32" = single-pass TF32 (fast config), "simt" = the CUDA-core fp32 kernel everywhere.  Layers the tensor-core
# kernel does not cover (stride 2, Cin % 32 != 0, Cout % 64 != 0) always use the CUDA-core kernel.
CONV_MODE = [os.environ.get("TTDG_CONV", "tf32x3")]
PARAM_EPOCH = [0]              # bumped by FlatSGD.step: invalidates the K-major weight copies below
WGRAD_TC = [os.environ.get("TTDG_WGRAD_TC", "1") == "1"]       # weight gradients on tensor cores (MN-major operands)
TC_STRIDE2 = [os.environ.get("TTDG_TC_STRIDE2", "1") == "1"]   # 1x1 stride-2 convs on tensor cores (TMA element strides)


def set_conv_mode(mode):
    """"bf16" = BASELINE.json configs[2]: the BACKBONE (stem output, res2-res5, FPN laterals) keeps bf16 activations in HBM
    and runs tcgen05.mma.kind::f16 from bf16 copies of the fp32 master weights; the FPN output convolutions write the fp32
    pyramid (node features of the matching stage, RoIAlign input); the heads cast their inputs to bf16 once and run kind::f16
    too (predictors write fp32); the backward runs single-pass TF32 on fp32 gradients; the matching stage is unchanged
    (fp32 / fp64)."""
    assert mode in ("simt", "tf32x3", "tf32", "bf16")
    CONV_MODE[0] = mode


def tf32_split(x):
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    check(_C.lib().ttdg_tf32_split(_p(x), _p(hi), _p(lo), x.numel(), _stream()), "tf32_split")
    return hi, lo


def _weights_kmajor(w, transposed, precise, owner=None):
    """K-major tensor-core operand copies of a [R][S][Cin][Cout] parameter: transposed = forward ([taps][Cout][Cin]),
    not transposed = data gradient (the array itself).  Cached on the owning layer until the parameter changes: an
    optimizer step (PARAM_EPOCH; the fused SGD kernel writes through raw pointers), an in-place torch update
    (``_version``) or a re-allocation (``data_ptr``)."""
    stamp = (PARAM_EPOCH[0], w._version, w.data_ptr())
    cache = owner.__dict__.setdefault("_wk_cache", {}) if owner is not None else None
    if cache is not None:
        hit = cache.get((transposed, precise))
        if hit is not None and hit[0] == stamp:
            return hit[1], hit[2]
    R, S, Cin, Cout = w.shape
    if transposed:
        hi = torch.empty(R * S, Cout, Cin, dtype=torch.float32, device=w.device)
        lo = torch.empty_like(hi) if precise else None
        check(_C.lib().ttdg_weight_transpose_split(_p(w), R * S, Cin, Cout, _p(hi), _p(lo), _stream()), "weight_transpose_split")
    else:
        hi, lo = tf32_split(w.detach().contiguous())
        if not precise:
            lo = None
    if cache is not None:
        cache[(transposed, precise)] = (stamp, hi, lo)
    return hi, lo


def _weights_bf16(w, owner):
    """[taps][Cout][Cin] bf16 K-major copy of a [R][S][Cin][Cout] fp32 master parameter (cached like _weights_kmajor)."""
    stamp = (PARAM_EPOCH[0], w._version, w.data_ptr())
    cache = owner.__dict__.setdefault("_wk_cache", {}) if owner is not None else None
    if cache is not None:
        hit = cache.get("bf16")
        if hit is not None and hit[0] == stamp:
            return hit[1]
    R, S, Cin, Cout = w.shape
    wt = torch.empty(R * S, Cout, Cin, dtype=torch.bfloat16, device=w.device)
    check(_C.lib().ttdg_weight_transpose_bf16(_p(w), R * S, Cin, Cout, _p(wt), _stream()), "weight_transpose_bf16")
    if cache is not None:
        cache["bf16"] = (stamp, wt)
    return wt


_REFRESH_TABLES = {}


def refresh_weight_copies(modules):
    """Re-derive, in ONE launch, every cached tensor-core weight copy (K-major hi / lo, data-gradient hi / lo, bf16) of the
    given layers from their (just updated) fp32 master weights - in place, so the copies keep their addresses and the cached
    TMA descriptors stay valid.  Call after FlatSGD.step (which bumps PARAM_EPOCH); layers / variants that have no cached copy
    yet are created lazily on first use as before."""
    jobs, keep = [], []
    for m in modules:
        cache = m.__dict__.get("_wk_cache")
        if not cache:
            continue
        w = m.weight
        R, S, Cin, Cout = w.shape
        for key, ent in cache.items():
            if key == "bf16":
                jobs.append((w.data_ptr(), ent[1].data_ptr(), 0, R * S, Cin, Cout, 2))
            else:
                transposed, precise = key
                jobs.append((w.data_ptr(), ent[1].data_ptr(), ent[2].data_ptr() if ent[2] is not None else 0, R * S, Cin, Cout,
                             0 if transposed else 1))
            keep.append((m, key))
    if not jobs:
        return
    sig = tuple(jobs)
    dev = modules[0].weight.device
    tab = _REFRESH_TABLES.get(sig)
    if tab is None:
        rows, tot = [], 0
        for j in jobs:
            rows.append(list(j) + [tot])
            tot += j[3] * ((j[4] + 31) // 32) * ((j[5] + 31) // 32)
        _REFRESH_TABLES.clear()                              # one live table per model is enough
        tab = (torch.tensor(rows, dtype=torch.int64).to(dev), tot)
        _REFRESH_TABLES[sig] = tab
    check(_C.lib().ttdg_weight_refresh(_p(tab[0]), len(jobs), tab[1], _stream()), "weight_refresh")
    for m, key in keep:                                      # the copies are current again
        ent = m._wk_cache[key]
        stamp = (PARAM_EPOCH[0], m.weight._version, m.weight.data_ptr())
        m._wk_cache[key] = (stamp,) + tuple(ent[1:])


RPN_NMS_PER_LEVEL = [True]     # False: the level-concatenated ordering + ttdg_nms sweep (the reference formulation; tests compare)


def _tc_ok(Cin, Cout, stride, R=1, pad=0):
    """Tensor-core kernel coverage: stride 1, or the strided 1x1 convs (TMA element strides)."""
    if CONV_MODE[0] == "simt" or Cin % 32 or Cout % 64:
        return False
    return stride == 1 or (stride == 2 and R == 1 and pad == 0 and TC_STRIDE2[0])


# ---------------------------------------------------------------------------------------------- raw op wrappers
def conv_forward(x, w, scale, bias, residual, res_mode, relu, R, S, stride, pad, out=None, owner=None, out_bf16=None):
    """x: N x H x W x Cin (NHWC contiguous, fp32 - or bf16 inside the bf16 backbone) -> N x Ho x Wo x Cout (bf16 when
    out_bf16; default: the input's dtype)."""
    N, H, W, Cin = x.shape
    Cout = w.shape[-1]
    Ho, Wo = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
    if x.dtype == torch.bfloat16:
        if not (_tc_ok(Cin, Cout, stride, R, pad) and Cin % 64 == 0):
            raise _C.TTDGError(f"no bf16 kernel for this layer (Cin {Cin}, Cout {Cout}, stride {stride}, k {R})")
        out_bf16 = True if out_bf16 is None else out_bf16
        y = torch.empty(N, Ho, Wo, Cout, dtype=torch.bfloat16 if out_bf16 else torch.float32, device=x.device) if out is None else out
        res_bf16 = residual is not None and residual.dtype == torch.bfloat16
        check(_C.lib().ttdg_conv_tc_bf16(_p(x), _p(_weights_bf16(w, owner)), _p(scale), _p(bias), _p(residual), int(res_bf16),
                                         int(res_mode), int(relu), 0, N, H, W, Cin, Cout, R, S, pad, stride, 1, 0, 0, _p(y),
                                         int(y.dtype == torch.bfloat16), _stream()), "conv_tc_bf16")
        return y
    y = torch.empty(N, Ho, Wo, Cout, dtype=torch.float32, device=x.device) if out is None else out
    if _tc_ok(Cin, Cout, stride, R, pad):
        precise = CONV_MODE[0] == "tf32x3"
        w_hi, w_lo = _weights_kmajor(w, True, precise, owner)
        check(_C.lib().ttdg_conv_tc(_p(x), _p(w_hi), _p(w_lo), _p(scale), _p(bias), _p(residual), int(res_mode), int(relu),
                                    0, N, H, W, Cin, Cout, R, S, pad, stride, 1, 0, 0, _p(y), _stream()), "conv_tc")
        return y
    check(_C.lib().ttdg_conv_fwd(_p(x), _p(w), _p(scale), _p(bias), _p(residual), int(res_mode), int(relu), N, H, W, Cin, Cout, R, S,
                                 stride, pad, _p(y), _stream()), "conv_fwd")
    return y


class _ConvFn(torch.autograd.Function):
    """y = relu?(conv(x, w) * scale + bias + residual) with the gradients the TTT loss needs."""

    @staticmethod
    def forward(ctx, x, w, bias_param, residual, scale, bias_const, res_mode, relu, R, S, stride, pad, owner, out_bf16=None):
        bias = bias_param if bias_param is not None else bias_const
        y = conv_forward(x, w, scale, bias, residual, res_mode, relu, R, S, stride, pad, owner=owner, out_bf16=out_bf16)
        ctx.owner = owner
        ctx.save_for_backward(x, w, y if relu else None, scale)
        ctx.meta = (res_mode, relu, R, S, stride, pad, bias_param is not None, residual is not None and residual.requires_grad)
        ctx.res_dtype = residual.dtype if residual is not None else None
        return y

    @staticmethod
    def backward(ctx, g):
        L = _C.lib()
        x, w, y, scale = ctx.saved_tensors
        res_mode, relu, R, S, stride, pad, has_bias, res_grad = ctx.meta
        N, H, W, Cin = x.shape
        Cout = w.shape[-1]
        # bf16 backbone: the backward works on fp32 copies with the single-pass TF32 kernels (activations were saved in bf16:
        # the ReLU mask reads them directly, the weight gradient needs them widened); gradients leave in the dtype autograd
        # expects for each input
        x_dtype = x.dtype
        g = g.float().contiguous() if g.dtype != torch.float32 else g.contiguous()
        s = _stream()
        # ReLU mask (from the stored output) and FrozenBN scale: ONE elementwise pass.  d_pre = g * [y > 0] is only materialised
        # when something consumes it (a residual branch, a bias gradient); d_conv = d_pre * scale feeds dgrad / wgrad.
        need_pre = res_grad or (has_bias and ctx.needs_input_grad[2])
        d_pre = d_conv = None
        if relu and scale is not None and y.dtype == torch.float32:
            d_conv = torch.empty_like(g)
            if need_pre:
                d_pre = torch.empty_like(g)
                check(L.ttdg_relu_bn_bwd2(_p(g), _p(y), _p(scale), Cout, g.numel(), _p(d_pre), _p(d_conv), s), "relu_bn_bwd2")
            else:
                check(L.ttdg_relu_bn_bwd(_p(g), _p(y), _p(scale), Cout, g.numel(), _p(d_conv), s), "relu_bn_bwd")
        elif relu and scale is not None and not need_pre:      # bf16 backbone: mask from the stored bf16 output, same single pass
            d_conv = torch.empty_like(g)
            check(L.ttdg_relu_bn_bwd_bf16y(_p(g), _p(y), _p(scale), Cout, g.numel(), _p(d_conv), s), "relu_bn_bwd_bf16y")
        else:
            if relu:                                       # dPre = g * [y > 0]
                d_pre = torch.empty_like(g)
                if y.dtype == torch.bfloat16:
                    check(L.ttdg_relu_bn_bwd_bf16y(_p(g), _p(y), None, Cout, g.numel(), _p(d_pre), s), "relu_bwd_bf16y")
                else:
                    check(L.ttdg_relu_bn_bwd(_p(g), _p(y), None, Cout, g.numel(), _p(d_pre), s), "relu_bwd")
            else:
                d_pre = g
            if scale is not None:                          # through the FrozenBN scale
                d_conv = torch.empty_like(d_pre)
                check(L.ttdg_relu_bn_bwd(_p(d_pre), None, _p(scale), Cout, d_pre.numel(), _p(d_conv), s), "bn_bwd")
            else:
                d_conv = d_pre
        g_res = None
        if res_grad:
            if res_mode == 1:
                g_res = d_pre
            else:                                          # FPN top-down: sum over the 2x2 children
                Ho, Wo = g.shape[1], g.shape[2]
                g_res = torch.zeros(N, Ho // 2, Wo // 2, Cout, dtype=torch.float32, device=g.device)
                check(L.ttdg_resample2(_p(d_pre), _p(g_res), N, Ho // 2, Wo // 2, Cout, 2, s), "upsample_bwd")
            if ctx.res_dtype != torch.float32:
                g_res = g_res.to(ctx.res_dtype)
        # Weight / bias gradients are ACCUMULATED by their kernels (atomics), so when the parameter already owns a gradient
        # buffer - the views of FlatSGD's flat bucket, zeroed by zero_grad() - they add straight into it and autograd gets
        # None: no zero-filled temporary and no separate "grad += temporary" pass per parameter (132 launches per step).
        def _acc(param):
            gr = getattr(param, "grad", None) if param is not None else None
            return gr if (gr is not None and gr.is_contiguous() and gr.dtype == torch.float32 and gr.data_ptr() % 16 == 0) else None
        owner = ctx.owner
        g_bias = None
        if has_bias and ctx.needs_input_grad[2]:
            acc_b = _acc(owner.bias if owner is not None else None)
            gb = acc_b if acc_b is not None else torch.zeros(Cout, dtype=torch.float32, device=g.device)
            check(L.ttdg_bias_grad(_p(d_pre), d_pre.numel() // Cout, Cout, _p(gb), s), "bias_grad")
            g_bias = None if acc_b is not None else gb
        g_x = g_w = None
        if ctx.needs_input_grad[0]:
            g_x = (torch.zeros if stride == 2 else torch.empty)(N, H, W, Cin, dtype=torch.float32, device=g.device)
            if _tc_ok(Cout, Cin, stride, R, pad):          # GEMM k = Cout, n = Cin; taps mirrored, pad' = R - 1 - pad
                precise = CONV_MODE[0] == "tf32x3"         # (stride 2, 1x1: results land on the even pixels of the zeroed g_x)
                w_hi, w_lo = _weights_kmajor(w, False, precise, ctx.owner)
                check(L.ttdg_conv_tc(_p(d_conv), _p(w_hi), _p(w_lo), None, None, None, 0, 0, 1, N, g.shape[1], g.shape[2], Cout,
                                     Cin, R, S, R - 1 - pad, 1, stride, H, W, _p(g_x), s), "conv_tc_dgrad")
            else:
                check(L.ttdg_conv_dgrad(_p(d_conv), _p(w), N, H, W, Cin, Cout, R, S, stride, pad, _p(g_x), s), "conv_dgrad")
        if g_x is not None and x_dtype != torch.float32:
            g_x = g_x.to(x_dtype)
        if ctx.needs_input_grad[1]:
            if x_dtype != torch.float32:
                x = x.float()
            acc_w = _acc(owner.weight if owner is not None else None)
            if acc_w is not None and acc_w.shape != w.shape:
                acc_w = None
            gw = acc_w if acc_w is not None else torch.zeros_like(w)
            if WGRAD_TC[0] and Cin % 128 == 0 and _tc_ok(Cin, Cout, stride, R, pad):
                check(L.ttdg_wgrad_tc(_p(x), _p(d_conv), int(CONV_MODE[0] == "tf32x3"), N, H, W, Cin, Cout, R, S, stride, pad, _p(gw), s),
                      "wgrad_tc")
            else:
                check(L.ttdg_conv_wgrad(_p(x), _p(d_conv), N, H, W, Cin, Cout, R, S, stride, pad, _p(gw), s), "conv_wgrad")
            g_w = None if acc_w is not None else gw
        return g_x, g_w, g_bias, g_res, None, None, None, None, None, None, None, None, None, None


class FrozenBN(nn.Module):
    """State holder of d2's FrozenBatchNorm2d (buffers only); folded into (scale, bias) by the owning Conv2d."""

    def __init__(self, c):
        super().__init__()
        self.register_buffer("weight", torch.ones(c))
        self.register_buffer("bias", torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.eps = 1e-5


class Conv2d(nn.Module):
    """Convolution / linear layer in the kernels' layout with d2-compatible state-dict I/O.

    kind: 'conv' (torch weight Cout x Cin x R x S), 'linear' (Cout x Cin), 'linear_chw' (Cout x C*H*W flattened in
    (C, H, W) order - the box head's fc1; our activations are (H, W, C)), 'deconv' (ConvTranspose2d Cin x Cout x 2 x 2,
    stored as a 1x1 conv with 4 Cout outputs ordered (a, b, c))."""

    def __init__(self, cin, cout, k=1, stride=1, pad=0, bias=True, norm=False, kind="conv", chw=None, cout_pad=4):
        super().__init__()
        self.cin, self.cout, self.k, self.stride, self.pad, self.kind, self.chw = cin, cout, k, stride, pad, kind, chw
        # cout_pad = 64 puts the narrow prediction heads (15 / 60 / 3 / 8 / 2 outputs) on the tensor-core kernel, whose N
        # tile is 64: the zero columns cost nothing next to a 128-wide CUDA-core tile that is 90 % padding
        self.cin_p, self.cout_p = _pad4(cin), (cout + cout_pad - 1) // cout_pad * cout_pad
        if kind == "deconv":
            self.weight = nn.Parameter(torch.zeros(1, 1, self.cin_p, 4 * self.cout_p))
            self.bias = nn.Parameter(torch.zeros(4 * self.cout_p))
        else:
            self.weight = nn.Parameter(torch.zeros(k, k, self.cin_p, self.cout_p))
            self.bias = nn.Parameter(torch.zeros(self.cout_p)) if bias else None
        self.norm = FrozenBN(cout) if norm else None
        self.register_buffer("fold_scale", None, persistent=False)
        self.register_buffer("fold_bias", None, persistent=False)

    # ---- state dict in Detectron2's names / shapes
    def _save_to_state_dict(self, destination, prefix, keep_vars):
        w = self.weight.detach()
        if self.kind == "conv":
            t = w[:, :, :self.cin, :self.cout].permute(3, 2, 0, 1).contiguous()
        elif self.kind == "linear":
            t = w[0, 0, :self.cin, :self.cout].t().contiguous()
        elif self.kind == "linear_chw":
            C, H, W = self.chw
            t = w[0, 0, :, :self.cout].reshape(H, W, C, self.cout).permute(3, 2, 0, 1).reshape(self.cout, C * H * W).contiguous()
        else:   # deconv: stored [ci][(a, b, co)] -> torch [ci][co][a][b]
            t = w[0, 0, :self.cin].reshape(self.cin, 2, 2, self.cout_p)[..., :self.cout].permute(0, 3, 1, 2).contiguous()
        destination[prefix + "weight"] = t
        if self.bias is not None:
            b = self.bias.detach()
            destination[prefix + "bias"] = (b.reshape(4, self.cout_p)[0, :self.cout] if self.kind == "deconv" else b[:self.cout]).clone()
        if self.norm is not None:
            for k in ("weight", "bias", "running_mean", "running_var"):
                destination[prefix + "norm." + k] = getattr(self.norm, k).clone()

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        key = prefix + "weight"
        if key not in state_dict:
            missing_keys.append(key)
            return
        t = state_dict[key].to(torch.float32)
        with torch.no_grad():
            w = torch.zeros_like(self.weight)
            if self.kind == "conv":
                w[:, :, :self.cin, :self.cout] = t.permute(2, 3, 1, 0)
            elif self.kind == "linear":
                w[0, 0, :self.cin, :self.cout] = t.t()
            elif self.kind == "linear_chw":
                C, H, W = self.chw
                w[0, 0, :, :self.cout] = t.reshape(self.cout, C, H, W).permute(2, 3, 1, 0).reshape(H * W * C, self.cout)
            else:
                w[0, 0, :self.cin] = torch.nn.functional.pad(t.permute(0, 2, 3, 1), (0, self.cout_p - self.cout)).reshape(self.cin, 4 * self.cout_p)
            self.weight.copy_(w)
            if self.bias is not None:
                if prefix + "bias" not in state_dict:
                    missing_keys.append(prefix + "bias")
                else:
                    b = torch.zeros_like(self.bias)
                    src = state_dict[prefix + "bias"].to(torch.float32)
                    if self.kind == "deconv":
                        b.view(4, self.cout_p)[:, :self.cout] = src
                    else:
                        b[:self.cout] = src
                    self.bias.copy_(b)
            if self.norm is not None:
                for k in ("weight", "bias", "running_mean", "running_var"):
                    if prefix + "norm." + k not in state_dict:
                        missing_keys.append(prefix + "norm." + k)
                    else:
                        getattr(self.norm, k).copy_(state_dict[prefix + "norm." + k])
                self.fold()

    def fold(self):
        """FrozenBatchNorm2d: scale = w * rsqrt(var + eps), bias = b - mean * scale."""
        n = self.norm
        scale = n.weight * (n.running_var + n.eps).rsqrt()
        bias = n.bias - n.running_mean * scale
        self.fold_scale = torch.nn.functional.pad(scale, (0, self.cout_p - self.cout)).contiguous()
        self.fold_bias = torch.nn.functional.pad(bias, (0, self.cout_p - self.cout)).contiguous()

    def forward(self, x, relu=False, residual=None, res_mode=0, out_bf16=None):
        if self.norm is not None and self.fold_scale is None:
            self.fold()
        scale = self.fold_scale if self.norm is not None else None
        bias_const = self.fold_bias if self.norm is not None else None
        k = 1 if self.kind != "conv" else self.k
        if torch.is_grad_enabled() and (x.requires_grad or self.weight.requires_grad):
            return _ConvFn.apply(x, self.weight, self.bias, residual, scale, bias_const, res_mode, relu, k, k, self.stride, self.pad, self,
                                 out_bf16)
        bias = self.bias if self.bias is not None else bias_const
        return conv_forward(x, self.weight, scale, None if bias is None else bias.detach(), residual, res_mode, relu, k, k,
                            self.stride, self.pad, owner=self, out_bf16=out_bf16)


class Bottleneck(nn.Module):
    def __init__(self, cin, mid, cout, stride, shortcut):
        super().__init__()
        self.shortcut = Conv2d(cin, cout, 1, stride, 0, bias=False, norm=True) if shortcut else None
        self.conv1 = Conv2d(cin, mid, 1, stride, 0, bias=False, norm=True)          # STRIDE_IN_1X1
        self.conv2 = Conv2d(mid, mid, 3, 1, 1, bias=False, norm=True)
        self.conv3 = Conv2d(mid, cout, 1, 1, 0, bias=False, norm=True)

    def forward(self, x):
        out = self.conv1(x, relu=True)
        out = self.conv2(out, relu=True)
        sc = self.shortcut(x) if self.shortcut is not None else x
        return self.conv3(out, relu=True, residual=sc, res_mode=1)                   # relu(bn(conv3) + shortcut)


GRAD_READY_HOOK = [None]      # callable(k): the gradients of bucket k are complete (FlatSGD.grad_ready), see ResNet50.forward
STEM_TC = [os.environ.get("TTDG_STEM_TC", "1") == "1"]        # stem on tensor cores (padded image, overlapping TMA windows)
STEM_LEFT, STEM_EXTRA = 3, 8                                   # padded image rows: 3 zero pixels | image | 5 zero pixels


class Stem(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = Conv2d(3, 64, 7, 2, 3, bias=False, norm=True)

    def _weights_tc(self, precise):
        """[7][64][32] K-major stem weights: wk[r][co][4 s + c] = w[r][s][c][co], zeros for the 8th pixel of the window."""
        w = self.conv1.weight
        stamp = (PARAM_EPOCH[0], w._version, w.data_ptr(), precise)
        hit = self.__dict__.get("_wk_stem")
        if hit is not None and hit[0] == stamp:
            return hit[1], hit[2]
        wk = torch.zeros(7, 64, 32, dtype=torch.float32, device=w.device)
        wk[:, :, :28] = w.detach().permute(0, 3, 1, 2).reshape(7, 64, 28)
        hi, lo = tf32_split(wk)
        if not precise:
            lo = None
        self.__dict__["_wk_stem"] = (stamp, hi, lo)
        return hi, lo

    def forward(self, x, width=None):
        """x: N x H x W x 4 (plain) or, with ``width``, the padded image N x H x (width + 8) x 4 for the tensor-core stem."""
        c = self.conv1
        if width is None:
            return c(x, relu=True)
        if c.fold_scale is None:
            c.fold()
        N, H, Wp, _ = x.shape
        hi, lo = self._weights_tc(CONV_MODE[0] == "tf32x3")
        bf16 = CONV_MODE[0] == "bf16"                      # the image is fp32 (3 channels); the first ACTIVATION is bf16
        y = torch.empty(N, H // 2, width // 2, 64, dtype=torch.bfloat16 if bf16 else torch.float32, device=x.device)
        check(_C.lib().ttdg_stem_tc2(_p(x), Wp, _p(hi), _p(lo), _p(c.fold_scale), _p(c.fold_bias), 1, N, H, width, _p(y), int(bf16),
                                     _stream()), "stem_tc")
        return y


class ResNet50(nn.Module):
    def __init__(self):
        super().__init__()
        self.stem = Stem()
        cin = 64
        for name, blocks, mid, cout, stride in RES_STAGES:
            layers = []
            for b in range(blocks):
                layers.append(Bottleneck(cin, mid, cout, stride if b == 0 else 1, b == 0))
                cin = cout
            setattr(self, name, nn.Sequential(*layers))

    def forward(self, x, width=None):
        L = _C.lib()
        with torch.no_grad():                                   # FREEZE_AT = 2: stem + res2
            y = self.stem(x, width)
            N, H, W, C = y.shape
            p = torch.empty(N, (H - 1) // 2 + 1, (W - 1) // 2 + 1, C, dtype=y.dtype, device=y.device)
            if y.dtype == torch.bfloat16:
                check(L.ttdg_maxpool3x3s2_bf16(_p(y), N, H, W, C, _p(p), _stream()), "maxpool_bf16")
            else:
                check(L.ttdg_maxpool3x3s2(_p(y), N, H, W, C, _p(p), _stream()), "maxpool")
            y = self.res2(p)
        out = {"res2": y}
        for k, name in enumerate(("res3", "res4", "res5")):
            y = getattr(self, name)(y)
            out[name] = y
            # Overlapped gradient all-reduce: when the gradient w.r.t. a stage's OUTPUT is complete, every parameter above it
            # has its gradient (weight gradients are produced by the same autograd nodes as the data gradients).  With
            # adapted_parameters() ordered [affinity, FPN, res5 | res4 | res3]: d(res4 out) -> bucket 0, d(res3 out) -> bucket 1.
            if GRAD_READY_HOOK[0] is not None and name != "res5" and y.requires_grad:
                bucket = 1 - k
                y.register_hook(lambda g, b=bucket: (GRAD_READY_HOOK[0](b), None)[1])
        return out


class _Subsample2(torch.autograd.Function):
    """p6 = max_pool2d(p5, kernel_size=1, stride=2) (d2 LastLevelMaxPool)."""

    @staticmethod
    def forward(ctx, x):
        N, H, W, C = x.shape
        y = torch.empty(N, H // 2, W // 2, C, dtype=torch.float32, device=x.device)
        check(_C.lib().ttdg_resample2(_p(x), _p(y), N, H // 2, W // 2, C, 0, _stream()), "subsample")
        ctx.shape = (N, H, W, C)
        return y

    @staticmethod
    def backward(ctx, g):
        N, H, W, C = ctx.shape
        gx = torch.zeros(N, H, W, C, dtype=torch.float32, device=g.device)
        check(_C.lib().ttdg_resample2(_p(g.contiguous()), _p(gx), N, H // 2, W // 2, C, 1, _stream()), "subsample_bwd")
        return gx


class Backbone(nn.Module):
    """build_resnet_fpn_backbone: ResNet-50 bottom-up + FPN (sum fusion, no norm) + LastLevelMaxPool."""

    def __init__(self):
        super().__init__()
        self.bottom_up = ResNet50()
        for lvl, c in zip((2, 3, 4, 5), (256, 512, 1024, 2048)):
            setattr(self, f"fpn_lateral{lvl}", Conv2d(c, 256, 1, 1, 0, bias=True))
            setattr(self, f"fpn_output{lvl}", Conv2d(256, 256, 3, 1, 1, bias=True))

    out_features = ("p2", "p3", "p4", "p5", "p6")

    def pyramid(self, x, width=None):
        """[p2 .. p6] as N x H x W x 256 (NHWC) maps.  x: the preprocessed image tensor (``preprocess``); width: the image
        width when x carries the stem's zero padding (also read from the ``stem_width`` attribute ``preprocess`` sets)."""
        if width is None:
            width = getattr(x, "stem_width", None)
        res = self.bottom_up(x, width)
        outs, prev = {}, None
        for lvl in (5, 4, 3, 2):
            lat = getattr(self, f"fpn_lateral{lvl}")
            prev = lat(res[f"res{lvl}"]) if prev is None else lat(res[f"res{lvl}"], residual=prev, res_mode=2)
            outs[lvl] = getattr(self, f"fpn_output{lvl}")(prev, out_bf16=False)      # the pyramid is fp32 in every mode
        return [outs[2], outs[3], outs[4], outs[5], _Subsample2.apply(outs[5])]

    def forward(self, x, width=None):
        """d2 Backbone.forward (called as ``self.backbone(images.tensor)``, rcnn.py:226): {"p2": .., "p6": ..} in NCHW - views
        of the NHWC maps the kernels write, so turning them back (``_nhwc``) costs nothing."""
        return dict(zip(self.out_features, (f.permute(0, 3, 1, 2) for f in self.pyramid(x, width))))


class RPNHead(nn.Module):
    def __init__(self, num_anchors=15):
        super().__init__()
        self.conv = Conv2d(256, 256, 3, 1, 1, bias=True)
        self.objectness_logits = Conv2d(256, num_anchors, 1, 1, 0, bias=True, cout_pad=64)
        self.anchor_deltas = Conv2d(256, num_anchors * 4, 1, 1, 0, bias=True, cout_pad=64)


def cell_anchors():
    a = []
    for s in ANCHOR_SIZES:
        area = float(s) ** 2
        for r in ANCHOR_RATIOS:
            w = math.sqrt(area / r)
            h = r * w
            a.append([-w / 2.0, -h / 2.0, w / 2.0, h / 2.0])
    return torch.tensor(a, dtype=torch.float32)


_NMS_SCRATCH = {}


def _nhwc(f):
    """NCHW view of an NHWC map (what Backbone.forward hands out) -> the NHWC map; a real NCHW tensor is re-laid out."""
    return f.permute(0, 2, 3, 1).contiguous()


def _sizes(image_sizes, n):
    """(h, w) for all images, or one (h, w) per image (d2 ImageList.image_sizes) -> (list of n sizes, all equal?)."""
    if len(image_sizes) == 2 and not isinstance(image_sizes[0], (tuple, list)):
        return [tuple(image_sizes)] * n, True
    sizes = [tuple(int(v) for v in s) for s in image_sizes]
    assert len(sizes) == n
    return sizes, all(s == sizes[0] for s in sizes)


def nms_sorted(boxes, cats, thresh, max_keep):
    """Per-category NMS over boxes sorted by descending score, for a batch of images in one launch pair.
    boxes: B x n x 4 (or n x 4), cats: B x n int32.  Returns (keep int32 [B][max_keep], n_keep int32 [B])."""
    single = boxes.dim() == 2
    if single:
        boxes, cats = boxes.unsqueeze(0), cats.unsqueeze(0)
    boxes, cats = boxes.contiguous(), cats.contiguous()
    B, n = boxes.shape[0], boxes.shape[1]
    L = _C.lib()
    keep = torch.zeros(B, max_keep, dtype=torch.int32, device=boxes.device)
    n_keep = torch.zeros(B, dtype=torch.int32, device=boxes.device)
    need = max(L.ttdg_nms_scratch_bytes(B, n), 8)
    scratch = _NMS_SCRATCH.get(boxes.device)
    if scratch is None or scratch.numel() < need:       # persistent, grow-only: an 80 MB block that comes and goes makes
        scratch = torch.empty(need, dtype=torch.uint8, device=boxes.device)     # the caching allocator cudaMalloc/cudaFree every step
        _NMS_SCRATCH[boxes.device] = scratch
    check(L.ttdg_nms(_p(boxes), _p(cats), B, n, float(thresh), int(max_keep), _p(keep), _p(n_keep), _p(scratch), _stream()), "nms")
    return (keep[0], n_keep) if single else (keep, n_keep)


class RPN(nn.Module):
    """PseudoLabRPN (proposal_generator/rpn.py:16-55) with compute_loss=False: StandardRPNHead + d2
    find_top_rpn_proposals.  pre-NMS top-k per level is 2000 in train mode (the TTT pass) / 1000 in eval mode."""

    def __init__(self):
        super().__init__()
        self.rpn_head = RPNHead(len(ANCHOR_SIZES) * len(ANCHOR_RATIOS))
        self.nms_thresh, self.post_topk = 0.7, 1000
        self._cell = cell_anchors()

    in_features = ("p2", "p3", "p4", "p5", "p6")

    def forward(self, feats, image_size, training):
        return self.predict(feats, image_size, training)

    @torch.no_grad()
    def predict_padded(self, feats, image_size, training):
        """feats: [p2 .. p6] NHWC; image_size: (h, w) or one (h, w) per image (boxes are clipped to the image's own size, d2
        find_top_rpn_proposals).  Everything after the head convolutions runs in five launches without touching the host
        (csrc/select.cu): top-k + sort + decode for all (image, level) pairs, per-image ordering across levels, per-level NMS
        (2 launches), compaction.  Returns (boxes N x 1000 x 4, objectness logits N x 1000 - padding rows hold -inf -, counts
        N int32 on the device)."""
        L = _C.lib()
        A = self._cell.shape[0]
        pre_topk = 2000 if training else 1000
        N = feats[0].shape[0]
        dev = feats[0].device
        sizes, _ = _sizes(image_size, N)
        logits_l, deltas_l, hw, ks = [], [], [], []
        for l, f in enumerate(feats):
            f = f.detach()
            _, H, W, _ = f.shape
            if CONV_MODE[0] == "bf16":                                  # configs[2]: the heads' GEMMs run kind::f16 as well
                f = f.to(torch.bfloat16)
            t = self.rpn_head.conv(f, relu=True)
            logits_l.append(self.rpn_head.objectness_logits(t, out_bf16=False))     # N x H x W x 64 (15 used), fp32
            deltas_l.append(self.rpn_head.anchor_deltas(t, out_bf16=False))         # N x H x W x 64 (60 used)
            hw += [H, W]
            ks.append(min(H * W * A, pre_topk))
        nl, Kt = len(feats), sum(ks)
        boxes = torch.empty(N, Kt, 4, dtype=torch.float32, device=dev)
        scores = torch.empty(N, Kt, dtype=torch.float32, device=dev)
        valid = torch.empty(N, Kt, dtype=torch.uint8, device=dev)
        vp = lambda ts: ctypes.cast((ctypes.c_void_p * nl)(*[t.data_ptr() for t in ts]), ctypes.c_void_p)
        ip = lambda vs: ctypes.cast((ctypes.c_int32 * len(vs))(*vs), ctypes.c_void_p)
        fp = lambda vs: ctypes.cast((ctypes.c_float * len(vs))(*vs), ctypes.c_void_p)
        check(L.ttdg_rpn_select(vp(logits_l), vp(deltas_l), ip(hw), ip(list(STRIDES[:nl])), ip(ks), nl, logits_l[0].shape[-1],
                                deltas_l[0].shape[-1], A, fp(self._cell.reshape(-1).tolist()), N,
                                fp([float(v) for s_ in sizes for v in s_]), _p(boxes), _p(scores), _p(valid), _stream()), "rpn_select")
        if RPN_NMS_PER_LEVEL[0]:
            # batched_nms with the level as category = independent problems of <= 2000 boxes per (image, level), each in shared
            # memory (40 CTAs) instead of one 8960-box sweep per image; then one ordering pass picks the first post_topk kept
            kept = torch.empty(N, Kt, dtype=torch.uint8, device=dev)
            check(L.ttdg_rpn_nms_levels(_p(boxes), _p(valid), ip(ks), nl, N, float(self.nms_thresh), int(self.post_topk), _p(kept),
                                        _stream()), "rpn_nms_levels")
            out_b = torch.empty(N, self.post_topk, 4, dtype=torch.float32, device=dev)
            out_s = torch.empty(N, self.post_topk, dtype=torch.float32, device=dev)
            counts = torch.empty(N, dtype=torch.int32, device=dev)
            check(L.ttdg_top_candidates(_p(boxes), _p(scores), _p(kept), N, Kt, int(self.post_topk), -float("inf"), _p(out_b), _p(out_s),
                                        _p(counts), _stream()), "top_candidates")
            return out_b, out_s, counts
        key = (tuple(ks), dev)
        lvl = self.__dict__.setdefault("_lvl_cache", {}).get(key)
        if lvl is None:                                                 # level of every candidate position: the NMS category
            lvl = torch.cat([torch.full((k,), l, dtype=torch.int32) for l, k in enumerate(ks)]).to(dev)
            self._lvl_cache[key] = lvl
        b_sorted, s_sorted, cats, n_valid = sort_candidates(boxes, scores, valid, lvl, 0, -float("inf"))
        keep, n_keep = nms_sorted(b_sorted, cats, self.nms_thresh, self.post_topk)          # all images, one launch pair
        out_b, out_s, _, counts = gather_kept(b_sorted, s_sorted, cats, keep, n_keep, n_valid, self.post_topk, -float("inf"), False)
        return out_b, out_s, counts

    @torch.no_grad()
    def predict(self, feats, image_size, training):
        """List form of predict_padded: per image (proposal boxes k x 4, objectness logits k), sorted by logit; one host read."""
        out_b, out_s, counts = self.predict_padded(feats, image_size, training)
        cnt = counts.cpu().tolist()
        return [(out_b[n, :c], out_s[n, :c]) for n, c in enumerate(cnt)]


def sort_candidates(boxes, scores, valid, cats_in, cat_mod, invalid_score):
    """boxes B x n x 4, scores B x n (+ valid B x n uint8 or None: valid <=> score > 0) -> candidates of every image ordered by
    score (invalid last) with the NMS category: cats_in[position] or position % cat_mod."""
    B, n = scores.shape
    dev = scores.device
    b_out, s_out = torch.empty_like(boxes), torch.empty_like(scores)
    cats = torch.empty(B, n, dtype=torch.int32, device=dev)
    n_valid = torch.empty(B, dtype=torch.int32, device=dev)
    check(_C.lib().ttdg_sort_candidates(_p(boxes), _p(scores), _p(valid), _p(cats_in), int(cat_mod), B, n, float(invalid_score), _p(b_out),
                                        _p(s_out), _p(cats), _p(n_valid), _stream()), "sort_candidates")
    return b_out, s_out, cats, n_valid


def gather_kept(b_sorted, s_sorted, cats, keep, n_keep, n_valid, max_keep, pad_score, want_cats=True):
    B, n = s_sorted.shape
    dev = s_sorted.device
    out_b = torch.empty(B, max_keep, 4, dtype=torch.float32, device=dev)
    out_s = torch.empty(B, max_keep, dtype=torch.float32, device=dev)
    out_c = torch.empty(B, max_keep, dtype=torch.int64, device=dev) if want_cats else None
    counts = torch.empty(B, dtype=torch.int32, device=dev)
    check(_C.lib().ttdg_gather_kept(_p(b_sorted), _p(s_sorted), _p(cats), _p(keep), _p(n_keep), _p(n_valid), B, n, int(max_keep),
                                    float(pad_score), _p(out_b), _p(out_s), _p(out_c), _p(counts), _stream()), "gather_kept")
    return out_b, out_s, out_c, counts


class BoxHead(nn.Module):
    def __init__(self):
        super().__init__()
        self.fc1 = Conv2d(256 * 7 * 7, 1024, kind="linear_chw", chw=(256, 7, 7))
        self.fc2 = Conv2d(1024, 1024, kind="linear")


class BoxPredictor(nn.Module):
    def __init__(self, num_classes):
        super().__init__()
        self.cls_score = Conv2d(1024, num_classes + 1, kind="linear", cout_pad=64)
        self.bbox_pred = Conv2d(1024, num_classes * 4, kind="linear", cout_pad=64)


class MaskHead(nn.Module):
    def __init__(self, num_classes):
        super().__init__()
        for i in range(1, 5):
            setattr(self, f"mask_fcn{i}", Conv2d(256, 256, 3, 1, 1, bias=True))
        self.deconv = Conv2d(256, 256, kind="deconv")
        self.predictor = Conv2d(256, num_classes, 1, 1, 0, bias=True, cout_pad=64)


def roi_align(feats4, rois, pooled):
    L = _C.lib()
    n = rois.shape[0]
    C = feats4[0].shape[-1]
    out = torch.empty(n, pooled, pooled, C, dtype=torch.float32, device=rois.device)
    ptrs = (ctypes.c_void_p * 4)(*[f.data_ptr() for f in feats4])
    hw = (ctypes.c_int32 * 8)(*[v for f in feats4 for v in (f.shape[1], f.shape[2])])
    check(L.ttdg_roi_align(ctypes.cast(ptrs, ctypes.c_void_p), ctypes.cast(hw, ctypes.c_void_p), _p(rois), n, C, pooled, _p(out),
                           _stream()), "roi_align")
    return out


def _rois(boxes_per_image):
    dev = boxes_per_image[0].device
    counts = [len(b) for b in boxes_per_image]
    if len(set(counts)) == 1:                               # the usual case: same count per image, no host -> device copy
        idx = torch.arange(len(counts), dtype=torch.float32, device=dev).repeat_interleave(counts[0])
    else:
        idx = torch.repeat_interleave(torch.arange(len(counts), dtype=torch.float32), torch.tensor(counts)).to(dev)
    return torch.cat([idx.unsqueeze(1), torch.cat(boxes_per_image)], dim=1).contiguous()


class ROIHeads(nn.Module):
    """StandardROIHeadsPseudoLab (roi_heads/roi_heads.py:65-114, 173-205) in inference form: box branch, and for
    every branch except 'TTT' the mask branch (forward_with_given_boxes)."""

    def __init__(self, num_classes=2):
        super().__init__()
        self.num_classes = num_classes
        self.box_head = BoxHead()
        self.box_predictor = BoxPredictor(num_classes)
        self.mask_head = MaskHead(num_classes)
        self.score_thresh, self.nms_thresh, self.topk = 0.05, 0.5, 100

    @torch.no_grad()
    def forward_box_padded(self, feats, pboxes, pcounts, image_size):
        """_forward_box + FastRCNNOutputLayers.inference on PADDED proposals (N x P x 4 with per-image counts on the device):
        RoIAlign, the two FCs, class-specific decoding, score threshold, per-class NMS, top 100 - without a host round trip.
        Returns (boxes N x 100 x 4, scores N x 100, classes N x 100 int64, counts N int32); padding rows: score 0, class -1."""
        L = _C.lib()
        K = self.num_classes
        feats4 = [f.detach() for f in feats[:4]]
        dev = feats4[0].device
        N, P = pboxes.shape[0], pboxes.shape[1]
        R = N * P
        pb = pboxes.reshape(R, 4).contiguous()
        rois = torch.empty(R, 5, dtype=torch.float32, device=dev)
        check(L.ttdg_rois_from_padded(_p(pb), _p(pcounts), N, P, _p(rois), _stream()), "rois_from_padded")
        x = roi_align(feats4, rois, 7).reshape(R, 1, 1, 7 * 7 * 256)
        if CONV_MODE[0] == "bf16":
            x = x.to(torch.bfloat16)
        x = self.box_head.fc1(x, relu=True)
        x = self.box_head.fc2(x, relu=True)
        cls = self.box_predictor.cls_score(x, out_bf16=False).reshape(R, -1)
        reg = self.box_predictor.bbox_pred(x, out_bf16=False).reshape(R, -1)
        cand_b = torch.empty(R * K, 4, dtype=torch.float32, device=dev)
        cand_s = torch.empty(R * K, dtype=torch.float32, device=dev)
        sizes, same = _sizes(image_size, N)
        if same:
            check(L.ttdg_box_predict(_p(cls), cls.shape[1], _p(reg), reg.shape[1], _p(pb), R, K, float(sizes[0][0]), float(sizes[0][1]),
                                     self.score_thresh, _p(cand_b), _p(cand_s), _stream()), "box_predict")
        else:                                                           # clip to each image's own size
            for i in range(N):
                o = i * P
                check(L.ttdg_box_predict(_p(cls[o:o + P]), cls.shape[1], _p(reg[o:o + P]), reg.shape[1], _p(pb[o:o + P]), P, K,
                                         float(sizes[i][0]), float(sizes[i][1]), self.score_thresh, _p(cand_b[o * K:(o + P) * K]),
                                         _p(cand_s[o * K:(o + P) * K]), _stream()), "box_predict")
        check(L.ttdg_mask_padded_candidates(_p(cand_s), _p(pcounts), N, P, K, _stream()), "mask_padded")
        # candidate t = roi * K + class: the NMS category is t % K
        bb, ss, cats, n_valid = sort_candidates(cand_b.view(N, P * K, 4), cand_s.view(N, P * K), None, None, K, -1.0)
        keep, n_keep = nms_sorted(bb, cats, self.nms_thresh, self.topk)
        return gather_kept(bb, ss, cats, keep, n_keep, n_valid, self.topk, 0.0, True)

    @torch.no_grad()
    def forward_box(self, feats, proposals, image_size):
        """List form: proposals = per image (boxes k x 4, logits k) -> per image (boxes, scores, classes); one host read."""
        dev = feats[0].device
        lens = [len(p[0]) for p in proposals]
        P = max(max(lens), 1)
        if all(n == P for n in lens):
            pboxes = torch.stack([p[0] for p in proposals])
        else:
            pboxes = torch.zeros(len(proposals), P, 4, dtype=torch.float32, device=dev)
            for i, p in enumerate(proposals):
                pboxes[i, :lens[i]] = p[0]
        pcounts = torch.tensor(lens, dtype=torch.int32).to(dev, non_blocking=True)
        b, s_, c, counts = self.forward_box_padded(feats, pboxes, pcounts, image_size)
        cnt = counts.cpu().tolist()
        return [(b[i, :n], s_[i, :n], c[i, :n]) for i, n in enumerate(cnt)]

    @torch.no_grad()
    def mask_logits(self, feats, dets):
        """Mask branch up to the predictor (d2 _forward_mask in inference form): R x 28 x 28 x Kpad logits (K used) for the
        R = sum of detections of the batch, or None when there are none."""
        L = _C.lib()
        feats4 = [f.detach() for f in feats[:4]]
        boxes = [d[0] for d in dets]
        R = sum(len(b) for b in boxes)
        if R == 0:
            return None
        x = roi_align(feats4, _rois(boxes), 14)
        if CONV_MODE[0] == "bf16":
            x = x.to(torch.bfloat16)
        for i in range(1, 5):
            x = getattr(self.mask_head, f"mask_fcn{i}")(x, relu=True)
        y4 = self.mask_head.deconv(x, relu=True, out_bf16=False)     # R x 14 x 14 x (4 * 256), fp32 for the shuffle / predictor
        y = torch.empty(R, 28, 28, 256, dtype=torch.float32, device=y4.device)
        check(L.ttdg_pixel_shuffle2(_p(y4), R, 14, 14, 256, _p(y), _stream()), "pixel_shuffle")
        return self.mask_head.predictor(y)                           # R x 28 x 28 x 64 (K used)

    @torch.no_grad()
    def paste(self, logits, dets, out_size, image_size):
        """d2 detector_postprocess + paste_masks_in_image for the whole batch: boxes rescaled from the network input size to
        the requested output size, clipped, empty ones dropped, the class's 28 x 28 logits pasted (sigmoid, bilinear, >= 0.5).
        out_size / image_size: (h, w) for all images or one per image.  Same sizes everywhere = one paste launch and one host
        sync for the batch; per-image results are views."""
        L = _C.lib()
        B = len(dets)
        dev = dets[0][0].device
        counts = [len(d[0]) for d in dets]
        R = sum(counts)
        outs, same_o = _sizes(out_size, B)
        ins, same_i = _sizes(image_size, B)
        bs = torch.cat([d[0] for d in dets]) if R else torch.zeros(0, 4, dtype=torch.float32, device=dev)
        scores = torch.cat([d[1] for d in dets]) if R else torch.zeros(0, dtype=torch.float32, device=dev)
        classes = (torch.cat([d[2] for d in dets]) if R else torch.zeros(0, dtype=torch.int64, device=dev)).contiguous()
        if same_o and same_i:
            H, W = outs[0]
            sx, sy = W / ins[0][1], H / ins[0][0]
            if sx != 1.0 or sy != 1.0:
                bs = torch.stack((bs[:, 0] * sx, bs[:, 1] * sy, bs[:, 2] * sx, bs[:, 3] * sy), dim=1)
            bs = torch.stack((bs[:, 0].clamp(0, W), bs[:, 1].clamp(0, H), bs[:, 2].clamp(0, W), bs[:, 3].clamp(0, H)), dim=1).contiguous()
            masks = torch.empty(R, H, W, dtype=torch.uint8, device=dev)     # the paste kernel writes every pixel
            if R:
                check(L.ttdg_mask_paste(_p(logits), logits.shape[-1], 28, _p(bs), _p(classes), R, H, W, 0.5, _p(masks), _stream()), "mask_paste")
            per_image = [(bs[o:o + n], masks[o:o + n]) for o, n in zip(_offsets(counts), counts)]
        else:                                                               # every image pasted at its own size
            per_image = []
            for i, (o, n) in enumerate(zip(_offsets(counts), counts)):
                H, W = outs[i]
                sx, sy = W / ins[i][1], H / ins[i][0]
                b = bs[o:o + n]
                b = torch.stack(((b[:, 0] * sx).clamp(0, W), (b[:, 1] * sy).clamp(0, H), (b[:, 2] * sx).clamp(0, W),
                                 (b[:, 3] * sy).clamp(0, H)), dim=1).contiguous()
                mk = torch.empty(n, H, W, dtype=torch.uint8, device=dev)
                if n:
                    check(L.ttdg_mask_paste(_p(logits[o:o + n]), logits.shape[-1], 28, _p(b), _p(classes[o:o + n]), n, H, W, 0.5, _p(mk),
                                            _stream()), "mask_paste")
                per_image.append((b, mk))
        all_kept = True
        if R:
            allb = torch.cat([b for b, _ in per_image])
            keep = ((allb[:, 2] - allb[:, 0]) > 0) & ((allb[:, 3] - allb[:, 1]) > 0)        # Boxes.nonempty()
            all_kept = bool(keep.all().item())
        results = []
        for (b, mk), o, n in zip(per_image, _offsets(counts), counts):
            sl = slice(o, o + n)
            mb = mk.view(torch.bool)                                        # the paste kernel writes 0 / 1
            if all_kept:
                results.append({"pred_boxes": b, "scores": scores[sl], "pred_classes": classes[sl], "pred_masks": mb})
            else:
                k = keep[sl]
                results.append({"pred_boxes": b[k], "scores": scores[sl][k], "pred_classes": classes[sl][k], "pred_masks": mb[k]})
        return results

    @torch.no_grad()
    def forward_mask(self, feats, dets, out_size, image_size):
        """Mask branch + detector_postprocess: returns per image a dict with pred_boxes / scores / pred_classes /
        pred_masks (bool R x H x W)."""
        return self.paste(self.mask_logits(feats, dets), dets, out_size, image_size)


def _offsets(counts):
    o, out = 0, []
    for n in counts:
        out.append(o)
        o += n
    return out


def preprocess(images_u8, device, stem_padded=False):
    """d2 preprocess_image + ImageList.from_tensors: list of uint8 3 x H_i x W_i -> N x Hp x Wp x 4 fp32 NHWC, mean-subtracted
    (std 1), every image in the top-left corner of a zero canvas of the batch's maximum size rounded up to a multiple of 32
    (size_divisibility; d2 pads AFTER normalisation, so the padding is 0).  stem_padded: rows carry the stem's zero padding as
    well (3 pixels left, 5 right), the layout ttdg_stem_tc reads; returns (tensor N x Hp x (Wp + 8) x 4, Wp)."""
    images_u8 = list(images_u8)
    shapes = [tuple(im.shape) for im in images_u8]
    assert all(len(sh) == 3 and sh[0] == 3 for sh in shapes) and all(im.dtype == torch.uint8 for im in images_u8)
    same = all(sh == shapes[0] for sh in shapes)
    Hm, Wm = max(sh[1] for sh in shapes), max(sh[2] for sh in shapes)
    Hp, Wq = (Hm + 31) // 32 * 32, (Wm + 31) // 32 * 32
    left, Wp = (STEM_LEFT, Wq + STEM_EXTRA) if stem_padded else (0, Wm)
    N = len(images_u8)
    L = _C.lib()
    if same:
        if images_u8[0].device.type == "cpu":               # host images (pinned by the loader): one async H2D copy each, straight
            x = torch.empty((N,) + shapes[0], dtype=torch.uint8, device=device)                                     # into the batch
            for n, im in enumerate(images_u8):
                x[n].copy_(im, non_blocking=True)
        else:
            x = torch.stack(images_u8).contiguous()
        H, W = shapes[0][1:]
        if Hp != H:                                         # bottom padding: image by image into the taller zeroed buffer
            out = torch.zeros(N, Hp, Wp, 4, dtype=torch.float32, device=device)
            for n in range(N):
                check(L.ttdg_preprocess(_p(x[n]), 1, H, W, Wp, left, *PIXEL_MEAN, _p(out[n]), _stream()), "preprocess")
        else:
            out = torch.empty(N, H, Wp, 4, dtype=torch.float32, device=device)
            check(L.ttdg_preprocess(_p(x), N, H, W, Wp, left, *PIXEL_MEAN, _p(out), _stream()), "preprocess")
    else:                                                   # mixed sizes: every image into its corner of the zero canvas
        out = torch.zeros(N, Hp, Wp, 4, dtype=torch.float32, device=device)
        for n, im in enumerate(images_u8):
            xi = im.to(device, non_blocking=True).contiguous()
            check(L.ttdg_preprocess(_p(xi), 1, im.shape[1], im.shape[2], Wp, left, *PIXEL_MEAN, _p(out[n]), _stream()), "preprocess")
    if stem_padded:
        out.stem_width = Wq
        return out, Wq
    if Wq != Wm:
        out = torch.nn.functional.pad(out, (0, 0, 0, Wq - Wm)).contiguous()
    return out


class MaskRCNN(nn.Module):
    """Backbone + PseudoLabRPN + StandardROIHeadsPseudoLab with d2's module / parameter names.  The three sub-module classes
    can be substituted (the registries of train_net.py pass the reference-shaped adapters)."""

    def __init__(self, num_classes=2, backbone_cls=None, rpn_cls=None, roi_heads_cls=None):
        super().__init__()
        self.backbone = (backbone_cls or Backbone)()
        self.proposal_generator = (rpn_cls or RPN)()
        self.roi_heads = (roi_heads_cls or ROIHeads)(num_classes)

    def adapted_parameters(self):
        """Parameters the test-time loss reaches: res3-res5 and FPN (stem + res2 are frozen, FREEZE_AT = 2; RPN / ROI
        heads are not in the TTT loss graph)."""
        return [p for group in self.adapted_parameter_groups() for p in group]

    def adapted_parameter_groups(self):
        """[FPN + res5, res4, res3]: the order in which the backward pass completes their gradients (gradient buckets)."""
        fpn = []
        for lvl in (2, 3, 4, 5):
            fpn += list(getattr(self.backbone, f"fpn_lateral{lvl}").parameters())
            fpn += list(getattr(self.backbone, f"fpn_output{lvl}").parameters())
        bu = self.backbone.bottom_up
        return [fpn + list(bu.res5.parameters()), list(bu.res4.parameters()), list(bu.res3.parameters())]

    def refresh_weight_copies(self):
        """FlatSGD(on_step=...): all tensor-core copies of the adapted weights in one launch (see refresh_weight_copies)."""
        mods = self.__dict__.get("_adapted_convs")
        if mods is None:
            ids = {id(p) for p in self.adapted_parameters()}
            mods = [m for m in self.modules() if isinstance(m, Conv2d) and id(m.weight) in ids]
            self.__dict__["_adapted_convs"] = mods
        refresh_weight_copies(mods)

    def preprocess_image(self, images_u8):
        """-> (image tensor for ``backbone.pyramid``, [(h, w) per image]): d2 preprocess_image (rcnn.py:219)."""
        _need_cuda(*[p for p in [self.backbone.fpn_output2.weight]])
        dev = self.backbone.fpn_output2.weight.device
        sizes = [tuple(im.shape[-2:]) for im in images_u8]
        if STEM_TC[0] and CONV_MODE[0] != "simt":
            x, _ = preprocess(images_u8, dev, stem_padded=True)
        else:
            x = preprocess(images_u8, dev)
        return x, sizes

    def features(self, images_u8):
        return self.backbone.pyramid(self.preprocess_image(images_u8)[0])

    def detect_ttt(self, images_u8):
        """rcnn.py:331-345: features (NHWC, grad-carrying), RPN proposals and box-head detections in TRAIN mode."""
        x, sizes = self.preprocess_image(images_u8)
        feats = self.backbone.pyramid(x)
        props, dets = self._detect(feats, sizes, training=True)
        return feats, props, dets

    @torch.no_grad()
    def _detect(self, feats, sizes, training):
        """RPN -> box head on padded device-side tensors; ONE host read (both count vectors) turns them into the per-image
        lists of the API."""
        pb, ps, pc = self.proposal_generator.predict_padded(feats, sizes, training)
        b, s, c, dc = self.roi_heads.forward_box_padded(feats, pb, pc, sizes)
        cnt = torch.stack((pc, dc)).cpu().tolist()
        props = [(pb[n, :k], ps[n, :k]) for n, k in enumerate(cnt[0])]
        dets = [(b[n, :k], s[n, :k], c[n, :k]) for n, k in enumerate(cnt[1])]
        return props, dets

    @torch.no_grad()
    def inference(self, images_u8, out_sizes=None):
        """GeneralizedRCNN.inference (rcnn.py:181-182): eval-mode detections with pasted masks.  out_sizes: (h, w) for all
        images or one per image (the dataset dict's original 'height' / 'width'); default = the network input sizes."""
        x, sizes = self.preprocess_image(images_u8)
        feats = self.backbone.pyramid(x)
        props, dets = self._detect(feats, sizes, training=False)
        return self.roi_heads.forward_mask(feats, dets, out_sizes or sizes, sizes), feats, props, dets
