"""Torch-facing operators of the matching stage; every one is a call into libttdg_sm100.so.

torch is plumbing here (device memory, the current stream, autograd bookkeeping).  No op has a CPU or
PyTorch-eager fallback: CPU tensors raise, a missing library raises (``_C.lib()``).
"""
import ctypes
import itertools

import torch

from . import _C
from ._C import check

NU = 32           # universe size (reference rcnn.py:116)


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _C.TTDGError("ttdg_b200 ops need CUDA tensors (there is no CPU fallback)")


def _f32c(t):
    return t.detach().to(torch.float32).contiguous()


def _i64(rows, device):
    return torch.tensor(rows, dtype=torch.int64, device=device)


# ------------------------------------------------------------------------------------------------ dense helpers
def linear(x, w, b=None):
    """y = x w^T + b with fp64 accumulation (ttdg_linear_f64acc)."""
    _need_cuda(x, w, b)
    x, w = _f32c(x), _f32c(w)
    m, k = x.shape
    n = w.shape[0]
    y = torch.empty(m, n, dtype=torch.float32, device=x.device)
    bb = None if b is None else _f32c(b)
    check(_C.lib().ttdg_linear_f64acc(_p(x), k, _p(w), w.stride(0), _p(bb), _p(y), n, m, n, k, _stream()), "linear")
    return y


def gemm(a, b, trans_a=False, trans_b=False, out=None, out_dtype=torch.float32, accumulate=False, m=None, n=None, k=None,
         lda=None, ldb=None, ldc=None):
    """C = op(A) op(B) (+ C) with fp64 accumulation; A, B, C may be fp32 or fp64 (row-major, 2-D views with a
    unit inner stride)."""
    _need_cuda(a, b, out)
    assert a.stride(-1) == 1 and b.stride(-1) == 1
    lda = a.stride(0) if lda is None else lda
    ldb = b.stride(0) if ldb is None else ldb
    if m is None:
        m = a.shape[1] if trans_a else a.shape[0]
    if k is None:
        k = a.shape[0] if trans_a else a.shape[1]
    if n is None:
        n = b.shape[0] if trans_b else b.shape[1]
    if out is None:
        out = torch.empty(m, n, dtype=out_dtype, device=a.device)
    ldc = out.stride(0) if ldc is None else ldc
    f64 = lambda t: 1 if t.dtype == torch.float64 else 0
    check(_C.lib().ttdg_gemm_f64acc(int(trans_a), int(trans_b), m, n, k, _p(a), f64(a), lda, _p(b), f64(b), ldb, _p(out),
                                    f64(out), ldc, int(accumulate), _stream()), "gemm")
    return out


# ------------------------------------------------------------------------------------------------ Sinkhorn
def _items_dense(batch, n1s, n2s, N1, N2, with_third=True):
    rows = []
    for b in range(batch):
        off = b * N1 * N2
        # a square item inherits the orientation of the padded batch (pygmtools transposes a tall batch globally)
        rows.append([off, off, -1 if with_third else off, int(n1s[b]), int(n2s[b]), N2, N2, N2, 1 if N1 > N2 else 0])
    return rows


class _SinkhornSmall(torch.autograd.Function):
    """Per-item log-Sinkhorn on a (padded) batch b x N1 x N2; entries outside n1[b] x n2[b] come out 0."""

    @staticmethod
    def forward(ctx, s, n1s, n2s, tau, max_iter, dummy_row):
        s_c = _f32c(s)
        B, N1, N2 = s_c.shape
        full = all(int(a) == N1 for a in n1s) and all(int(b) == N2 for b in n2s)
        out = torch.empty_like(s_c) if full else torch.zeros_like(s_c)
        items = _i64(_items_dense(B, n1s, n2s, N1, N2), s.device)
        max_dim = max(max(int(a) for a in n1s), max(int(b) for b in n2s), 1)
        check(_C.lib().ttdg_sinkhorn_small_fwd(_p(s_c), _p(out), _p(items), B, max_dim, float(tau), int(max_iter),
                                               int(bool(dummy_row)), _stream()), "sinkhorn_small_fwd")
        ctx.save_for_backward(s_c)
        ctx.meta = (list(map(int, n1s)), list(map(int, n2s)), float(tau), int(max_iter), bool(dummy_row), max_dim, full)
        return out

    @staticmethod
    def backward(ctx, gout):
        (s_c,) = ctx.saved_tensors
        n1s, n2s, tau, max_iter, dummy_row, max_dim, full = ctx.meta
        B, N1, N2 = s_c.shape
        g = _f32c(gout)
        gin = torch.empty_like(s_c) if full else torch.zeros_like(s_c)
        items = _i64(_items_dense(B, n1s, n2s, N1, N2, with_third=False), s_c.device)
        check(_C.lib().ttdg_sinkhorn_small_bwd(_p(s_c), _p(g), _p(gin), _p(items), B, max_dim, tau, max_iter,
                                               int(dummy_row), _stream()), "sinkhorn_small_bwd")
        return gin, None, None, None, None, None


def sinkhorn(s, nrows=None, ncols=None, dummy_row=False, max_iter=10, tau=1.0):
    """Semantics of ``pygmtools.sinkhorn(backend='pytorch')`` as the reference calls it (utils/sinkhorn.py:87):
    2-D or 3-D input, optional ragged sizes.  Small matrices (<= 96) -> fp64 shared-memory kernel
    (differentiable); larger -> the fp32 cluster-resident kernel (forward only, uniform batch)."""
    _need_cuda(s)
    squeeze = s.dim() == 2
    if s.dim() not in (2, 3):
        raise ValueError(f"the input argument s is expected to be 2- or 3-dimensional, got {s.dim()}")
    s3 = s.unsqueeze(0) if squeeze else s
    B, N1, N2 = s3.shape
    n1s = [N1] * B if nrows is None else [int(v) for v in torch.as_tensor(nrows).reshape(-1).tolist()]
    n2s = [N2] * B if ncols is None else [int(v) for v in torch.as_tensor(ncols).reshape(-1).tolist()]
    small = _C.limit("small_max_dim")
    if max(max(n1s), max(n2s)) <= small:
        out = _SinkhornSmall.apply(s3, n1s, n2s, tau, max_iter, dummy_row)
    else:
        if s3.requires_grad and torch.is_grad_enabled():
            raise ValueError(f"differentiable Sinkhorn is limited to matrices up to {small} x {small}")
        if nrows is not None or ncols is not None:
            raise ValueError("ragged batches are limited to the small-matrix path")
        out = sinkhorn_stream(s3, tau=tau, max_iter=max_iter, dummy_row=dummy_row)
    return out.squeeze(0) if squeeze else out


def sinkhorn_stream(s, tau, max_iter, dummy_row=False, out=None):
    """Large uniform batch b x n1 x n2 (fp32, forward only): one cluster per matrix, matrix resident in
    distributed shared memory.  Tall matrices are handled by transposing (the operator's own rule)."""
    _need_cuda(s)
    s3 = _f32c(s)
    if s3.dim() == 2:
        s3 = s3.unsqueeze(0)
    transposed = s3.shape[2] < s3.shape[1]
    if transposed:
        s3 = s3.transpose(1, 2).contiguous()
    B, n1, n2 = s3.shape
    o = torch.empty_like(s3) if (out is None or transposed) else out
    check(_C.lib().ttdg_sinkhorn_stream_fwd(_p(s3), _p(o), B, n1, n2, float(tau), int(max_iter), int(bool(dummy_row)),
                                            None, _stream()), "sinkhorn_stream_fwd")
    if transposed:
        o = o.transpose(1, 2).contiguous()
    return o


# ------------------------------------------------------------------------------------------------ LAP
def hungarian(s, n1=None, n2=None):
    """utils/hungarian.py:8-65: max-weight assignment as a 0/1 matrix (2-D or batched 3-D input)."""
    _need_cuda(s)
    squeeze = s.dim() == 2
    if s.dim() not in (2, 3):
        raise ValueError("input data shape not understood: {}".format(tuple(s.shape)))
    s3 = _f32c(s.unsqueeze(0) if squeeze else s)
    B, N1, N2 = s3.shape
    a = [N1] * B if n1 is None else [int(v) for v in torch.as_tensor(n1).reshape(-1).tolist()]
    b = [N2] * B if n2 is None else [int(v) for v in torch.as_tensor(n2).reshape(-1).tolist()]
    lim = _C.limit("lap_max_dim")
    if max(a + b) > lim:
        raise ValueError(f"hungarian: matrices are limited to {lim} x {lim}")
    perm = torch.zeros_like(s3)
    items = _i64([[i * N1 * N2, i * N1 * N2, a[i], b[i], N2, N2] for i in range(B)], s.device)
    check(_C.lib().ttdg_lap_solve(_p(s3), _p(perm), _p(items), B, _stream()), "lap_solve")
    return perm.squeeze(0) if squeeze else perm


# ------------------------------------------------------------------------------------------------ focal BCE
class _FocalBCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, y):
        p_c, y_c = _f32c(p), _f32c(y)
        loss = torch.empty((), dtype=torch.float32, device=p.device)
        scratch = torch.empty(_C.lib().ttdg_focal_bce_scratch_bytes(), dtype=torch.uint8, device=p.device)
        check(_C.lib().ttdg_focal_bce_fwd(_p(p_c), _p(y_c), p_c.numel(), _p(loss), _p(scratch), _stream()), "focal_bce_fwd")
        ctx.save_for_backward(p_c, y_c)
        return loss

    @staticmethod
    def backward(ctx, g):
        p_c, y_c = ctx.saved_tensors
        gp = torch.empty_like(p_c)
        gl = _f32c(g).reshape(1)
        check(_C.lib().ttdg_focal_bce_bwd(_p(p_c), _p(y_c), p_c.numel(), _p(gl), _p(gp), _stream()), "focal_bce_bwd")
        return gp, None


def focal_bce(p, y):
    """utils/losses.py:83-103 (gamma 2, alpha 0.25, clamp 1e-6, mean)."""
    _need_cuda(p, y)
    if p.shape != y.shape:
        raise ValueError("focal_bce: shape mismatch")
    return _FocalBCE.apply(p, y)


# ------------------------------------------------------------------------------------------------ attention adjacency
def attention_logits(x, wq, bq, wk, bk):
    q = linear(x, wq, bq)
    k = linear(x, wk, bk)
    return gemm(q, k, trans_b=True)                       # M x M (cross-graph blocks are not read afterwards)


def attention_adjacency(x, sizes, wq, bq, wk, bk, keep_masks=None, p_drop=0.0, seed=0, offset=0, scale=None):
    """Block-diagonal adjacency A (M x M) of mgm:497-502 for graphs stacked in x (M x 256)."""
    _need_cuda(x)
    M, d = x.shape
    G = len(sizes)
    offs = [0] + list(itertools.accumulate(int(n) for n in sizes))
    assert offs[-1] == M
    S = attention_logits(x, wq, bq, wk, bk)
    node_off = torch.tensor(offs, dtype=torch.int32, device=x.device)
    A = torch.empty(M, M, dtype=torch.float32, device=x.device)
    km, mo = None, None
    if keep_masks is not None:
        km = torch.cat([_f32c(m).reshape(-1) for m in keep_masks])
        mo = _i64([0] + list(itertools.accumulate(int(n) * int(n) for n in sizes))[:-1], x.device)
    scale = float((d // 1) ** -0.5) if scale is None else float(scale)
    check(_C.lib().ttdg_attn_adjacency(_p(S), _p(node_off), G, M, scale, _p(km), _p(mo), float(p_drop), int(seed),
                                       int(offset), _p(A), _stream()), "attn_adjacency")
    return A


# ------------------------------------------------------------------------------------------------ affinity
def _pair_tables(sizes, pairs, device):
    offs = [0] + list(itertools.accumulate(int(n) for n in sizes))
    rows, out_off, tot = [], [], 0
    for (s, t) in pairs:
        rows.append([offs[s], sizes[s], offs[t], sizes[t]])
        out_off.append(tot)
        tot += sizes[s] * sizes[t]
    return _i64(rows, device), _i64(out_off, device), out_off, tot


class _AffinityPairs(torch.autograd.Function):
    """Affinity.forward (utils/affinity.py:44-57) for many (src, tgt) graph pairs at once, separable form.
    Returns the pair blocks stored back to back in one flat tensor."""

    @staticmethod
    def forward(ctx, X, w_sr, w_tg, w0, b0, w1, b1, sizes, pairs):
        L = _C.lib()
        X_c = _f32c(X)
        M, d = X_c.shape
        hidden = w0.shape[0]
        Xp, Yp = linear(X_c, w_sr), linear(X_c, w_tg)
        ac = torch.empty(M, 2 * hidden, dtype=torch.float64, device=X.device)
        w0c, b0c, w1c, b1c = _f32c(w0), _f32c(b0), _f32c(w1).reshape(-1), _f32c(b1).reshape(-1)
        check(L.ttdg_affinity_hidden(_p(Xp), _p(Yp), _p(w0c), _p(b0c), M, d, hidden, _p(ac), _stream()), "affinity_hidden")
        ptab, otab, out_off, tot = _pair_tables(sizes, pairs, X.device)
        out = torch.empty(tot, dtype=torch.float32, device=X.device)
        check(L.ttdg_affinity_pairs_fwd(_p(ac), _p(w1c), _p(b1c), _p(ptab), _p(otab), len(pairs), max(sizes), hidden, _p(out),
                                        _stream()), "affinity_pairs_fwd")
        ctx.save_for_backward(X_c, Xp, Yp, ac, _f32c(w_sr), _f32c(w_tg), w0c, w1c, ptab, otab)
        ctx.meta = (list(sizes), len(pairs), tot, hidden, w1.shape, b1.shape)
        return out

    @staticmethod
    def backward(ctx, gout):
        L = _C.lib()
        X_c, Xp, Yp, ac, w_sr, w_tg, w0c, w1c, ptab, otab = ctx.saved_tensors
        sizes, n_pairs, tot, hidden, w1_shape, b1_shape = ctx.meta
        M, d = X_c.shape
        dev = X_c.device
        g = _f32c(gout)
        g_ac = torch.empty(M, 2 * hidden, dtype=torch.float64, device=dev)
        g_w1 = torch.empty(hidden, dtype=torch.float32, device=dev)
        g_b1 = torch.empty(1, dtype=torch.float32, device=dev)
        scratch = torch.empty(L.ttdg_affinity_bwd_scratch_bytes(M, hidden), dtype=torch.uint8, device=dev)
        check(L.ttdg_affinity_pairs_bwd(_p(ac), _p(w1c), _p(ptab), _p(otab), n_pairs, hidden, M, max(sizes), _p(g), tot, _p(g_ac),
                                        _p(g_w1), _p(g_b1), _p(scratch), _stream()), "affinity_pairs_bwd")
        g_a, g_c = g_ac[:, :hidden], g_ac[:, hidden:]
        # fc_M.0: weight (hidden x 2d) = [W0a | W0b], bias = sum over nodes of g_c
        g_w0 = torch.empty(hidden, 2 * d, dtype=torch.float32, device=dev)
        gemm(g_a, Xp, trans_a=True, out=g_w0[:, :d])
        gemm(g_c, Yp, trans_a=True, out=g_w0[:, d:])
        ones = torch.ones(M, 1, dtype=torch.float32, device=dev)
        g_b0 = gemm(g_c, ones, trans_a=True).reshape(-1)
        # projections
        g_Xp = gemm(g_a, w0c[:, :d], out_dtype=torch.float64)          # M x d
        g_Yp = gemm(g_c, w0c[:, d:], out_dtype=torch.float64)
        g_sr = gemm(g_Xp, X_c, trans_a=True)
        g_tg = gemm(g_Yp, X_c, trans_a=True)
        g_X = gemm(g_Xp, w_sr)
        gemm(g_Yp, w_tg, out=g_X, accumulate=True)
        return g_X, g_sr, g_tg, g_w0, g_b0, g_w1.reshape(w1_shape), g_b1.reshape(b1_shape), None, None


def affinity_pairs(X, w_sr, w_tg, w0, b0, w1, b1, sizes, pairs):
    _need_cuda(X, w_sr, w_tg, w0, b0, w1, b1)
    return _AffinityPairs.apply(X, w_sr, w_tg, w0, b0, w1, b1, [int(n) for n in sizes], [tuple(p) for p in pairs])


# ------------------------------------------------------------------------------------------------ GA-GM
# Called (no arguments) right after the solver of a test-time-adaptation step has been launched: the solver keeps a cluster of <= 8
# SMs busy for ~10 ms while nothing else of the step can run (the loss needs its result), so the caller may put independent work on
# ANOTHER stream here - adapteacher/engine/trainer.py evaluates a batch of the previous dataset with a weight snapshot.
SOLVER_WINDOW_HOOK = [None]
GAGM_CLUSTER_SMS = 8


def gagm_solve(A, W, U0, ms, n_univ=NU, init_tau=0.1, min_tau=1e-2, sk_gamma=0.5, max_iter=200, sk_iter=20,
               converge_tol=1e-3, quad_weight=0.5, mode=0, step_projector=0, return_info=False, trace_cap=0):
    """GA_GM.gagm (mgm:300-389) on the device: one persistent cluster, no host round trips."""
    _need_cuda(A, W, U0)
    L = _C.lib()
    A_c, W_c, U0_c = _f32c(A), _f32c(W), _f32c(U0)
    ms = [int(v) for v in ms]
    G, M = len(ms), sum(ms)
    assert A_c.shape == (M, M) and W_c.shape == (M, M) and U0_c.shape == (M, n_univ)
    ms_h = (ctypes.c_int32 * G)(*ms)
    U = torch.empty(M, n_univ, dtype=torch.float32, device=A.device)
    info = torch.zeros(16, dtype=torch.int32, device=A.device)
    scratch = torch.empty(L.ttdg_gagm_scratch_bytes(M, G), dtype=torch.uint8, device=A.device)
    trace = meta = None
    if trace_cap > 0:                   # test hook: the whole trajectory, fp64
        trace = torch.zeros(trace_cap + 1, M, n_univ, dtype=torch.float64, device=A.device)
        meta = torch.zeros(trace_cap, 2, dtype=torch.float64, device=A.device)
    check(L.ttdg_gagm_solve(_p(A_c), _p(W_c), _p(U0_c), ctypes.cast(ms_h, ctypes.c_void_p), G, M, n_univ, float(init_tau),
                            float(min_tau), float(sk_gamma), int(max_iter), int(sk_iter), float(converge_tol),
                            float(quad_weight), int(mode), int(step_projector), _p(U), _p(info), _p(scratch), _p(trace),
                            _p(meta), int(trace_cap), _stream()), "gagm_solve")
    if trace_cap > 0:
        return U, info, trace, meta
    return (U, info) if return_info else U


# ------------------------------------------------------------------------------------------------ MGM3_unsup core
def mgm_pairs(G):
    """(src, tgt) pairs with src >= tgt in the order of itertools.product (mgm:507-512)."""
    return [(s, t) for s in range(G) for t in range(G) if s >= t]


class _MatchingLoss(torch.autograd.Function):
    """aff (flat pair blocks of the learned affinity, differentiable) -> pairwise Sinkhorn -> Wds; GA-GM on
    (A, Wds.detach(), U0) -> U; focal-BCE matching loss between the Sinkhorn blocks and U_i U_j^T
    (mgm:504-564).  Only `aff` carries gradient (mgm:225, :532)."""

    @staticmethod
    def forward(ctx, aff, A, U0, sizes, cfg, U_override):
        L = _C.lib()
        dev = aff.device
        sizes = [int(n) for n in sizes]
        G, M = len(sizes), sum(sizes)
        offs = [0] + list(itertools.accumulate(sizes))
        pairs = mgm_pairs(G)
        aff_c = _f32c(aff)
        Wds = torch.empty(M, M, dtype=torch.float32, device=dev)
        rows, tot = [], 0
        for (s, t) in pairs:
            main = offs[s] * M + offs[t]
            mirror = offs[t] * M + offs[s] if s != t else -1
            rows.append([tot, main, mirror, sizes[s], sizes[t], sizes[t], M, M, 0])
            tot += sizes[s] * sizes[t]
        items = _i64(rows, dev)
        # utils/sinkhorn.py via mgm:467-468, 519-522: max_iter 20, tau 0.05, dummy_row
        check(L.ttdg_sinkhorn_small_fwd(_p(aff_c), _p(Wds), _p(items), len(pairs), max(sizes), cfg["sk_tau"], cfg["sk_iter"], 1,
                                        _stream()), "sinkhorn_small_fwd")
        info = torch.zeros(16, dtype=torch.int32, device=dev)
        if U_override is None:
            U, info = gagm_solve(A, Wds, U0, sizes, NU, cfg["ga_tau0"], cfg["ga_min_tau"], cfg["ga_gamma"], cfg["ga_iter"],
                                 cfg["ga_sk_iter"], cfg["ga_tol"], cfg["quad_weight"], return_info=True)
            if SOLVER_WINDOW_HOOK[0] is not None:
                SOLVER_WINDOW_HOOK[0]()
        else:
            U = _f32c(U_override)
        node_off = torch.tensor(offs, dtype=torch.int32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        flags = torch.zeros(1, dtype=torch.int32, device=dev)
        scratch = torch.empty(L.ttdg_matching_loss_scratch_bytes(G), dtype=torch.uint8, device=dev)
        check(L.ttdg_matching_loss_fwd(_p(Wds), _p(U), _p(node_off), G, M, NU, _p(loss), _p(flags), _p(scratch), _stream()),
              "matching_loss_fwd")
        ctx.save_for_backward(aff_c, Wds, U, node_off)
        ctx.meta = (sizes, cfg, rows, tot)
        ctx.mark_non_differentiable(Wds, U, flags, info)
        return loss, Wds, U, flags, info

    @staticmethod
    def backward(ctx, gloss, _gW, _gU, _gf, _gi):
        L = _C.lib()
        aff_c, Wds, U, node_off = ctx.saved_tensors
        sizes, cfg, rows, tot = ctx.meta
        G, M = len(sizes), sum(sizes)
        dev = aff_c.device
        gl = _f32c(gloss).reshape(1)
        gW = torch.empty(M, M, dtype=torch.float32, device=dev)
        check(L.ttdg_matching_loss_bwd(_p(Wds), _p(U), _p(node_off), G, M, NU, _p(gl), _p(gW), _stream()), "matching_loss_bwd")
        g_aff = torch.zeros(tot, dtype=torch.float32, device=dev)
        # only the off-diagonal pairs feed the loss; {s_off, gout_off, gin_off, n1, n2, ld_s, ld_gout, ld_gin}
        brow = [[r[0], r[1], r[0], r[3], r[4], r[5], M, r[5], 0] for r in rows if r[2] >= 0]
        if brow:
            items = _i64(brow, dev)
            check(L.ttdg_sinkhorn_small_bwd(_p(aff_c), _p(gW), _p(g_aff), _p(items), len(brow), max(sizes), cfg["sk_tau"],
                                            cfg["sk_iter"], 1, _stream()), "sinkhorn_small_bwd")
        return g_aff, None, None, None, None, None


MGM_DEFAULTS = dict(sk_iter=20, sk_tau=0.05,                       # mgm:467-468
                    ga_iter=200, ga_sk_iter=20, ga_tau0=0.1, ga_gamma=0.5, ga_tol=1e-3, ga_min_tau=1e-2,   # mgm:469-474
                    quad_weight=0.5)                               # mgm:457


def matching_loss(aff, A, U0, sizes, cfg=None, U_override=None):
    cfg = dict(MGM_DEFAULTS, **(cfg or {}))
    return _MatchingLoss.apply(aff, A, U0, sizes, cfg, U_override)


# ------------------------------------------------------------------------------------------------ node sampler
STRIDES = (4, 8, 16, 32, 64)


class _SamplerGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, *feats):
        L = _C.lib()
        dev = feats[0].device
        B, C = plan["B"], feats[0].shape[1]
        n_total = plan["n_total"]
        nodes = torch.empty(n_total, C, dtype=torch.float32, device=dev)
        labels = torch.empty(n_total, dtype=torch.int64, device=dev)
        ptrs = (ctypes.c_void_p * 5)(*[f.data_ptr() for f in feats])
        strides = (ctypes.c_int64 * 15)(*[v for f in feats for v in (f.stride(0), f.stride(1), f.stride(3))])
        for f in feats:
            assert f.dtype == torch.float32 and f.stride(2) == f.shape[3] * f.stride(3), "feature maps must be dense per image"
        check(L.ttdg_sampler_gather(ctypes.cast(ptrs, ctypes.c_void_p), ctypes.cast(strides, ctypes.c_void_p),
                                    ctypes.cast(plan["hw"], ctypes.c_void_p), B, C, n_total, _p(plan["label"]),
                                    _p(plan["sel_idx"]), plan["max_per_level"], _p(plan["node_off"]), _p(nodes), _p(labels),
                                    _stream()), "sampler_gather")
        ctx.plan = plan
        ctx.shapes = [(tuple(f.shape), tuple(f.stride())) for f in feats]
        ctx.mark_non_differentiable(labels)
        return nodes, labels

    @staticmethod
    def backward(ctx, gnodes, _gl):
        L = _C.lib()
        plan = ctx.plan
        g = _f32c(gnodes)
        dev = g.device
        # same memory layout as the feature maps (NHWC pyramids stay NHWC: no transposes on the way back)
        grads = [torch.empty_strided(shape, stride, dtype=torch.float32, device=dev).zero_() for shape, stride in ctx.shapes]
        ptrs = (ctypes.c_void_p * 5)(*[t.data_ptr() for t in grads])
        strides = (ctypes.c_int64 * 15)(*[v for t in grads for v in (t.stride(0), t.stride(1), t.stride(3))])
        check(L.ttdg_sampler_scatter_bwd(_p(g), ctypes.cast(ptrs, ctypes.c_void_p), ctypes.cast(strides, ctypes.c_void_p),
                                         ctypes.cast(plan["hw"], ctypes.c_void_p), plan["B"], g.shape[1], plan["n_total"],
                                         _p(plan["sel_idx"]), plan["max_per_level"], _p(plan["node_off"]), _stream()),
              "sampler_scatter_bwd")
        return (None, *grads)


def sample_nodes(features, boxes_per_image, classes_per_image, sample_dist=10):
    """PrototypeComputation.__call__ (build_graph.py:160-250).  ``features``: 5 maps B x C x H x W (strides 4..64,
    any dense layout); boxes / classes: per-image tensors.  Returns (list of n_i x C nodes, list of int64 labels)
    or (None, None).  One small device->host copy (5 counts per image) sizes the ragged outputs."""
    _need_cuda(*features)
    L = _C.lib()
    dev = features[0].device
    keep = [i for i, b in enumerate(boxes_per_image) if len(b) > 0]
    if not keep:
        return None, None
    B = len(keep)                        # listed image b reads feature image b (build_graph.py:79 vs :181)
    boxes = torch.cat([_f32c(boxes_per_image[i]).reshape(-1, 4) for i in keep]).to(dev)
    classes = torch.cat([classes_per_image[i].reshape(-1).to(torch.int64) for i in keep]).to(dev)
    box_off = torch.tensor([0] + list(itertools.accumulate(len(boxes_per_image[i]) for i in keep)), dtype=torch.int32,
                           device=dev)
    hw = (ctypes.c_int32 * 10)(*[v for f in features for v in (f.shape[2], f.shape[3])])
    Ltot = sum(f.shape[2] * f.shape[3] for f in features)
    max_per_level = 32
    label = torch.empty(B, Ltot, dtype=torch.int32, device=dev)
    counts = torch.empty(B * 5, dtype=torch.int32, device=dev)
    sel_idx = torch.zeros(B * 5 * max_per_level, dtype=torch.int32, device=dev)
    check(L.ttdg_sampler_select(_p(boxes), _p(classes), _p(box_off), B, ctypes.cast(hw, ctypes.c_void_p), int(sample_dist),
                                max_per_level, _p(label), _p(counts), _p(sel_idx), _stream()), "sampler_select")
    cnt = counts.cpu().tolist()          # the one host sync of the sampler (ragged output shapes)
    node_off_h = [0] + list(itertools.accumulate(cnt))
    plan = dict(B=B, hw=hw, label=label, sel_idx=sel_idx, max_per_level=max_per_level, n_total=node_off_h[-1],
                node_off=torch.tensor(node_off_h, dtype=torch.int32, device=dev))
    nodes, labels = _SamplerGather.apply(plan, *features)
    per_img = [sum(cnt[5 * b:5 * b + 5]) for b in range(B)]
    return list(torch.split(nodes, per_img)), list(torch.split(labels, per_img))


# ------------------------------------------------------------------------------------------------ evaluator (counts)
def mask_gt_stats(gt):
    """gt: G x H x W uint8 / bool (CUDA) -> int64 [G][5] = {n, sum rows, sum cols, split_y, split_x} (dice_metric.py:227-229)."""
    _need_cuda(gt)
    gt = gt.to(torch.uint8).contiguous()
    G, H, W = gt.shape
    out = torch.zeros(G, 5, dtype=torch.int64, device=gt.device)
    check(_C.lib().ttdg_mask_gt_stats(_p(gt), G, H, W, _p(out), _stream()), "mask_gt_stats")
    return out


def mask_pair_counts(pred, gt, pairs, gt_stats):
    """Per (prediction index, ground-truth index) pair the 4 quadrants x (n11, n10, n01, n00) pixel counts: int64 [n][16]."""
    _need_cuda(pred, gt, gt_stats)
    pred = (pred.view(torch.uint8) if pred.dtype == torch.bool else pred.to(torch.uint8)).contiguous()
    gt = gt.to(torch.uint8).contiguous()
    H, W = pred.shape[-2:]
    if tuple(gt.shape[-2:]) != (H, W):
        raise ValueError("prediction and ground-truth masks differ in size")
    pt = torch.tensor(list(pairs), dtype=torch.int32, device=pred.device).reshape(-1, 2)
    if len(pt) and (int(pt[:, 0].max()) >= pred.shape[0] or int(pt[:, 1].max()) >= gt.shape[0] or int(pt.min()) < 0):
        raise ValueError("pair index out of range")
    out = torch.zeros(len(pt), 16, dtype=torch.int64, device=pred.device)
    check(_C.lib().ttdg_mask_pair_counts(_p(pred), _p(gt), _p(pt), len(pt), _p(gt_stats), H, W, _p(out), _stream()), "mask_pair_counts")
    return out


# ------------------------------------------------------------------------------------------------ test data path
_RESIZE_TABLES = {}


def resize_tables(in_size, out_size, device):
    """Device copy of Pillow's fixed-point coefficient table of one axis (built on the host by the library, cached per
    (in, out, device)): (bounds int32 [out, 2], kk int32 [out, ksize], ksize)."""
    key = (int(in_size), int(out_size), str(device))
    hit = _RESIZE_TABLES.get(key)
    if hit is None:
        L = _C.lib()
        ks = L.ttdg_resize_ksize(int(in_size), int(out_size))
        if ks < 1:
            raise ValueError("resize: sizes must be positive")
        bounds = torch.empty(out_size, 2, dtype=torch.int32)
        kk = torch.empty(out_size, ks, dtype=torch.int32)
        got = L.ttdg_resize_coeffs_u8(int(in_size), int(out_size), _p(bounds), _p(kk))
        if got != ks:
            raise _C.TTDGError(f"ttdg_resize_coeffs_u8 failed ({got})")
        if len(_RESIZE_TABLES) > 256:
            _RESIZE_TABLES.clear()
        hit = (bounds.to(device), kk.to(device), ks)
        _RESIZE_TABLES[key] = hit
    return hit


def resize_bilinear_u8(img_hwc, new_h, new_w, planar=True, flip=False):
    """PIL.Image.resize((new_w, new_h), Image.BILINEAR) on the device, bit-exact (d2 ResizeTransform.apply_image for uint8 images).
    img_hwc: uint8 H x W x C CUDA tensor (C = 1, 3, 4) -> uint8 C x new_h x new_w (planar) or new_h x new_w x C; flip reverses the
    channel order (INPUT.FORMAT 'BGR')."""
    _need_cuda(img_hwc)
    if img_hwc.dtype != torch.uint8 or img_hwc.dim() != 3:
        raise ValueError("resize_bilinear_u8 expects a uint8 H x W x C tensor")
    img_hwc = img_hwc.contiguous()
    H, W, C = img_hwc.shape
    dev = img_hwc.device
    bx = kx = by = ky = tmp = None
    ksx = ksy = 0
    if new_w != W:
        bx, kx, ksx = resize_tables(W, new_w, dev)
        tmp = torch.empty(H, new_w, C, dtype=torch.uint8, device=dev)
    if new_h != H:
        by, ky, ksy = resize_tables(H, new_h, dev)
    out = torch.empty((C, new_h, new_w) if planar else (new_h, new_w, C), dtype=torch.uint8, device=dev)
    check(_C.lib().ttdg_resize_bilinear_u8(_p(img_hwc), H, W, C, _p(bx), _p(kx), int(ksx), _p(by), _p(ky), int(ksy), int(new_h), int(new_w),
                                           _p(tmp), _p(out), int(bool(planar)), int(bool(flip)), _stream()), "resize_bilinear_u8")
    return out
