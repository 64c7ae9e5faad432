"""Restatement of ``pygmtools.sinkhorn(..., backend='pytorch')`` (pygmtools 0.3.8).

TEST INFRASTRUCTURE (see oracle/__init__.py).  **Parity unpinned**: pygmtools is a
third-party dependency of the reference (``requirements.txt:58``) that is not
vendored under /root/reference and not installed in this image.  The single call
site is ``adapteacher/modeling/GModule/utils/sinkhorn.py:85-87``; it is reached
from ``multi_graph_matching.py:129-130,143`` (U_sup), ``:330-353`` (GA-GM, batched,
no grad), ``:411,435`` (HiPPI) and ``:467-468,519-522`` (MGM3_unsup, per-item,
differentiable).  The algorithm below follows the library's published log-space
Sinkhorn: tau scaling, dummy rows filled with -100,
``max_iter`` alternating row / column log-sum-exp normalisations starting with a
row step, per-item transposes for ragged batches, ``exp`` at the end.

Plain PyTorch on CPU tensors, differentiable (autograd flows through every step
exactly as in the library's per-item path).
"""
import torch

NEG_INF = -float("inf")


def _sinkhorn_core(s, nrows=None, ncols=None, dummy_row=False, max_iter=10, tau=1.0,
                   batched_operation=False):
    """s: (b, n1, n2).  Mirrors the pytorch backend function step by step."""
    batch_size = s.shape[0]

    # global orientation: the working matrix always has shape[1] <= shape[2]
    if s.shape[2] >= s.shape[1]:
        transposed = False
    else:
        s = s.transpose(1, 2)
        nrows, ncols = ncols, nrows
        transposed = True

    if nrows is None:
        nrows = torch.tensor([s.shape[1]] * batch_size, device=s.device)
    if ncols is None:
        ncols = torch.tensor([s.shape[2]] * batch_size, device=s.device)
    nrows = torch.as_tensor(nrows, device=s.device).long()
    ncols = torch.as_tensor(ncols, device=s.device).long()

    # per-item orientation for ragged batches: items whose valid block has more
    # rows than columns are transposed in place (padded with -inf columns)
    transposed_batch = nrows > ncols
    if torch.any(transposed_batch):
        s_t = s.transpose(1, 2)
        s_t = torch.cat((
            s_t[:, :s.shape[1], :],
            torch.full((batch_size, s.shape[1], s.shape[2] - s.shape[1]), NEG_INF,
                       device=s.device, dtype=s.dtype)), dim=2)
        s = torch.where(transposed_batch.view(batch_size, 1, 1), s_t, s)
        nrows, ncols = (torch.where(transposed_batch, ncols, nrows),
                        torch.where(transposed_batch, nrows, ncols))

    log_s = s / tau

    if dummy_row:
        assert log_s.shape[2] >= log_s.shape[1]
        n_dummy = log_s.shape[2] - log_s.shape[1]
        ori_nrows = nrows
        nrows = ncols.clone()
        log_s = torch.cat((log_s, torch.full((batch_size, n_dummy, log_s.shape[2]), NEG_INF,
                                             device=log_s.device, dtype=log_s.dtype)), dim=1)
        for b in range(batch_size):
            log_s[b, ori_nrows[b]:nrows[b], :ncols[b]] = -100.0

    row_mask = torch.zeros(batch_size, log_s.shape[1], 1, dtype=torch.bool, device=log_s.device)
    col_mask = torch.zeros(batch_size, 1, log_s.shape[2], dtype=torch.bool, device=log_s.device)
    for b in range(batch_size):
        row_mask[b, :nrows[b], 0] = True
        col_mask[b, 0, :ncols[b]] = True

    if batched_operation:
        valid = row_mask & col_mask
        log_s = torch.where(valid, log_s, torch.full_like(log_s, NEG_INF))
        for i in range(max_iter):
            if i % 2 == 0:
                log_sum = torch.logsumexp(log_s, 2, keepdim=True)
                log_s = log_s - torch.where(row_mask, log_sum, torch.zeros_like(log_sum))
            else:
                log_sum = torch.logsumexp(log_s, 1, keepdim=True)
                log_s = log_s - torch.where(col_mask, log_sum, torch.zeros_like(log_sum))
        ret_log_s = log_s
    else:
        ret_log_s = torch.full((batch_size, log_s.shape[1], log_s.shape[2]), NEG_INF,
                               device=log_s.device, dtype=log_s.dtype)
        for b in range(batch_size):
            rs = slice(0, int(nrows[b]))
            cs = slice(0, int(ncols[b]))
            log_s_b = log_s[b, rs, cs]
            for i in range(max_iter):
                if i % 2 == 0:
                    log_s_b = log_s_b - torch.logsumexp(log_s_b, 1, keepdim=True)
                else:
                    log_s_b = log_s_b - torch.logsumexp(log_s_b, 0, keepdim=True)
            ret_log_s[b, rs, cs] = log_s_b

    if dummy_row:
        if n_dummy > 0:
            ret_log_s = ret_log_s[:, :-n_dummy]
        for b in range(batch_size):
            ret_log_s[b, ori_nrows[b]:nrows[b], :ncols[b]] = NEG_INF

    if torch.any(transposed_batch):
        s_t = ret_log_s.transpose(1, 2)
        s_t = torch.cat((
            s_t[:, :ret_log_s.shape[1], :],
            torch.full((batch_size, ret_log_s.shape[1], ret_log_s.shape[2] - ret_log_s.shape[1]),
                       NEG_INF, device=log_s.device, dtype=log_s.dtype)), dim=2)
        ret_log_s = torch.where(transposed_batch.view(batch_size, 1, 1), s_t, ret_log_s)

    if transposed:
        ret_log_s = ret_log_s.transpose(1, 2)

    return torch.exp(ret_log_s)


def sinkhorn(s, n1=None, n2=None, unmatch1=None, unmatch2=None, dummy_row=False, max_iter=10,
             tau=1.0, batched_operation=False, backend="pytorch"):
    """Signature of the library's top-level ``pygmtools.sinkhorn`` as the reference calls it
    (``utils/sinkhorn.py:87``): accepts a 2-D matrix or a 3-D batch."""
    assert backend == "pytorch" and unmatch1 is None and unmatch2 is None
    if s.dim() == 2:
        s3, squeeze = s.unsqueeze(0), True
    elif s.dim() == 3:
        s3, squeeze = s, False
    else:
        raise ValueError(f"the input argument s is expected to be 2- or 3-dimensional, got {s.dim()}")
    out = _sinkhorn_core(s3, n1, n2, dummy_row, max_iter, tau, batched_operation)
    return out.squeeze(0) if squeeze else out
