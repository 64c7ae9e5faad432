"""Plain-PyTorch CPU port of the reference's multi-graph-matching stage.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Unlike ``ref_shim`` this file travels to the GPU
box, where it is the checker for the CUDA path and the ``cpu_baseline`` ("port") of bench.py.
It is pinned against the reference's own modules in ``tests/test_oracle_vs_reference.py`` (run in
the build container) and against ``tests/golden/*.npz``.

Functional style: every function takes explicit weights (a state dict with the reference's key
names) instead of nn.Modules.  Citations are to /root/reference/adapteacher/modeling/GModule.
"""
import itertools

import numpy as np
import scipy.optimize
import torch
import torch.nn.functional as F

from oracle.pygm_sinkhorn import sinkhorn as pygm_sinkhorn


# ---------------------------------------------------------------- operators
def sinkhorn(s, nrows=None, ncols=None, dummy_row=False, max_iter=10, tau=1.0, batched_operation=False):
    """utils/sinkhorn.py:58-87 (log_forward path; ``epsilon`` is ignored there)."""
    return pygm_sinkhorn(s, n1=nrows, n2=ncols, dummy_row=dummy_row, max_iter=max_iter, tau=tau,
                         batched_operation=batched_operation, backend="pytorch")


def hungarian(s):
    """utils/hungarian.py:8-65 for a single 2-D score matrix (n1/n2 None, nproc 1)."""
    cost = s.detach().cpu().numpy() * -1
    row, col = scipy.optimize.linear_sum_assignment(cost)
    perm = np.zeros_like(cost)
    perm[row, col] = 1
    return torch.from_numpy(perm).to(s.device)


def affinity(sd, X, Y, prefix="node_affinity."):
    """utils/affinity.py:44-57: project, broadcast-concat to N1 x N2 x 512, MLP 512-512-ReLU-1."""
    Xp = F.linear(X, sd[prefix + "project_sr.weight"])
    Yp = F.linear(Y, sd[prefix + "project_tg.weight"])
    n1, n2, c = Xp.shape[0], Yp.shape[0], Xp.shape[1]
    cat = torch.cat([Xp.unsqueeze(1).expand(n1, n2, c), Yp.unsqueeze(0).expand(n1, n2, c)], dim=-1)
    h = F.relu(F.linear(cat, sd[prefix + "fc_M.0.weight"], sd[prefix + "fc_M.0.bias"]))
    return F.linear(h, sd[prefix + "fc_M.2.weight"], sd[prefix + "fc_M.2.bias"]).squeeze()


def attention_adjacency(sd, x, keep_mask=None, p=0.1, prefix="intra_domain_graph."):
    """utils/attentions.py:60-86 (version 'v2', one head) keeping only the attention map, which is
    all ``MGM3_unsup`` uses (mgm:498).  ``keep_mask`` (n x n of 0/1) is the dropout keep-mask of
    attentions.py:40; ``None`` = eval mode (no dropout)."""
    k = F.linear(x.unsqueeze(1), sd[prefix + "linear_k.weight"], sd[prefix + "linear_k.bias"])
    q = F.linear(x.unsqueeze(1), sd[prefix + "linear_q.weight"], sd[prefix + "linear_q.bias"])
    n, d = x.shape
    k = k.view(n, 1, d).transpose(0, 1)
    q = q.view(n, 1, d).transpose(0, 1)
    scale = (d // 1) ** -0.5
    att = torch.softmax(torch.bmm(q, k.transpose(1, 2)) * scale, dim=2)
    if keep_mask is not None:
        att = att * (keep_mask.reshape(att.shape).to(att.dtype) / (1.0 - p))
    return att.squeeze()


def focal_bce(pt, target, gamma=2, alpha=0.25, eps=1e-6):
    """utils/losses.py:83-103 (mean reduction), reached from PermutationLoss.forward :419-455."""
    pt = pt.clamp(min=eps, max=1 - eps)
    loss = -alpha * (1 - pt) ** gamma * target * torch.log(pt) \
        - (1 - alpha) * pt ** gamma * (1 - target) * torch.log(1 - pt)
    return loss.mean()


# ---------------------------------------------------------------- GA-GM solver
def gagm_step(A, W, U, ms, n_univ, projector, tau, sk_iter=20, quad_weight=0.5, noise=None, return_V=False):
    """One iteration of the loop body of ``gagm`` (multi_graph_matching.py:312-359): V update, projection
    (batched Sinkhorn :330-353 or per-graph Hungarian :324-328) and the G == 2 quirk (:358-359)."""
    G = len(ms)
    offs = np.concatenate([[0], np.cumsum(ms)]).tolist()
    msl = torch.tensor(ms)
    UUt = torch.mm(U, U.t())
    V = torch.linalg.multi_dot([A, UUt, A, U]) * quad_weight * 2 + torch.mm(W, U)   # cluster weight == 1
    V = V / G
    if noise is not None:      # sensitivity experiments only: relative perturbation of V
        V = V * (1.0 + noise[0] * torch.randn(V.shape, generator=noise[1], dtype=V.dtype))
    if projector == "hungarian":
        U = torch.cat([hungarian(V[offs[g]:offs[g + 1], :n_univ]) for g in range(G)], dim=0)
    else:
        if all(m == ms[0] for m in ms):
            Vb = V.reshape(G, -1, n_univ)
            if ms[0] <= n_univ:
                U = sinkhorn(Vb, dummy_row=True, max_iter=sk_iter, tau=tau,
                             batched_operation=True).reshape(-1, n_univ)
            else:
                U = sinkhorn(Vb.transpose(1, 2), dummy_row=True, max_iter=sk_iter, tau=tau,
                             batched_operation=True).transpose(1, 2).reshape(-1, n_univ)
        else:
            mx = max(ms)
            Vp = torch.stack([F.pad(V[offs[g]:offs[g + 1], :n_univ], (0, 0, 0, mx - ms[g]))
                              for g in range(G)], dim=0)
            Ub = sinkhorn(Vp, msl, dummy_row=True, max_iter=sk_iter, tau=tau, batched_operation=True)
            U = torch.cat([Ub[g, :ms[g], :] for g in range(G)], dim=0)
    if G == 2:
        U[:ms[0], :] = torch.eye(ms[0], n_univ, dtype=U.dtype)
    return (U, V) if return_V else U


def gagm(A, W, U0, ms, n_univ, init_tau=0.1, min_tau=1e-2, max_iter=200, sk_iter=20, sk_gamma=0.5,
         converge_tol=1e-3, quad_weight=0.5, trace=None, noise=None, precise=False):
    """multi_graph_matching.py:300-389 with projector0='sinkhorn', hung_iter=True, num_clusters=1
    (cluster_M == 1).  ``ms`` is a python list of graph sizes.  Returns U (M x n_univ).

    ``precise=True`` runs the identical iteration in float64 (inputs are the fp32 A, W, U0 widened
    exactly).  The fp32 trajectory of the reference is NOT reproducible by the reference itself: a
    1e-7 change of W (e.g. MKL with 8 threads instead of 1) flips most of U on several seeded cases,
    because ~200 discrete Hungarian/Sinkhorn iterations amplify rounding noise.  The float64
    trajectory is the noise-free limit of the same algorithm; the CUDA solver (fp64 internally) is
    checked against it bit for bit, and against the reference's own fp32 result on the cases where
    that result is stable."""
    if precise:
        A, W, U0 = A.double(), W.double(), U0.double()
    G = len(ms)
    U = U0
    lastU = torch.zeros_like(U)
    tau = init_tau
    projector = "sinkhorn"
    while True:
        for i in range(max_iter):
            lastU2, lastU = lastU, U
            if trace is not None:
                trace.append((projector, tau, i, lastU.clone()))
            U = gagm_step(A, W, U, ms, n_univ, projector, tau, sk_iter, quad_weight, noise)
            if torch.norm(U - lastU) < converge_tol or torch.norm(U - lastU2) == 0:
                break
        # (mgm:364-371) "not converged" with hung_iter=True is a no-op
        if projector == "hungarian":
            break
        elif tau > min_tau:
            tau *= sk_gamma
        else:
            projector = "hungarian"
    return U.float() if precise else U


# ---------------------------------------------------------------- MGM3_unsup.forward
def mgm3_unsup_forward(sd, nodes, labels, U_univ, keep_masks=None, univ_size=32, quad_weight=0.5,
                       return_aux=False, precise=False, U_override=None):
    """multi_graph_matching.py:487-569.  ``nodes``: list of n_i x 256 (may require grad);
    ``labels`` unused by the loss; ``U_univ``: 32 x 256."""
    if nodes is None or len(nodes) == 1:
        return None
    ms = [int(n.shape[0]) for n in nodes]
    G, M = len(ms), sum(ms)
    offs = np.concatenate([[0], np.cumsum(ms)]).tolist()

    dt = nodes[0].dtype          # float64 inputs run the whole pipeline in float64 (noise-free gradients)
    A = torch.zeros(M, M, dtype=dt)
    for g, x in enumerate(nodes):
        adj = attention_adjacency(sd, x, None if keep_masks is None else keep_masks[g])
        A[offs[g]:offs[g + 1], offs[g]:offs[g + 1]] += adj.reshape(ms[g], ms[g]).detach()
    A.fill_diagonal_(0)

    Wds = torch.zeros(M, M, dtype=dt)
    for si, ti in itertools.product(range(G), repeat=2):
        if si < ti:
            continue
        Wij = affinity(sd, nodes[si], nodes[ti]).reshape(ms[si], ms[ti])
        if ms[ti] >= ms[si]:
            ds = sinkhorn(Wij, dummy_row=True, max_iter=20, tau=0.05)
        else:
            ds = sinkhorn(Wij.t(), dummy_row=True, max_iter=20, tau=0.05).t()
        Wds[offs[si]:offs[si + 1], offs[ti]:offs[ti + 1]] += ds
        if si != ti:
            Wds[offs[ti]:offs[ti + 1], offs[si]:offs[si + 1]] += ds.t()

    U0 = torch.cat([torch.mm(x, U_univ.t()) for x in nodes], dim=0).detach()
    if U_override is not None:
        Ub = U_override.to(dt)
    else:
        Ub = gagm(A, Wds.detach(), U0, ms, univ_size, quad_weight=quad_weight, precise=precise)
    Us = [Ub[offs[g]:offs[g + 1]] for g in range(G)]

    loss = 0
    npairs = 0
    for i1, i2 in itertools.combinations(range(G), 2):
        if ms[i2] >= ms[i1]:
            s = Wds[offs[i1]:offs[i1 + 1], offs[i2]:offs[i2 + 1]]
        else:
            s = Wds[offs[i2]:offs[i2 + 1], offs[i1]:offs[i1 + 1]].t()
        x_gt = torch.mm(Us[i1], Us[i2].t())
        assert torch.all((s >= 0) & (s <= 1)) and torch.all((x_gt >= 0) & (x_gt <= 1))
        loss = loss + (torch.tensor(0., dtype=dt) + focal_bce(s, x_gt))
        npairs += 1
    loss = loss / npairs
    if return_aux:
        return loss, {"A": A.detach(), "Wds": Wds.detach(), "U0": U0, "U": Ub}
    return loss


# ---------------------------------------------------------------- node sampler
INF = 100000000


def sample_nodes(features, boxes_per_image, classes_per_image, sample_dist=10):
    """build_graph.py:160-250 (+ :27-68, :70-115, :133-157).  ``features``: 5 maps B x C x H x W
    (strides 4..64); ``boxes_per_image``: list of (k_i x 4) xyxy; ``classes_per_image``: list of int64
    (k_i).  Returns (nodes list, labels list) or (None, None).  Reproduces the image-index
    misalignment after an image without boxes."""
    strides = [4, 8, 16, 32, 64]
    ranges = [[-1, 64], [64, 128], [128, 256], [256, 512], [512, INF]]
    if not any(len(b) for b in boxes_per_image):
        return None, None
    C = features[0].shape[1]
    locs, sizes_of_interest = [], []
    for l, f in enumerate(features):
        h, w = f.shape[-2:]
        sx = torch.arange(0, w * strides[l], step=strides[l], dtype=torch.float32)
        sy = torch.arange(0, h * strides[l], step=strides[l], dtype=torch.float32)
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        p = torch.stack((xx.reshape(-1), yy.reshape(-1)), dim=1) + strides[l] // 2
        locs.append(p)
        sizes_of_interest.append(torch.tensor(ranges[l], dtype=torch.float32)[None].expand(len(p), -1))
    npl = [len(p) for p in locs]
    pts = torch.cat(locs, 0)
    soi = torch.cat(sizes_of_interest, 0)
    xs, ys = pts[:, 0], pts[:, 1]
    labels_img = []
    for boxes, classes in zip(boxes_per_image, classes_per_image):
        if len(boxes) == 0:
            continue                                    # (:79) skipped -> later index misalignment
        lab = classes + 1
        area = (boxes[:, 2] - boxes[:, 0] + 1) * (boxes[:, 3] - boxes[:, 1] + 1)
        ltrb = torch.stack([xs[:, None] - boxes[:, 0][None], ys[:, None] - boxes[:, 1][None],
                            boxes[:, 2][None] - xs[:, None], boxes[:, 3][None] - ys[:, None]], dim=2)
        inside = ltrb.min(dim=2)[0] > 0
        mx = ltrb.max(dim=2)[0]
        cared = (mx >= soi[:, [0]]) & (mx <= soi[:, [1]])
        a = area[None].repeat(len(pts), 1)
        a[inside == 0] = INF
        a[cared == 0] = INF
        amin, ind = a.min(dim=1)
        lab = lab[ind]
        lab[amin == INF] = 0
        labels_img.append(torch.split(lab, npl, dim=0))
    out_nodes, out_labels = [], []
    for b in range(len(labels_img)):
        pn, pl = [], []
        for l, lab in enumerate(labels_img[b]):
            feat = features[l][b].permute(1, 2, 0).reshape(-1, C)       # (:181) b = position in the shortened list
            pos = lab > 0
            fa, la = feat[pos], lab[pos]
            step = len(la) // sample_dist
            if step > 1:
                fa, la = fa[::step], la[::step]
            pn.append(fa)
            pl.append(la)
        out_nodes.append(torch.cat(pn, 0))
        out_labels.append(torch.cat(pl, 0))
    return out_nodes, out_labels
