"""Seeded synthetic inputs shared by the bench, the tests and the golden-vector generator.

SURVEY.md section 8(d): node features ``randn(n_i, 256)``, labels ``randint(1, K+1)``, dropout
keep-masks from ``gen(seed + 7)``; MGM parameters drawn name-by-name from one generator so the
reference modules (via ``oracle/ref_shim.py``) and our mirror load the *same* state dict without
depending on module construction order.  Distributions follow the reference constructors
(``utils/affinity.py:33-42``: N(0, 0.01), zero bias; ``multi_graph_matching.py:124``:
``randn + 1/univ``; ``nn.Linear`` default uniform(+-1/sqrt(fan_in)); LayerNorm ones/zeros).
"""
import math

import torch

MGM_CASES = {
    # name: (sizes, seed)      sizes follow SURVEY 8(d)
    "g4x40": ((40, 40, 40, 40), 11),          # shapes of reference rcnn.py:525-531
    "ragged4": ((23, 40, 31, 57), 12),
    "g5x30": ((30, 30, 30, 30, 30), 13),      # all n < 32
    "g2x32": ((32, 32), 14),                  # G == 2 quirk (mgm:358-359), square
    "g3mix32": ((32, 45, 20), 15),            # ragged with an n == 32 item and max_n > 32
    "g8x60": ((60,) * 8, 16),
}


def _uniform(gen, shape, bound):
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2.0 - 1.0) * bound


def mgm_unsup_state(seed=0, dim=256):
    """State dict for ``MGM3_unsup`` (keys as measured in SURVEY 8b)."""
    g = torch.Generator().manual_seed(1_000_003 + seed)
    sd = {}
    p = "node_affinity."
    sd[p + "fc_M.0.weight"] = torch.randn(512, 512, generator=g) * 0.01
    sd[p + "fc_M.0.bias"] = torch.zeros(512)
    sd[p + "fc_M.2.weight"] = torch.randn(1, 512, generator=g) * 0.01
    sd[p + "fc_M.2.bias"] = torch.zeros(1)
    sd[p + "project_sr.weight"] = torch.randn(dim, dim, generator=g) * 0.01
    sd[p + "project_tg.weight"] = torch.randn(dim, dim, generator=g) * 0.01
    p = "intra_domain_graph."
    b = 1.0 / math.sqrt(dim)
    for name in ("linear_k", "linear_v", "linear_q", "linear_final"):
        sd[p + name + ".weight"] = _uniform(g, (dim, dim), b)
        sd[p + name + ".bias"] = _uniform(g, (dim,), b)
    sd[p + "layer_norm.weight"] = torch.ones(dim)
    sd[p + "layer_norm.bias"] = torch.zeros(dim)
    return sd


def perturb_affinity_state(sd, seed=0, scale=0.06):
    """A 'trained-looking' variant: larger affinity weights + non-zero biases so the Sinkhorn input
    is not nearly constant (fresh-init affinities are ~1e-4 apart, which makes every LAP a near tie)."""
    g = torch.Generator().manual_seed(2_000_003 + seed)
    out = dict(sd)
    p = "node_affinity."
    out[p + "fc_M.0.weight"] = torch.randn(512, 512, generator=g) * scale
    out[p + "fc_M.0.bias"] = torch.randn(512, generator=g) * 0.1
    out[p + "fc_M.2.weight"] = torch.randn(1, 512, generator=g) * scale
    out[p + "fc_M.2.bias"] = torch.randn(1, generator=g) * 0.1
    out[p + "project_sr.weight"] = torch.randn(256, 256, generator=g) * scale
    out[p + "project_tg.weight"] = torch.randn(256, 256, generator=g) * scale
    return out


def universe(seed=0, univ_size=32, dim=256):
    """``U_sup.U`` (multi_graph_matching.py:124)."""
    g = torch.Generator().manual_seed(3_000_003 + seed)
    return torch.randn(univ_size, dim, generator=g) + 1.0 / univ_size


def mgm_inputs(sizes, seed, num_classes=2, dim=256, p_drop=0.1):
    """nodes (list of n_i x dim), labels (list of int64 in 1..K), dropout keep-masks (list n_i x n_i)."""
    g = torch.Generator().manual_seed(seed)
    nodes = [torch.randn(n, dim, generator=g) for n in sizes]
    labels = [torch.randint(1, num_classes + 1, (n,), generator=g) for n in sizes]
    gm = torch.Generator().manual_seed(seed + 7)
    masks = [(torch.rand(n, n, generator=gm) >= p_drop).to(torch.float32) for n in sizes]
    return nodes, labels, masks


def fundus_like_image(idx, size=512, polyp=False):
    """uint8 3 x S x S synthetic image plus its ground-truth boxes / classes / masks."""
    g = torch.Generator().manual_seed(1000 + idx)
    S = size
    img = torch.randn(3, S, S, generator=g) * 10.0 + 40.0
    yy, xx = torch.meshgrid(torch.arange(S, dtype=torch.float32), torch.arange(S, dtype=torch.float32),
                            indexing="ij")
    u = torch.rand(8, generator=g)
    if polyp:
        cx, cy = (0.3 + 0.4 * u[0]) * S, (0.3 + 0.4 * u[1]) * S
        rx, ry = (0.1 + 0.1 * u[2]) * S, (0.1 + 0.1 * u[3]) * S
        shapes = [(cx, cy, rx, ry, 0, 150.0)]
    else:
        cx, cy = (0.4 + 0.2 * u[0]) * S, (0.4 + 0.2 * u[1]) * S
        rx, ry = (0.13 + 0.05 * u[2]) * S, (0.13 + 0.05 * u[3]) * S
        k = 0.4 + 0.3 * u[4]
        shapes = [(cx, cy, rx, ry, 0, 140.0), (cx, cy, rx * k, ry * k, 1, 220.0)]
    boxes, classes, masks = [], [], []
    for (cx, cy, rx, ry, cls, val) in shapes:
        m = ((xx - cx) / rx) ** 2 + ((yy - cy) / ry) ** 2 <= 1.0
        img = torch.where(m[None], torch.full_like(img, val) * torch.tensor([1.0, 0.8, 0.6]).view(3, 1, 1), img)
        boxes.append([float(cx - rx), float(cy - ry), float(cx + rx), float(cy + ry)])
        classes.append(cls)
        masks.append(m)
    img = img + torch.randn(3, S, S, generator=g) * 5.0
    img = img.clamp(0, 255).to(torch.uint8)
    return {"image": img, "height": S, "width": S, "image_id": idx,
            "gt_boxes": torch.tensor(boxes), "gt_classes": torch.tensor(classes), "gt_masks": torch.stack(masks)}


SAMPLER_CASES = {
    "samp_a": dict(S=256, C=16, seed=777,
                   boxes=[[[40., 50., 200., 220.], [90., 100., 150., 170.]], [[10., 10., 250., 250.]]],
                   classes=[[0, 1], [1]]),
    # image 1 has no boxes -> exercises the index misalignment quirk (build_graph.py:79 vs :181)
    "samp_skip": dict(S=128, C=16, seed=778,
                      boxes=[[[8., 8., 100., 90.]], [], [[30., 20., 120., 110.], [50., 40., 90., 80.]]],
                      classes=[[1], [], [0, 1]]),
    "samp_big": dict(S=512, C=8, seed=779,
                     boxes=[[[100., 120., 420., 400.], [200., 210., 330., 320.], [5., 5., 40., 30.]]],
                     classes=[[0, 1, 1]]),
}


def sampler_feats(name):
    c = SAMPLER_CASES[name]
    g = torch.Generator().manual_seed(c["seed"])
    B = len(c["boxes"])
    return [torch.randn(B, c["C"], c["S"] // s, c["S"] // s, generator=g) for s in (4, 8, 16, 32, 64)]
