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


# ---------------------------------------------------------------------------------------------- detector weights
RES_STAGES = (("res2", 3, 64, 256, 1), ("res3", 4, 128, 512, 2), ("res4", 6, 256, 1024, 2), ("res5", 3, 512, 2048, 2))


def detector_state(seed=0, num_classes=2, num_anchors=15):
    """Random-init state dict of the Mask R-CNN R50-FPN detector with Detectron2 0.5 parameter names and
    init distributions ([recalled]): c2_msra_fill (kaiming normal, fan_out) for ResNet / mask-head
    convs, c2_xavier_fill (kaiming uniform, a=1) for FPN and box FCs, N(0, .01) for RPN and cls_score, N(0, .001) for
    bbox_pred and the mask predictor, zero biases.  FrozenBN: gamma 1, beta 0, and - unlike a fresh d2 model - mildly
    randomised running statistics so the fold scale/bias path is exercised (mean N(0, .1), var U(.5, 1.5)); no trained
    checkpoint is available offline (README.md:96)."""
    g = torch.Generator().manual_seed(5_000_003 + seed)
    sd = {}

    def msra(name, cout, cin, k):
        std = math.sqrt(2.0 / (cout * k * k))
        sd[name + ".weight"] = torch.randn(cout, cin, k, k, generator=g) * std

    def xavier(name, shape, fan_in):
        bound = math.sqrt(3.0 / fan_in)               # kaiming_uniform_(a=1)
        sd[name + ".weight"] = _uniform(g, shape, bound)
        sd[name + ".bias"] = torch.zeros(shape[0])

    def normal(name, shape, std):
        sd[name + ".weight"] = torch.randn(*shape, generator=g) * std
        sd[name + ".bias"] = torch.zeros(shape[0])

    def bn(name, c):
        sd[name + ".weight"] = torch.ones(c)
        sd[name + ".bias"] = torch.zeros(c)
        sd[name + ".running_mean"] = torch.randn(c, generator=g) * 0.1
        sd[name + ".running_var"] = torch.rand(c, generator=g) + 0.5

    p = "backbone.bottom_up."
    msra(p + "stem.conv1", 64, 3, 7)
    bn(p + "stem.conv1.norm", 64)
    cin = 64
    for stage, blocks, mid, cout, _ in RES_STAGES:
        for b in range(blocks):
            q = f"{p}{stage}.{b}."
            if b == 0:
                msra(q + "shortcut", cout, cin, 1)
                bn(q + "shortcut.norm", cout)
            msra(q + "conv1", mid, cin, 1)
            bn(q + "conv1.norm", mid)
            msra(q + "conv2", mid, mid, 3)
            bn(q + "conv2.norm", mid)
            msra(q + "conv3", cout, mid, 1)
            bn(q + "conv3.norm", cout)
            cin = cout
    for lvl, c in zip((2, 3, 4, 5), (256, 512, 1024, 2048)):
        xavier(f"backbone.fpn_lateral{lvl}", (256, c, 1, 1), c)
        xavier(f"backbone.fpn_output{lvl}", (256, 256, 3, 3), 256 * 9)
    r = "proposal_generator.rpn_head."
    normal(r + "conv", (256, 256, 3, 3), 0.01)
    normal(r + "objectness_logits", (num_anchors, 256, 1, 1), 0.01)
    normal(r + "anchor_deltas", (num_anchors * 4, 256, 1, 1), 0.01)
    h = "roi_heads."
    xavier(h + "box_head.fc1", (1024, 256 * 7 * 7), 256 * 7 * 7)
    xavier(h + "box_head.fc2", (1024, 1024), 1024)
    normal(h + "box_predictor.cls_score", (num_classes + 1, 1024), 0.01)
    normal(h + "box_predictor.bbox_pred", (num_classes * 4, 1024), 0.001)
    for i in range(1, 5):
        msra(h + f"mask_head.mask_fcn{i}", 256, 256, 3)
        sd[h + f"mask_head.mask_fcn{i}.bias"] = torch.zeros(256)
    std = math.sqrt(2.0 / (256 * 2 * 2))
    sd[h + "mask_head.deconv.weight"] = torch.randn(256, 256, 2, 2, generator=g) * std      # ConvTranspose2d: (in, out, kh, kw)
    sd[h + "mask_head.deconv.bias"] = torch.zeros(256)
    normal(h + "mask_head.predictor", (num_classes, 256, 1, 1), 0.001)
    return sd


def calibrate_frozen_bn(sd, size=128, n_images=2, seed=0):
    """Sets every FrozenBN's running statistics to the statistics its input actually has on a few synthetic images
    (what training would have left in a real checkpoint), so activations stay O(1) through the 16 residual blocks
    and the detector produces a realistic workload (about a thousand proposals per image after NMS).  Plain torch on
    the CPU, deterministic for a given machine; both the oracle and the CUDA path load the resulting dict."""
    import torch.nn.functional as F
    x = torch.stack([fundus_like_image(9000 + seed + i, size)["image"].float() for i in range(n_images)])
    x = x - torch.tensor([103.530, 116.280, 123.675]).reshape(1, 3, 1, 1)

    def conv_bn(x, name, stride=1, pad=0, relu=True):
        y = F.conv2d(x, sd[name + ".weight"], None, stride, pad)
        sd[name + ".norm.running_mean"] = y.mean(dim=(0, 2, 3))
        sd[name + ".norm.running_var"] = y.var(dim=(0, 2, 3), unbiased=False) + 1e-3
        scale = sd[name + ".norm.weight"] * (sd[name + ".norm.running_var"] + 1e-5).rsqrt()
        y = (y - sd[name + ".norm.running_mean"].reshape(1, -1, 1, 1)) * scale.reshape(1, -1, 1, 1) + sd[name + ".norm.bias"].reshape(1, -1, 1, 1)
        return F.relu(y) if relu else y

    p = "backbone.bottom_up."
    y = F.max_pool2d(conv_bn(x, p + "stem.conv1", 2, 3), 3, 2, 1)
    for stage, blocks, _, _, stride in RES_STAGES:
        for b in range(blocks):
            q = f"{p}{stage}.{b}."
            s = stride if b == 0 else 1
            out = conv_bn(y, q + "conv1", s)
            out = conv_bn(out, q + "conv2", 1, 1)
            out = conv_bn(out, q + "conv3", relu=False)
            sc = conv_bn(y, q + "shortcut", s, relu=False) if b == 0 else y
            y = F.relu(out + sc)
    return sd


_DET_CACHE = {}


def detector_state_calibrated(seed=0, num_classes=2):
    key = (seed, num_classes)
    if key not in _DET_CACHE:
        with torch.no_grad():
            _DET_CACHE[key] = calibrate_frozen_bn(detector_state(seed, num_classes), seed=seed)
    return {k: v.clone() for k, v in _DET_CACHE[key].items()}
