"""Plain-PyTorch CPU restatement of the detector half of the hot path: Detectron2 0.5 Mask R-CNN R50-FPN as
configured by configs/Base-RCNN-FPN.yaml + configs/test_segment.yaml and driven by
adapteacher/modeling/meta_arch/rcnn.py:154-357, proposal_generator/rpn.py:16-55, roi_heads/roi_heads.py:65-205.

TEST INFRASTRUCTURE (see oracle/__init__.py).  **Parity unpinned**: Detectron2 (requirements.txt:13, 0.5+cu111) is a
third-party dependency that is neither vendored under /root/reference nor installable in this image; its operators
are restated from SURVEY.md Appendix A ([recalled]) on top of torch / torchvision.ops (roi_align aligned=True, nms,
batched_nms).  Functional style over a state dict with d2's parameter names.  NCHW fp32 throughout.

Cross-check (tests/test_oracle_vs_torchvision.py): where torchvision implements the same published algorithm independently -
box decoding (BoxCoder.decode), proposal selection (RegionProposalNetwork.filter_proposals), FPN level assignment (LevelMapper),
box-head inference (RoIHeads.postprocess_detections, background class moved) - this restatement reproduces torchvision's outputs on
random head outputs (float64, mixed image sizes).  That pins the selection / decoding SEMANTICS to a second implementation; what stays
unpinned is what Detectron2 does differently from torchvision on purpose (unrounded anchors, aligned RoIAlign, grid_sample mask
pasting), restated from SURVEY Appendix A.
"""
import contextlib
import math

import torch
import torch.nn.functional as F
import torchvision.ops as tvo


@contextlib.contextmanager
def float64():
    """Run the restatement in float64 (state dict and images converted by the caller): the noise-free limit of the same
    algorithm, against which both the fp32 restatement and the CUDA path are measured in the parity tests."""
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        yield
    finally:
        torch.set_default_dtype(old)


def _dt():
    return torch.get_default_dtype()

PIXEL_MEAN = (103.530, 116.280, 123.675)      # d2 defaults, applied to whatever channel order arrives
STRIDES = (4, 8, 16, 32, 64)
ANCHOR_SIZES = (32, 64, 128, 256, 512)
ANCHOR_RATIOS = (0.5, 1.0, 2.0)
SCALE_CLAMP = math.log(1000.0 / 16)
STAGES = (("res2", 3, 1), ("res3", 4, 2), ("res4", 6, 2), ("res5", 3, 2))


# ---------------------------------------------------------------------------------------------- backbone
def frozen_bn(x, sd, name, eps=1e-5):
    scale = sd[name + ".weight"] * (sd[name + ".running_var"] + eps).rsqrt()
    bias = sd[name + ".bias"] - sd[name + ".running_mean"] * scale
    return x * scale.reshape(1, -1, 1, 1) + bias.reshape(1, -1, 1, 1)


def conv_bn(x, sd, name, stride=1, pad=0, relu=True):
    y = frozen_bn(F.conv2d(x, sd[name + ".weight"], None, stride, pad), sd, name + ".norm")
    return F.relu(y) if relu else y


def bottleneck(x, sd, q, stride, has_shortcut):
    out = conv_bn(x, sd, q + "conv1", stride=stride)                 # STRIDE_IN_1X1
    out = conv_bn(out, sd, q + "conv2", pad=1)
    out = conv_bn(out, sd, q + "conv3", relu=False)
    sc = conv_bn(x, sd, q + "shortcut", stride=stride, relu=False) if has_shortcut else x
    return F.relu(out + sc)


def preprocess(images_u8):
    """d2 preprocess_image + ImageList.from_tensors: list of uint8 C x H_i x W_i -> N x 3 x Hp x Wp float; every image is
    normalised first, then placed in the top-left corner of a ZERO canvas of the batch's maximum size rounded up to a
    multiple of 32 (size_divisibility)."""
    mean = torch.tensor(PIXEL_MEAN).reshape(3, 1, 1)
    ims = [im.to(_dt()) - mean for im in images_u8]                  # std = 1
    H, W = max(im.shape[-2] for im in ims), max(im.shape[-1] for im in ims)
    Hp, Wp = (H + 31) // 32 * 32, (W + 31) // 32 * 32
    return torch.stack([F.pad(im, (0, Wp - im.shape[-1], 0, Hp - im.shape[-2])) for im in ims])


def _sizes(image_size, n):
    """(h, w) for all images or one (h, w) per image -> list of n sizes."""
    if len(image_size) == 2 and not isinstance(image_size[0], (tuple, list)):
        return [tuple(image_size)] * n
    assert len(image_size) == n
    return [tuple(s) for s in image_size]


def backbone(sd, x):
    p = "backbone.bottom_up."
    with torch.no_grad():                                             # FREEZE_AT = 2: stem + res2 frozen
        y = conv_bn(x, sd, p + "stem.conv1", stride=2, pad=3)
        y = F.max_pool2d(y, 3, 2, 1)
        for b in range(3):
            y = bottleneck(y, sd, f"{p}res2.{b}.", 1, b == 0)
    res = {"res2": y}
    for stage, blocks, stride in STAGES[1:]:
        for b in range(blocks):
            y = bottleneck(y, sd, f"{p}{stage}.{b}.", stride if b == 0 else 1, b == 0)
        res[stage] = y
    # FPN top-down
    outs, prev = {}, None
    for lvl in (5, 4, 3, 2):
        lat = F.conv2d(res[f"res{lvl}"], sd[f"backbone.fpn_lateral{lvl}.weight"], sd[f"backbone.fpn_lateral{lvl}.bias"])
        prev = lat if prev is None else lat + F.interpolate(prev, scale_factor=2.0, mode="nearest")
        outs[lvl] = F.conv2d(prev, sd[f"backbone.fpn_output{lvl}.weight"], sd[f"backbone.fpn_output{lvl}.bias"], padding=1)
    p6 = F.max_pool2d(outs[5], kernel_size=1, stride=2, padding=0)
    return [outs[2], outs[3], outs[4], outs[5], p6]


# ---------------------------------------------------------------------------------------------- RPN
def cell_anchors():
    a = []
    for s in ANCHOR_SIZES:
        area = float(s) ** 2
        for r in ANCHOR_RATIOS:
            w = math.sqrt(area / r)
            h = r * w
            a.append([-w / 2.0, -h / 2.0, w / 2.0, h / 2.0])
    return torch.tensor(a)


def grid_anchors(h, w, stride):
    sx = torch.arange(0, w * stride, step=stride, dtype=_dt())
    sy = torch.arange(0, h * stride, step=stride, dtype=_dt())
    yy, xx = torch.meshgrid(sy, sx, indexing="ij")
    shifts = torch.stack((xx.reshape(-1), yy.reshape(-1), xx.reshape(-1), yy.reshape(-1)), dim=1)
    return (shifts.view(-1, 1, 4) + cell_anchors().view(1, -1, 4)).reshape(-1, 4)


def apply_deltas(deltas, boxes, weights):
    wx, wy, ww, wh = weights
    widths, heights = boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1]
    cx, cy = boxes[:, 0] + 0.5 * widths, boxes[:, 1] + 0.5 * heights
    dx, dy = deltas[:, 0::4] / wx, deltas[:, 1::4] / wy
    dw, dh = deltas[:, 2::4] / ww, deltas[:, 3::4] / wh
    dw, dh = torch.clamp(dw, max=SCALE_CLAMP), torch.clamp(dh, max=SCALE_CLAMP)
    pcx, pcy = dx * widths[:, None] + cx[:, None], dy * heights[:, None] + cy[:, None]
    pw, ph = torch.exp(dw) * widths[:, None], torch.exp(dh) * heights[:, None]
    out = torch.stack((pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph), dim=-1)
    return out.reshape(deltas.shape)


def clip_boxes(b, h, w):
    return torch.stack((b[:, 0].clamp(0, w), b[:, 1].clamp(0, h), b[:, 2].clamp(0, w), b[:, 3].clamp(0, h)), dim=1)


def rpn(sd, feats, image_size, training, nms_thresh=0.7, post_topk=1000):
    """StandardRPNHead + find_top_rpn_proposals.  Returns per image (boxes k x 4, logits k)."""
    r = "proposal_generator.rpn_head."
    pre_topk = 2000 if training else 1000                            # Base-RCNN-FPN.yaml:14-15 (TTT runs in train mode)
    N = feats[0].shape[0]
    top_boxes, top_scores, lvl_ids = [], [], []
    with torch.no_grad():
        for l, f in enumerate(feats):
            t = F.relu(F.conv2d(f, sd[r + "conv.weight"], sd[r + "conv.bias"], padding=1))
            logits = F.conv2d(t, sd[r + "objectness_logits.weight"], sd[r + "objectness_logits.bias"])
            deltas = F.conv2d(t, sd[r + "anchor_deltas.weight"], sd[r + "anchor_deltas.bias"])
            H, W = f.shape[-2:]
            logits = logits.permute(0, 2, 3, 1).flatten(1)                                   # N x HWA
            deltas = deltas.view(N, -1, 4, H, W).permute(0, 3, 4, 1, 2).flatten(1, -2)       # N x HWA x 4
            anchors = grid_anchors(H, W, STRIDES[l])
            k = min(logits.shape[1], pre_topk)
            sc, idx = logits.sort(descending=True, dim=1)
            sc, idx = sc[:, :k], idx[:, :k]
            props = torch.stack([apply_deltas(deltas[n][idx[n]], anchors[idx[n]], (1.0, 1.0, 1.0, 1.0)) for n in range(N)])
            top_boxes.append(props)
            top_scores.append(sc)
            lvl_ids.append(torch.full((k,), l, dtype=torch.int64))
        top_boxes, top_scores, lvl_ids = torch.cat(top_boxes, 1), torch.cat(top_scores, 1), torch.cat(lvl_ids)
        out = []
        sizes = _sizes(image_size, N)
        for n in range(N):
            b, s, lv = top_boxes[n], top_scores[n], lvl_ids
            valid = torch.isfinite(b).all(1) & torch.isfinite(s)
            b, s, lv = b[valid], s[valid], lv[valid]
            b = clip_boxes(b, sizes[n][0], sizes[n][1])
            keep = ((b[:, 2] - b[:, 0]) > 0) & ((b[:, 3] - b[:, 1]) > 0)
            b, s, lv = b[keep], s[keep], lv[keep]
            keep = tvo.batched_nms(b, s, lv, nms_thresh)[:post_topk]
            out.append((b[keep], s[keep]))
    return out


# ---------------------------------------------------------------------------------------------- ROI heads
def assign_levels(boxes, min_level=2, max_level=5, canonical_size=224, canonical_level=4):
    sizes = torch.sqrt((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1]))
    lv = torch.floor(canonical_level + torch.log2(sizes / canonical_size + 1e-8))
    return torch.clamp(lv, min=min_level, max=max_level).to(torch.int64) - min_level


def roi_pool(feats4, boxes_per_image, out_size):
    """ROIPooler with ROIAlignV2 (aligned, sampling_ratio 0) over p2..p5."""
    boxes = torch.cat(boxes_per_image)
    bidx = torch.cat([torch.full((len(b),), i, dtype=_dt()) for i, b in enumerate(boxes_per_image)])
    rois = torch.cat([bidx[:, None], boxes], dim=1)
    lv = assign_levels(boxes)
    C = feats4[0].shape[1]
    out = torch.zeros(len(boxes), C, out_size, out_size)
    for l, f in enumerate(feats4):
        sel = torch.nonzero(lv == l).squeeze(1)
        if len(sel):
            out[sel] = tvo.roi_align(f, rois[sel], out_size, spatial_scale=1.0 / STRIDES[l], sampling_ratio=0, aligned=True)
    return out


def box_head(sd, feats, proposals, image_size, score_thresh=0.05, nms_thresh=0.5, topk=100):
    """_forward_box + FastRCNNOutputLayers.inference.  proposals: list of (boxes, logits).  Returns per image
    (boxes, scores, classes)."""
    h = "roi_heads."
    with torch.no_grad():
        x = roi_pool(feats[:4], [p[0] for p in proposals], 7).flatten(1)
        x = F.relu(F.linear(x, sd[h + "box_head.fc1.weight"], sd[h + "box_head.fc1.bias"]))
        x = F.relu(F.linear(x, sd[h + "box_head.fc2.weight"], sd[h + "box_head.fc2.bias"]))
        scores = F.linear(x, sd[h + "box_predictor.cls_score.weight"], sd[h + "box_predictor.cls_score.bias"])
        deltas = F.linear(x, sd[h + "box_predictor.bbox_pred.weight"], sd[h + "box_predictor.bbox_pred.bias"])
        out, o = [], 0
        sizes = _sizes(image_size, len(proposals))
        for (boxes_p, _), image_size in zip(proposals, sizes):
            n = len(boxes_p)
            sc = F.softmax(scores[o:o + n], dim=-1)
            bx = apply_deltas(deltas[o:o + n], boxes_p, (10.0, 10.0, 5.0, 5.0))
            o += n
            valid = torch.isfinite(bx).all(1) & torch.isfinite(sc).all(1)
            bx, sc = bx[valid], sc[valid]
            sc = sc[:, :-1]
            K = bx.shape[1] // 4
            bx = clip_boxes(bx.reshape(-1, 4), image_size[0], image_size[1]).view(-1, K, 4)
            mask = sc > score_thresh
            inds = mask.nonzero()
            bsel, ssel = bx[mask], sc[mask]
            keep = tvo.batched_nms(bsel, ssel, inds[:, 1], nms_thresh)[:topk]
            out.append((bsel[keep], ssel[keep], inds[keep, 1]))
    return out


def mask_head(sd, feats, dets):
    """_forward_mask + mask_rcnn_inference: per image (R x 28 x 28) probabilities of the predicted class."""
    h = "roi_heads.mask_head."
    with torch.no_grad():
        x = roi_pool(feats[:4], [d[0] for d in dets], 14)
        for i in range(1, 5):
            x = F.relu(F.conv2d(x, sd[h + f"mask_fcn{i}.weight"], sd[h + f"mask_fcn{i}.bias"], padding=1))
        x = F.relu(F.conv_transpose2d(x, sd[h + "deconv.weight"], sd[h + "deconv.bias"], stride=2))
        logits = F.conv2d(x, sd[h + "predictor.weight"], sd[h + "predictor.bias"])
        cls = torch.cat([d[2] for d in dets])
        prob = logits[torch.arange(len(cls)), cls].sigmoid()
    return list(torch.split(prob, [len(d[0]) for d in dets]))


def paste_masks(probs, boxes, H, W, threshold=0.5):
    """detector_postprocess -> paste_masks_in_image (bilinear grid_sample, align_corners=False, >= threshold)."""
    n = len(boxes)
    if n == 0:
        return torch.zeros(0, H, W, dtype=torch.bool)
    x0, y0, x1, y1 = boxes[:, 0:1], boxes[:, 1:2], boxes[:, 2:3], boxes[:, 3:4]
    img_y = torch.arange(0, H, dtype=_dt()) + 0.5
    img_x = torch.arange(0, W, dtype=_dt()) + 0.5
    img_y = (img_y - y0) / (y1 - y0) * 2 - 1
    img_x = (img_x - x0) / (x1 - x0) * 2 - 1
    gx = img_x[:, None, :].expand(n, H, W)
    gy = img_y[:, :, None].expand(n, H, W)
    grid = torch.stack([gx, gy], dim=3)
    out = F.grid_sample(probs[:, None], grid, align_corners=False)
    return out[:, 0] >= threshold


def postprocess(dets, probs, image_size, out_size):
    res = []
    for (b, s, c), p, image_size, out_size in zip(dets, probs, _sizes(image_size, len(dets)), _sizes(out_size, len(dets))):
        sx, sy = out_size[1] / image_size[1], out_size[0] / image_size[0]
        b = b * torch.tensor([sx, sy, sx, sy])
        b = clip_boxes(b, out_size[0], out_size[1])
        keep = ((b[:, 2] - b[:, 0]) > 0) & ((b[:, 3] - b[:, 1]) > 0)
        b, s, c, p = b[keep], s[keep], c[keep], p[keep]
        res.append({"pred_boxes": b, "scores": s, "pred_classes": c, "pred_masks": paste_masks(p, b, out_size[0], out_size[1])})
    return res


# ---------------------------------------------------------------------------------------------- the two passes
def forward_ttt(sd, images_u8):
    """rcnn.py:331-357 up to the node sampler: features (grad-carrying), detections of the box head in TRAIN mode."""
    x = preprocess(images_u8)
    size = [tuple(im.shape[-2:]) for im in images_u8]
    feats = backbone(sd, x)
    props = rpn(sd, [f.detach() for f in feats], size, training=True)
    dets = box_head(sd, [f.detach() for f in feats], props, size)
    return feats, props, dets


def inference(sd, images_u8, out_sizes=None):
    """GeneralizedRCNN.inference (rcnn.py:181-182): eval-mode detections with pasted masks."""
    with torch.no_grad():
        x = preprocess(images_u8)
        size = [tuple(im.shape[-2:]) for im in images_u8]
        feats = backbone(sd, x)
        props = rpn(sd, feats, size, training=False)
        dets = box_head(sd, feats, props, size)
        probs = mask_head(sd, feats, dets)
        return postprocess(dets, probs, size, out_sizes or size), feats, props, dets, probs
