"""Cross-check of the detector oracle (oracle/detector_port.py - a restatement of Detectron2 0.5, which is not installable here)
against torchvision's INDEPENDENT implementation of the same published algorithms, wherever the two libraries' semantics coincide:
box decoding (Box2BoxTransform.apply_deltas = BoxCoder.decode), proposal selection (find_top_rpn_proposals =
RegionProposalNetwork.filter_proposals: per-level top-k, clip, drop empty, NMS per level, first post_nms_topk), FPN level assignment
(ROIPooler.assign_boxes_to_levels = LevelMapper) and box-head inference (fast_rcnn_inference = RoIHeads.postprocess_detections up to
the position of the background class).  Where they differ on purpose (torchvision rounds its cell anchors, pools with aligned = False
and pastes masks on an integer grid) nothing is compared.  CPU only."""
import math
import os
import sys

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import detector_port as dp  # noqa: E402

tv = pytest.importorskip("torchvision")
from torchvision.models.detection._utils import BoxCoder  # noqa: E402
from torchvision.models.detection.anchor_utils import AnchorGenerator  # noqa: E402
from torchvision.models.detection.roi_heads import RoIHeads  # noqa: E402
from torchvision.models.detection.rpn import RegionProposalNetwork, RPNHead  # noqa: E402
from torchvision.ops.poolers import LevelMapper  # noqa: E402


def _boxes(g, n, size=256.0):
    xy = torch.rand(n, 2, generator=g) * size * 0.8
    wh = torch.rand(n, 2, generator=g) * size * 0.4 + 2.0
    return torch.cat((xy, xy + wh), dim=1)


@pytest.mark.parametrize("weights", [(1.0, 1.0, 1.0, 1.0), (10.0, 10.0, 5.0, 5.0)])
def test_apply_deltas_is_torchvisions_box_decoding(weights):
    g = torch.Generator().manual_seed(3)
    boxes = _boxes(g, 500)
    deltas = torch.randn(500, 8, generator=g) * torch.tensor(weights).repeat(2) * 0.7
    deltas[0, 2] = 50.0                                                    # hits the exp clamp log(1000 / 16)
    ours = dp.apply_deltas(deltas, boxes, weights)
    ref = BoxCoder(weights, bbox_xform_clip=math.log(1000.0 / 16)).decode_single(deltas, boxes)
    assert torch.allclose(ours, ref.reshape(ours.shape), rtol=1e-6, atol=1e-4)


@pytest.mark.parametrize("training", [True, False])
def test_rpn_selection_is_torchvisions_filter_proposals(training):
    # float64 on both sides: torchvision orders by sigmoid(logit), whose fp32 rounding merges near-equal logits into ties
    with dp.float64():
        _rpn_selection_check(training)


def _rpn_selection_check(training):
    g = torch.Generator().manual_seed(5 + int(training))
    N, S = 2, 256
    C = 8
    sd = {"proposal_generator.rpn_head.conv.weight": torch.randn(C, C, 3, 3, generator=g) * 0.1,
          "proposal_generator.rpn_head.conv.bias": torch.randn(C, generator=g) * 0.1,
          "proposal_generator.rpn_head.objectness_logits.weight": torch.randn(15, C, 1, 1, generator=g),
          "proposal_generator.rpn_head.objectness_logits.bias": torch.randn(15, generator=g),
          "proposal_generator.rpn_head.anchor_deltas.weight": torch.randn(60, C, 1, 1, generator=g) * 0.3,
          "proposal_generator.rpn_head.anchor_deltas.bias": torch.randn(60, generator=g) * 0.1}
    sd = {k: v.double() for k, v in sd.items()}
    feats = [torch.randn(N, C, S // s, S // s, generator=g).double() for s in dp.STRIDES]
    sizes = [(256, 256), (200, 240)]
    ours = dp.rpn(sd, feats, sizes, training)
    # the same head outputs through torchvision's selection
    r = "proposal_generator.rpn_head."
    logits_l, deltas_l, anchors_l = [], [], []
    for l, f in enumerate(feats):
        t = F.relu(F.conv2d(f, sd[r + "conv.weight"], sd[r + "conv.bias"], padding=1))
        H, W = f.shape[-2:]
        lg = F.conv2d(t, sd[r + "objectness_logits.weight"], sd[r + "objectness_logits.bias"])
        dl = F.conv2d(t, sd[r + "anchor_deltas.weight"], sd[r + "anchor_deltas.bias"])
        logits_l.append(lg.permute(0, 2, 3, 1).flatten(1))
        deltas_l.append(dl.view(N, -1, 4, H, W).permute(0, 3, 4, 1, 2).reshape(N, -1, 4))
        anchors_l.append(dp.grid_anchors(H, W, dp.STRIDES[l]))             # d2's unrounded anchors (torchvision rounds its own)
    objectness = torch.cat(logits_l, 1)
    anchors = torch.cat(anchors_l)
    coder = BoxCoder((1.0, 1.0, 1.0, 1.0), bbox_xform_clip=math.log(1000.0 / 16))
    proposals = coder.decode(torch.cat(deltas_l, 1).reshape(-1, 4), [anchors] * N).view(N, -1, 4)
    tv_rpn = RegionProposalNetwork(AnchorGenerator(), RPNHead(C, 15), 0.7, 0.3, 256, 0.5, dict(training=2000, testing=1000),
                                   dict(training=1000, testing=1000), 0.7, score_thresh=0.0)
    tv_rpn.train(training)
    tv_rpn.min_size = 1e-12                                                # d2: min_box_size 0 = every box with w > 0 and h > 0
    boxes_tv, scores_tv = tv_rpn.filter_proposals(proposals, objectness.reshape(-1, 1), sizes, [t.shape[1] for t in logits_l])
    for n in range(N):
        b, s = ours[n]
        assert len(b) == len(boxes_tv[n]) > 100
        assert torch.allclose(b, boxes_tv[n], rtol=1e-9, atol=1e-8)
        assert torch.allclose(torch.sigmoid(s), scores_tv[n], rtol=1e-9, atol=1e-12)


def test_fpn_level_assignment_is_torchvisions_level_mapper():
    g = torch.Generator().manual_seed(9)
    boxes = torch.cat([_boxes(g, 400, 512.0), torch.tensor([[0.0, 0.0, 600.0, 600.0], [5.0, 5.0, 9.0, 9.0], [10.0, 20.0, 310.0, 330.0], [0.0, 0.0, 250.0, 400.0]])])
    ours = dp.assign_levels(boxes)
    ref = LevelMapper(2, 5, canonical_scale=224, canonical_level=4, eps=1e-6)([boxes])
    assert torch.equal(ours, ref)
    assert set(ours.tolist()) == {0, 1, 2, 3}


def test_box_head_inference_is_torchvisions_postprocess_detections():
    g = torch.Generator().manual_seed(13)
    K, C = 2, 4
    h = "roi_heads."
    sd = {h + "box_head.fc1.weight": torch.randn(32, C * 49, generator=g) * 0.1, h + "box_head.fc1.bias": torch.randn(32, generator=g) * 0.1,
          h + "box_head.fc2.weight": torch.randn(32, 32, generator=g) * 0.3, h + "box_head.fc2.bias": torch.randn(32, generator=g) * 0.1,
          h + "box_predictor.cls_score.weight": torch.randn(K + 1, 32, generator=g), h + "box_predictor.cls_score.bias": torch.randn(K + 1, generator=g),
          h + "box_predictor.bbox_pred.weight": torch.randn(K * 4, 32, generator=g) * 0.5, h + "box_predictor.bbox_pred.bias": torch.randn(K * 4, generator=g) * 0.1}
    feats = [torch.randn(2, C, 256 // s, 256 // s, generator=g) for s in dp.STRIDES]
    props = [(_boxes(g, 300), None), (_boxes(g, 250), None)]
    sizes = [(256, 256), (220, 256)]
    ours = dp.box_head(sd, feats, props, sizes)
    # the same logits / deltas through torchvision's post-processing (background class FIRST there, LAST in Detectron2)
    x = dp.roi_pool(feats[:4], [p[0] for p in props], 7).flatten(1)
    x = F.relu(F.linear(x, sd[h + "box_head.fc1.weight"], sd[h + "box_head.fc1.bias"]))
    x = F.relu(F.linear(x, sd[h + "box_head.fc2.weight"], sd[h + "box_head.fc2.bias"]))
    scores = F.linear(x, sd[h + "box_predictor.cls_score.weight"], sd[h + "box_predictor.cls_score.bias"])
    deltas = F.linear(x, sd[h + "box_predictor.bbox_pred.weight"], sd[h + "box_predictor.bbox_pred.bias"])
    tv_logits = torch.cat((scores[:, -1:], scores[:, :-1]), dim=1)
    tv_reg = torch.cat((torch.zeros(len(deltas), 4), deltas), dim=1)
    heads = RoIHeads(None, None, None, 0.5, 0.5, 512, 0.25, (10.0, 10.0, 5.0, 5.0), 0.05, 0.5, 100)
    heads.box_coder.bbox_xform_clip = math.log(1000.0 / 16)
    b_tv, s_tv, l_tv = heads.postprocess_detections(tv_logits, tv_reg, [p[0] for p in props], sizes)
    for n in range(2):
        b, s, c = ours[n]
        assert len(b) == len(b_tv[n]) > 10
        assert torch.allclose(b, b_tv[n], rtol=1e-6, atol=1e-4) and torch.allclose(s, s_tv[n], rtol=1e-6, atol=1e-7)
        assert torch.equal(c, l_tv[n] - 1)
