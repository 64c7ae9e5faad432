"""The reference-shaped calls on the GPU: the meta-architecture drives its sub-modules exactly as the reference does
(meta_arch/rcnn.py:219-226, 333-354), and batches of DIFFERENT image sizes (d2 ImageList padding, per-image clipping and
post-processing) agree with the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import _parity  # noqa: E402
from oracle import detector_port as dp  # noqa: E402  (checker only)
from ttdg_b200 import synth  # noqa: E402


def test_submodules_called_like_the_reference():
    m, sd_det, _, _ = _parity.build_model(2)
    m.train()
    ims = [synth.fundus_like_image(300 + i, 128) for i in range(3)]
    batched = [{"image": im["image"], "height": 128, "width": 128, "image_id": i} for i, im in enumerate(ims)]
    # reference rcnn.py:219, 226, 333-345
    images = m.preprocess_image(batched)
    features = m.backbone(images.tensor)
    assert list(features) == ["p2", "p3", "p4", "p5", "p6"]
    assert [tuple(f.shape) for f in features.values()] == [(3, 256, 128 // s, 128 // s) for s in (4, 8, 16, 32, 64)]
    proposals_rpn, losses = m.proposal_generator(images, features, None, compute_loss=False)
    assert losses == {} and len(proposals_rpn) == 3
    assert proposals_rpn[0].proposal_boxes.tensor.shape[1] == 4 and proposals_rpn[0].image_size == (128, 128)
    lg = proposals_rpn[0].objectness_logits
    assert len(lg) == len(proposals_rpn[0]) <= 1000 and bool((lg[:-1] >= lg[1:]).all())
    proposals_roih, _ = m.roi_heads(images, features, proposals_rpn, targets=None, compute_loss=False, branch="TTT")
    assert all(p.has("pred_boxes") and p.has("scores") and p.has("pred_classes") and not p.has("pred_masks") for p in proposals_roih)
    # the same detections as the fused internal path
    feats, props, dets = m._det[0].detect_ttt([b["image"] for b in batched])
    for p, d in zip(proposals_roih, dets):
        assert torch.equal(p.pred_boxes.tensor, d[0]) and torch.equal(p.scores, d[1]) and torch.equal(p.pred_classes, d[2])
    nodes, labels = m.graph_generator([f for f in features.values()], proposals_roih)
    assert len(nodes) == 3 and nodes[0].shape[1] == 256
    # eval mode: the mask branch is attached (roi_heads.py:112), soft 28 x 28 masks as mask_rcnn_inference leaves them
    m.eval()
    proposals, _ = m.proposal_generator(images, features, None)
    results, _ = m.roi_heads(images, features, proposals, None)
    assert results[0].pred_masks.shape[1:] == (1, 28, 28) and float(results[0].pred_masks.min()) >= 0.0
    out = m(batched)
    assert out[0]["instances"].pred_masks.dtype == torch.bool and out[0]["instances"].pred_masks.shape[1:] == (128, 128)


def test_mixed_size_batch_vs_oracle():
    """TEST.BATCH > 1 on a dataset whose images differ in size (Kvasir-SEG, ORIGA, BKAI in the reference's split table):
    ImageList padding to the batch maximum, RPN / box clipping to every image's own size, masks pasted at every image's own
    original 'height' / 'width'."""
    m, sd_det, _, _ = _parity.build_model(2)
    shapes = [(96, 128), (128, 160), (128, 128)]
    ims = []
    for i, (h, w) in enumerate(shapes):
        full = synth.fundus_like_image(700 + i, 160)
        ims.append({"image": full["image"][:, :h, :w].contiguous(), "gt": full["gt_masks"][:, :h, :w]})
    outs = [(2 * h, 2 * w) for h, w in shapes]                                  # the dataset dicts' original sizes
    batched = [{"image": im["image"], "height": o[0], "width": o[1], "image_id": i} for i, (im, o) in enumerate(zip(ims, outs))]
    m.eval()
    res = m(batched)
    ref = dp.inference(sd_det, [im["image"] for im in ims], outs)[0]
    for n, (o, r) in enumerate(zip(res, ref)):
        inst = o["instances"]
        assert inst.image_size == outs[n] and inst.pred_masks.shape[1:] == outs[n]
        assert len(inst) == len(r["scores"])
        b = inst.pred_boxes.tensor
        assert float(b[:, 2].max()) <= outs[n][1] and float(b[:, 3].max()) <= outs[n][0]       # clipped to the image's own size
        idx, ok = _parity.det_match((b, inst.scores, inst.pred_classes), (r["pred_boxes"], r["scores"], r["pred_classes"]), tol=0.2)
        assert float(ok.float().mean()) >= 0.9, float(ok.float().mean())         # free-running on tiny images (see tests/_parity.py)
        iou = _parity.iou(inst.pred_masks.cpu()[ok], r["pred_masks"][idx][ok])
        assert float(iou.mean()) > 0.93, float(iou.mean())       # small maps pasted at 2x: boundary pixels of tiny masks
    # the adaptation pass takes the same batch
    m.train()
    loss, _, _, feats = m(batched, branch="TTT")
    assert loss is not None and torch.isfinite(loss) and tuple(feats[0].shape[-2:]) == (32, 40)
    loss.backward()
