"""StandardROIHeadsPseudoLab; mirrors reference adapteacher/modeling/roi_heads/roi_heads.py:22-205 in inference form, with
the reference's call signature: ``forward(images, features, proposals, targets, compute_loss, branch)`` -> ``_forward_box``
+ ``FastRCNNOutputLayers.inference`` (:173-205); ``branch == 'TTT'`` returns the box predictions and skips the mask branch
(:109-110); any other branch continues with ``forward_with_given_boxes`` (:112).  RoIAlign, the FC / mask-head
convolutions, box decoding and NMS run on libttdg_sm100.so (``ttdg_b200.detector.ROIHeads``).  The loss branches
(``label_and_sample_proposals``, :80-102) belong to source training and are not on the test-time path."""
import torch

from ttdg_b200.detector import ROIHeads, _nhwc
from ttdg_b200.registry import ROI_HEADS_REGISTRY
from ttdg_b200.structures import Boxes, Instances


@ROI_HEADS_REGISTRY.register()
class StandardROIHeadsPseudoLab(ROIHeads):
    in_features = ("p2", "p3", "p4", "p5")

    def forward(self, images, features, proposals, targets=None, compute_loss=True, branch="", compute_val_loss=False):
        del images                                                       # roi_heads.py:76 (sizes travel with the proposals)
        if (self.training and compute_loss) or compute_val_loss:         # roi_heads.py:77-102
            raise NotImplementedError("ROI-head training losses are outside the test-time path (call with compute_loss=False)")
        del targets
        feats = [_nhwc(features[f]) for f in self.in_features]
        sizes = [p.image_size for p in proposals]
        pad = [getattr(p, "_padded", None) for p in proposals]
        if all(q is not None and q[0] is pad[0][0] and q[3] == i for i, q in enumerate(pad)) and len(pad) == pad[0][0].shape[0]:
            # proposals straight from PseudoLabRPN: the padded batch + device-side counts, one host read for the detections
            b, s, c, counts = self.forward_box_padded(feats, pad[0][0], pad[0][2], sizes)
            cnt = counts.cpu().tolist()
            dets = [(b[i, :n], s[i, :n], c[i, :n]) for i, n in enumerate(cnt)]
        else:
            dets = self.forward_box(feats, [(p.proposal_boxes.tensor, p.objectness_logits) for p in proposals], sizes)
        pred_instances = [Instances(size, pred_boxes=Boxes(b), scores=s, pred_classes=c) for (b, s, c), size in zip(dets, sizes)]
        if branch == 'TTT':                                              # roi_heads.py:109-110
            return pred_instances, None
        return self.forward_with_given_boxes(features, pred_instances, feats), None

    @torch.no_grad()
    def forward_with_given_boxes(self, features, instances, _feats=None):
        """d2 StandardROIHeads.forward_with_given_boxes -> _forward_mask -> mask_rcnn_inference: attaches ``pred_masks`` =
        sigmoid of the predicted class's 28 x 28 logits (R x 1 x 28 x 28).  The raw logits ride along (private attributes)
        so that detector_postprocess can paste the whole batch with one fused kernel."""
        feats = _feats if _feats is not None else [_nhwc(features[f]) for f in self.in_features]
        dets = [(i.pred_boxes.tensor, i.scores, i.pred_classes) for i in instances]
        logits = self.mask_logits(feats, dets)
        o = 0
        for inst in instances:
            n = len(inst.pred_boxes)
            if logits is None:
                inst.pred_masks = torch.zeros(0, 1, 28, 28, device=inst.scores.device)
            else:
                lg = logits[o:o + n]
                sel = torch.gather(lg, 3, inst.pred_classes.view(n, 1, 1, 1).expand(n, 28, 28, 1))
                inst.pred_masks = sel.permute(0, 3, 1, 2).sigmoid()
            inst._mask_logits = (logits, o, n)
            o += n
        return instances
