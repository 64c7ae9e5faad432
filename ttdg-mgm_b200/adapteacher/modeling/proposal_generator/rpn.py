"""PseudoLabRPN; mirrors reference adapteacher/modeling/proposal_generator/rpn.py:10-55: the d2 RPN with a ``compute_loss``
switch so that proposals can be produced in train mode without ground truth (the test-time-adaptation pass runs with
``model.training == True``, so the pre-NMS top-k is the TRAIN value, Base-RCNN-FPN.yaml:14-15).  Same call signature and
return value as the reference class; the head convolutions, anchor decoding, top-k and NMS run on libttdg_sm100.so
(``ttdg_b200.detector.RPN.predict``).  The loss branch (``label_and_sample_anchors`` / ``losses``, :43-48) belongs to source
training and is not on the test-time path."""
from ttdg_b200.detector import RPN, _nhwc
from ttdg_b200.registry import PROPOSAL_GENERATOR_REGISTRY
from ttdg_b200.structures import Boxes, Instances


@PROPOSAL_GENERATOR_REGISTRY.register()
class PseudoLabRPN(RPN):
    def forward(self, images, features, gt_instances=None, compute_loss=True, compute_val_loss=False):
        """images: ImageList (``image_sizes`` are what proposals are clipped to); features: {"p2": .., "p6": ..} as handed out
        by the backbone.  Returns (proposals: list[Instances(proposal_boxes, objectness_logits)], losses: {})."""
        if (self.training and compute_loss) or compute_val_loss:        # rpn.py:43-48
            raise NotImplementedError("RPN training losses are outside the test-time path (call with compute_loss=False)")
        feats = [_nhwc(features[f]) for f in self.in_features]
        boxes, logits, counts = self.predict_padded(feats, images.image_sizes, self.training)    # rpn.py:52-54 predict_proposals
        # Device-side selection keeps every image's proposals PADDED to POST_NMS_TOPK rows (padding: zero box, logit -inf) with
        # the real count on the device, so that no host round trip separates the RPN from the ROI heads; the padded batch rides
        # along for StandardROIHeadsPseudoLab (``num_valid()`` reads the count back if a caller needs it).
        proposals = []
        for n, size in enumerate(images.image_sizes):
            inst = Instances(size, proposal_boxes=Boxes(boxes[n]), objectness_logits=logits[n])
            inst._padded = (boxes, logits, counts, n)
            proposals.append(inst)
        return proposals, {}
