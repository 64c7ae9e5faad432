"""DAobjTwoStagePseudoLabGeneralizedRCNN; mirrors reference adapteacher/modeling/meta_arch/rcnn.py:67-357 for the
two branches the test-time path uses: ``branch='TTT'`` (:331-357) and eval-mode inference (:181-182 ->
d2 GeneralizedRCNN.inference).  Same sub-module names (``backbone``, ``proposal_generator``, ``roi_heads``,
``D_img``, ``graph_generator``, ``multi_matching_sup``, ``multi_matching_unsup``) and state-dict keys."""
import torch
import torch.nn as nn

from ttdg_b200.detector import MaskRCNN
from ttdg_b200.structures import Boxes, Instances
from adapteacher.modeling.GModule.build_graph import PrototypeComputation
from adapteacher.modeling.GModule.multi_graph_matching import MGM3_unsup, U_sup


class FCDiscriminator_img(nn.Module):
    """Image-level domain discriminator (rcnn.py:26-65): only its parameters are kept, for checkpoint compatibility;
    it is used by the adversarial training branches, never at test time."""

    def __init__(self, num_classes, ndf1=256, ndf2=128):
        super().__init__()
        self.conv1 = nn.Conv2d(num_classes, ndf1, kernel_size=3, padding=1)
        self.conv2 = nn.Conv2d(ndf1, ndf2, kernel_size=3, padding=1)
        self.conv3 = nn.Conv2d(ndf2, ndf2, kernel_size=3, padding=1)
        self.classifier = nn.Conv2d(ndf2, 1, kernel_size=3, padding=1)


class DAobjTwoStagePseudoLabGeneralizedRCNN(nn.Module):
    def __init__(self, num_classes=2, dis_type="p2"):
        super().__init__()
        det = MaskRCNN(num_classes)
        self.backbone = det.backbone
        self.proposal_generator = det.proposal_generator
        self.roi_heads = det.roi_heads
        self._det = [det]                                   # not a registered sub-module: shares the three above
        self.num_classes = num_classes
        self.dis_type = dis_type
        self.D_img = FCDiscriminator_img(256)               # rcnn.py:113
        self.graph_generator = PrototypeComputation(num_classes, 10)        # rcnn.py:115
        self.multi_matching_sup = U_sup(num_classes, 32)                    # rcnn.py:116
        self.multi_matching_unsup = MGM3_unsup(num_classes, 32)

    @property
    def device(self):
        return self.multi_matching_sup.U.device

    def train(self, mode=True):
        """nn.Module.train walks ~300 sub-modules; the test-time loop flips the mode twice per batch (trainer.py:469-485) and
        only two modules read the flag: this one (forward dispatch) and the attention of the matching head (dropout)."""
        if getattr(self, "_mode_initialised", False):
            self.training = mode
            for m in self.multi_matching_unsup.modules():
                m.training = mode
            return self
        self._mode_initialised = True
        return super().train(mode)

    def adapted_parameters(self):
        """Everything that receives a gradient in the TTT step: res3-res5, FPN and the affinity layer."""
        return self._det[0].adapted_parameters() + list(self.multi_matching_unsup.node_affinity.parameters())

    def forward(self, batched_inputs, branch="supervised", given_proposals=None, val_mode=False):
        images = [x["image"] for x in batched_inputs]
        if not self.training and not val_mode:              # rcnn.py:181-182
            return self.inference(batched_inputs)
        if branch != "TTT":
            raise NotImplementedError("only branch='TTT' and eval-mode inference are on the test-time path (SURVEY 8)")
        det = self._det[0]
        feats, props, dets = det.detect_ttt(images)         # rcnn.py:219-226, 333-345
        self.last_ttt = {"proposals": props, "detections": dets}            # (views; read by the parity tests only)
        size = tuple(images[0].shape[-2:])
        proposals_roih = [Instances(size, pred_boxes=Boxes(b), scores=s, pred_classes=c) for b, s, c in dets]
        features = [f.permute(0, 3, 1, 2) for f in feats]   # NCHW views of the NHWC pyramid (rcnn.py:351)
        nodes, labels = self.graph_generator(features, proposals_roih)      # rcnn.py:352
        loss = self.multi_matching_unsup(nodes, labels, self.multi_matching_sup.U)          # rcnn.py:353-354
        return loss, [], [], features

    @torch.no_grad()
    def inference(self, batched_inputs):
        images = [x["image"] for x in batched_inputs]
        h = batched_inputs[0].get("height", images[0].shape[-2])
        w = batched_inputs[0].get("width", images[0].shape[-1])
        results, _, _, _ = self._det[0].inference(images, (h, w))
        out = []
        for r in results:
            out.append({"instances": Instances((h, w), pred_boxes=Boxes(r["pred_boxes"]), scores=r["scores"],
                                               pred_classes=r["pred_classes"], pred_masks=r["pred_masks"])})
        return out
