"""DAobjTwoStagePseudoLabGeneralizedRCNN; mirrors reference adapteacher/modeling/meta_arch/rcnn.py:67-357 for the
two branches the test-time path uses: ``branch='TTT'`` (:331-357) and eval-mode inference (:181-182 ->
d2 GeneralizedRCNN.inference).  Same sub-module names (``backbone``, ``proposal_generator``, ``roi_heads``,
``D_img``, ``graph_generator``, ``multi_matching_sup``, ``multi_matching_unsup``), state-dict keys and - since round 2 - the
same CALLS between them: ``preprocess_image`` -> ImageList, ``backbone(images.tensor)`` -> {"p2".."p6"},
``proposal_generator(images, features, None, compute_loss=False)``, ``roi_heads(images, features, proposals, targets=None,
compute_loss=False, branch=branch)``, ``graph_generator(features, proposals_roih)``, ``multi_matching_unsup(nodes, labels, U)``.
The three detector sub-modules are looked up by name in the registries (``from_config``), as Detectron2 does."""
import torch
import torch.nn as nn

from ttdg_b200.detector import Backbone, MaskRCNN
from ttdg_b200.postprocess import detector_postprocess_batch
from ttdg_b200.registry import BACKBONE_REGISTRY, META_ARCH_REGISTRY, PROPOSAL_GENERATOR_REGISTRY, ROI_HEADS_REGISTRY
from ttdg_b200.structures import ImageList
from adapteacher.modeling.GModule.build_graph import PrototypeComputation
from adapteacher.modeling.GModule.multi_graph_matching import MGM3_unsup, U_sup
from adapteacher.modeling.proposal_generator.rpn import PseudoLabRPN
from adapteacher.modeling.roi_heads.roi_heads import StandardROIHeadsPseudoLab

BACKBONE_REGISTRY.register(Backbone, name="build_resnet_fpn_backbone")      # Base-RCNN-FPN.yaml:4


class FCDiscriminator_img(nn.Module):
    """Image-level domain discriminator (rcnn.py:26-65): only its parameters are kept, for checkpoint compatibility;
    it is used by the adversarial training branches, never at test time."""

    def __init__(self, num_classes, ndf1=256, ndf2=128):
        super().__init__()
        self.conv1 = nn.Conv2d(num_classes, ndf1, kernel_size=3, padding=1)
        self.conv2 = nn.Conv2d(ndf1, ndf2, kernel_size=3, padding=1)
        self.conv3 = nn.Conv2d(ndf2, ndf2, kernel_size=3, padding=1)
        self.classifier = nn.Conv2d(ndf2, 1, kernel_size=3, padding=1)


@META_ARCH_REGISTRY.register()
class DAobjTwoStagePseudoLabGeneralizedRCNN(nn.Module):
    def __init__(self, num_classes=2, dis_type="p2", backbone_cls=None, proposal_generator_cls=None, roi_heads_cls=None):
        super().__init__()
        det = MaskRCNN(num_classes, backbone_cls or Backbone, proposal_generator_cls or PseudoLabRPN,
                       roi_heads_cls or StandardROIHeadsPseudoLab)
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
        self.last_ttt = None

    @classmethod
    def from_config(cls, cfg):
        """rcnn.py:125-138: the sub-modules named by the yaml (test_segment.yaml:9-19, Base-RCNN-FPN.yaml:4)."""
        mdl = cfg.MODEL
        bb = getattr(getattr(mdl, "BACKBONE", None), "NAME", "build_resnet_fpn_backbone")
        pg = getattr(getattr(mdl, "PROPOSAL_GENERATOR", None), "NAME", "PseudoLabRPN")
        rh = getattr(mdl.ROI_HEADS, "NAME", "StandardROIHeadsPseudoLab")
        return cls(mdl.ROI_HEADS.NUM_CLASSES, getattr(getattr(cfg, "SEMISUPNET", None), "DIS_TYPE", "p2"),
                   BACKBONE_REGISTRY.get(bb), PROPOSAL_GENERATOR_REGISTRY.get(pg), ROI_HEADS_REGISTRY.get(rh))

    @property
    def device(self):
        return self.multi_matching_sup.U.device

    def train(self, mode=True):
        """nn.Module.train re-walks the module tree (~300 sub-modules) on every call; the test-time loop flips the mode twice
        per batch (trainer.py:469-485), so the flat module list is cached once - every sub-module still sees the flag."""
        mods = self.__dict__.get("_all_modules")
        if mods is None:
            mods = list(self.modules())
            self.__dict__["_all_modules"] = mods
        for m in mods:
            m.training = mode
        return self

    def adapted_parameters(self):
        """Everything that receives a gradient in the TTT step: res3-res5, FPN and the affinity layer, in the order
        the backward pass completes the gradients."""
        return [p for g in self.adapted_parameter_groups() for p in g]

    def adapted_parameter_groups(self):
        """Gradient buckets for FlatSGD(buckets=[len(g) for g in groups]): [affinity + FPN + res5, res4, res3]."""
        g = self._det[0].adapted_parameter_groups()
        return [list(self.multi_matching_unsup.node_affinity.parameters()) + g[0], g[1], g[2]]

    def refresh_weight_copies(self):
        self._det[0].refresh_weight_copies()

    def preprocess_image(self, batched_inputs):
        """rcnn.py:219 -> d2 preprocess_image: normalise, pad to a common size divisible by 32, keep every image's own size."""
        x, sizes = self._det[0].preprocess_image([b["image"] for b in batched_inputs])
        return ImageList(x, sizes)

    def forward(self, batched_inputs, branch="supervised", given_proposals=None, val_mode=False):
        if not self.training and not val_mode:              # rcnn.py:181-182
            return self.inference(batched_inputs)
        if branch != "TTT":
            raise NotImplementedError("only branch='TTT' and eval-mode inference are on the test-time path (SURVEY 8)")
        images = self.preprocess_image(batched_inputs)      # rcnn.py:219
        features = self.backbone(images.tensor)             # rcnn.py:226
        proposals_rpn, _ = self.proposal_generator(images, features, None, compute_loss=False)              # rcnn.py:333-335
        proposals_roih, ROI_predictions = self.roi_heads(images, features, proposals_rpn, targets=None,
                                                         compute_loss=False, branch=branch)                 # rcnn.py:338-345
        self.last_ttt = {"proposals": proposals_rpn,
                         "detections": [(p.pred_boxes.tensor, p.scores, p.pred_classes) for p in proposals_roih]}   # parity tests
        features = [feat[1] for feat in features.items()]   # rcnn.py:351
        nodes, labels = self.graph_generator(features, proposals_roih)      # rcnn.py:352
        loss = self.multi_matching_unsup(nodes, labels, self.multi_matching_sup.U)          # rcnn.py:353-354
        return loss, [], [], features

    @torch.no_grad()
    def inference(self, batched_inputs, detected_instances=None, do_postprocess=True):
        """d2 GeneralizedRCNN.inference: every image is post-processed to its OWN 'height' / 'width' (the original size in the
        dataset dict; default: the network input size)."""
        if detected_instances is not None:
            raise NotImplementedError("inference with given boxes is not on the test-time path")
        images = self.preprocess_image(batched_inputs)
        features = self.backbone(images.tensor)
        proposals, _ = self.proposal_generator(images, features, None)
        results, _ = self.roi_heads(images, features, proposals, None)
        if not do_postprocess:
            return results
        out_sizes = [(int(b.get("height", s[0])), int(b.get("width", s[1]))) for b, s in zip(batched_inputs, images.image_sizes)]
        return [{"instances": r} for r in detector_postprocess_batch(self.roi_heads, results, out_sizes)]


@META_ARCH_REGISTRY.register()
class TwoStagePseudoLabGeneralizedRCNN(DAobjTwoStagePseudoLabGeneralizedRCNN):
    """The reference's second meta-architecture (rcnn.py:423-494, imported by train_net.py:15): same detector without the
    domain-adaptive training branches.  At test time only eval-mode inference applies."""
