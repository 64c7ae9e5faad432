"""Graph-node sampler; mirrors reference adapteacher/modeling/GModule/build_graph.py:11-250
(``PrototypeComputation``): FCOS-style assignment of every FPN location to the smallest containing predicted
box within the level's size range, then every step-th positive per level becomes a graph node."""
import torch

from ttdg_b200 import ops


class PrototypeComputation(object):
    def __init__(self, num_cls, sample_dist):
        self.num_class = num_cls
        self.num_class_fgbg = num_cls + 1
        self.class_cond_nodes = sample_dist
        self.sample_dist = sample_dist

    def __call__(self, features, targets):
        """features: 5 maps B x C x H_l x W_l (p2..p6); targets: list of Instances-like objects exposing
        ``pred_boxes.tensor`` / ``pred_classes`` (TTT, build_graph.py:80-85) or ``gt_boxes.tensor`` / ``gt_classes``."""
        boxes, classes = [], []
        for t in targets:
            fields = getattr(t, "_fields", {})
            if "pred_boxes" in fields or hasattr(t, "pred_boxes"):
                boxes.append(t.pred_boxes.tensor)
                classes.append(t.pred_classes)
            else:
                boxes.append(t.gt_boxes.tensor)
                classes.append(t.gt_classes)
        return ops.sample_nodes(list(features), boxes, classes, self.sample_dist)
