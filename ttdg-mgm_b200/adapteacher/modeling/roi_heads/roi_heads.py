"""StandardROIHeadsPseudoLab; mirrors reference adapteacher/modeling/roi_heads/roi_heads.py:22-205 in inference form:
``forward_box`` = ``_forward_box`` + ``FastRCNNOutputLayers.inference`` (:173-205), ``forward_mask`` =
``forward_with_given_boxes`` (:112); ``branch == 'TTT'`` skips the mask branch (:109-110)."""
from ttdg_b200.detector import ROIHeads as StandardROIHeadsPseudoLab  # noqa: F401
