"""d2 ``detector_postprocess`` (detectron2/modeling/postprocessing.py, reached from GeneralizedRCNN._postprocess): rescale the
boxes of one image's Instances from the network input size to (output_height, output_width), clip, drop empty boxes, paste
the 28 x 28 masks into the image (``paste_masks_in_image``: bilinear, threshold 0.5) -> ``pred_masks`` bool R x H x W."""
from .structures import Boxes, Instances


def _paste(roi_heads, instances, out_sizes):
    lg = [getattr(i, "_mask_logits", None) for i in instances]
    if any(l is None for l in lg):
        raise ValueError("detector_postprocess needs instances produced by StandardROIHeadsPseudoLab.forward_with_given_boxes")
    batch = lg[0][0]
    contiguous = all(l[0] is batch for l in lg) and all(lg[k][1] == lg[k - 1][1] + lg[k - 1][2] for k in range(1, len(lg))) \
        and lg[0][1] == 0
    dets = [(i.pred_boxes.tensor, i.scores, i.pred_classes) for i in instances]
    sizes = [i.image_size for i in instances]
    if contiguous:                                        # the whole batch: one paste launch (or one per image for mixed sizes)
        return roi_heads.paste(batch, dets, out_sizes, sizes)
    return [roi_heads.paste(None if l[0] is None else l[0][l[1]:l[1] + l[2]], [d], [o], [s])[0]
            for l, d, o, s in zip(lg, dets, out_sizes, sizes)]


def detector_postprocess_batch(roi_heads, instances, out_sizes):
    res = _paste(roi_heads, instances, out_sizes)
    return [Instances(tuple(o), pred_boxes=Boxes(r["pred_boxes"]), scores=r["scores"], pred_classes=r["pred_classes"],
                      pred_masks=r["pred_masks"]) for r, o in zip(res, out_sizes)]


def detector_postprocess(roi_heads, results, output_height, output_width):
    return detector_postprocess_batch(roi_heads, [results], [(int(output_height), int(output_width))])[0]
