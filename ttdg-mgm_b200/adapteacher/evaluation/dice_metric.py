"""DiceEvaluator; mirrors reference adapteacher/evaluation/dice_metric.py:13-240: per predicted instance with score >=
``thres`` the best Dice / E-measure (IJCAI 2018) / S-measure (ICCV 2017) over the ground-truth masks of the same class,
x100, averaged over all kept instances.  Host-side metric code (numpy), like the reference; ground truth comes from a
list of dataset dicts (``image_id``, ``annotations`` with ``category_id`` and a binary ``mask`` / ``segmentation``
array - polygon / RLE decoding needs pycocotools and belongs to the data path)."""
import numpy as np
from scipy import ndimage


def dice(pred, gt):
    inter = np.logical_and(pred, gt).sum()
    return 2 * inter / (pred.sum() + gt.sum() + 1e-6)              # dice_metric.py:57-58


def enhanced_align(pred, gt):
    """E-measure of a binary prediction (dice_metric.py:110-143)."""
    pred = np.asarray(pred, dtype=np.float64)
    gt_b = np.asarray(gt, dtype=bool)
    th = min(2 * pred.mean(), 1.0)
    fm = (pred >= th).astype(np.float64)
    if gt_b.sum() == 0:
        enhanced = 1.0 - fm
    elif (~gt_b).sum() == 0:
        enhanced = fm
    else:
        g = gt_b.astype(np.float64)
        a_fm, a_gt = fm - fm.mean(), g - g.mean()
        align = 2.0 * (a_gt * a_fm) / (a_gt * a_gt + a_fm * a_fm + 1e-8)
        enhanced = (align + 1) ** 2 / 4
    h, w = gt_b.shape
    return enhanced.sum() / (h * w - 1 + 1e-8)


class Structure_measure:
    """S-measure (dice_metric.py:146-240), alpha = 0.5."""

    def __init__(self, alpha=0.5):
        self.alpha = alpha

    def get_score(self, pred, gt):
        pred = np.asarray(pred)
        gt = np.asarray(gt) > 0.5
        y = gt.mean()
        if y == 0:
            return 1 - pred.mean()
        if y == 1:
            return pred.mean()
        return self.alpha * self.object(pred, gt) + (1 - self.alpha) * self.region(pred, gt)

    @staticmethod
    def _s_object(x_in, sel):
        v = x_in[sel]
        x, sigma = v.mean(), v.std()
        return 2 * x / (x * x + 1 + sigma + 1e-8)

    def object(self, pred, gt):
        fg, bg = pred * gt, (1 - pred) * (1 - gt)
        u = gt.mean()
        return u * self._s_object(fg, gt) + (1 - u) * self._s_object(bg, np.logical_not(gt))

    @staticmethod
    def _ssim(a, b):
        b = np.float32(b)
        n = a.shape[0] * a.shape[1]
        x, y = a.mean(), b.mean()
        sxy = ((a - x) * (b - y)).sum() / (n - 1)
        alpha = 4 * x * y * sxy
        beta = (x * x + y * y) * (a.var() + b.var())
        if alpha != 0:
            return alpha / (beta + 1e-8)
        return 1 if beta == 0 else 0

    def region(self, pred, gt):
        cy, cx = ndimage.center_of_mass(gt)
        y, x = int(round(cy)) + 1, int(round(cx)) + 1
        h, w = gt.shape
        area = h * w
        quads = ((slice(0, y), slice(0, x), x * y), (slice(0, y), slice(x, w), y * (w - x)),
                 (slice(y, h), slice(0, x), (h - y) * x), (slice(y, h), slice(x, w), (h - y) * (w - x)))
        return sum(wt / area * self._ssim(pred[ys, xs], gt[ys, xs]) for ys, xs, wt in quads)


class DiceEvaluator:
    def __init__(self, dataset_name, thres, dataset_dicts=None):
        self.dataset_name = dataset_name
        self.dataset_dicts = dataset_dicts if dataset_dicts is not None else []
        self._by_id = {d["image_id"]: d for d in self.dataset_dicts}     # the reference scans linearly (:29-32)
        self.score_threshold = thres
        self.reset()

    def reset(self):
        self.dice_scores, self.ea_scores, self.sm_scores = [], [], []

    def process(self, inputs, outputs):
        for inp, out in zip(inputs, outputs):
            anns = self._by_id[inp["image_id"]]["annotations"]
            inst = out["instances"]
            masks = inst.pred_masks.cpu().numpy()                       # device -> host, as in the reference (:34-36)
            classes = inst.pred_classes.cpu().numpy()
            scores = inst.scores.cpu().numpy()
            keep = scores >= self.score_threshold
            gts = [(a["category_id"], np.asarray(a.get("mask", a.get("segmentation"))).astype(bool)) for a in anns]
            for c, m in zip(classes[keep], masks[keep]):
                best = [0.0, 0.0, 0.0]
                for gc, gm in gts:
                    if c == gc:
                        best[0] = max(best[0], dice(m, gm))
                        best[1] = max(best[1], enhanced_align(m, gm))
                        best[2] = max(best[2], Structure_measure().get_score(m, gm))
                self.dice_scores.append(best[0] * 100)
                self.ea_scores.append(best[1] * 100)
                self.sm_scores.append(best[2] * 100)

    def evaluate(self):
        return {"Dice Coefficient": np.mean(self.dice_scores), "Enhanced Alignment Metric": np.mean(self.ea_scores),
                "Structural Similarity Metric": np.mean(self.sm_scores)}
