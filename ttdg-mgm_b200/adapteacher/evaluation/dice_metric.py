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


# ---------------------------------------------------------------------------------------------- closed forms from counts
def metrics_from_counts(counts, split, shape):
    """Dice, E-measure and S-measure of one BINARY (prediction, ground truth) pair from the contingency tables the
    device produces (``ttdg_mask_pair_counts``): ``counts`` int[4 quadrants TL, TR, BL, BR][n11, n10, n01, n00]
    (prediction first), ``split`` = (y, x) of the S-measure's quadrant split, ``shape`` = (H, W).  Every expression
    follows the numpy code above value by value; a sum over pixels becomes count x value (so results agree with the
    array code to summation-order noise: 1e-12 for Dice / E-measure, 1e-6 for the S-measure whose reference SSIM runs in
    float32, dice_metric.py:204-205)."""
    c = np.asarray(counts, dtype=np.int64).reshape(4, 4)
    h, w = shape
    n = h * w
    a, b, cc, d = (int(v) for v in c.sum(0))                 # n11, n10 (pred only), n01 (gt only), n00
    n_pred, n_gt = a + b, a + cc
    dice_v = 2 * a / (n_pred + n_gt + 1e-6)
    # ---- E-measure: fm = pred >= min(2 * mean, 1); an empty prediction thresholds at 0 and becomes all ones
    if n_pred == 0:
        f11, f10, f01, f00 = cc, d, 0, 0                     # (fm, gt) contingency
    else:
        f11, f10, f01, f00 = a, b, cc, d
    if n_gt == 0:
        e_sum = float(f01 + f00)                             # enhanced = 1 - fm
    elif n - n_gt == 0:
        e_sum = float(f11 + f10)                             # enhanced = fm
    else:
        m_fm, m_gt = (f11 + f10) / n, n_gt / n
        e_sum = 0.0
        for cnt, fv, gv in ((f11, 1.0, 1.0), (f10, 1.0, 0.0), (f01, 0.0, 1.0), (f00, 0.0, 0.0)):
            a_fm, a_gt = fv - m_fm, gv - m_gt
            align = 2.0 * (a_gt * a_fm) / (a_gt * a_gt + a_fm * a_fm + 1e-8)
            e_sum += cnt * ((align + 1) ** 2 / 4)
    ea_v = e_sum / (n - 1 + 1e-8)
    # ---- S-measure
    y_mean = n_gt / n
    if n_gt == 0:
        sm_v = 1 - n_pred / n
    elif n_gt == n:
        sm_v = n_pred / n
    else:
        def s_object(ones, total):
            x = ones / total
            sigma = np.sqrt(max(x * (1 - x), 0.0))           # population std of a 0/1 sample
            return 2 * x / (x * x + 1 + sigma + 1e-8)
        obj = y_mean * s_object(a, n_gt) + (1 - y_mean) * s_object(d, n - n_gt)
        ys, xs = split
        sizes = (ys * xs, ys * (w - xs), (h - ys) * xs, (h - ys) * (w - xs))
        reg = 0.0
        for q in range(4):
            q11, q10, q01, q00 = (int(v) for v in c[q])
            nq = q11 + q10 + q01 + q00
            assert nq == sizes[q], "quadrant sizes disagree with the split"
            # an empty quadrant (ground-truth centroid on the last row / column) or a one-pixel quadrant divides by zero: numpy
            # semantics (nan / inf with a warning, as in the reference's array code), never a Python ZeroDivisionError
            with np.errstate(all="ignore"):
                nqf = np.float64(nq)
                x = np.float64(q11 + q10) / nqf                   # mean of the boolean prediction (float64)
                y32 = np.float32(q11 + q01) / np.float32(nq)      # mean of the float32 ground truth (float32 division)
                b1, b0 = np.float32(1) - y32, np.float32(0) - y32
                sxy = (q11 * ((1 - x) * b1) + q10 * ((1 - x) * b0) + q01 * ((0 - x) * b1) + q00 * ((0 - x) * b0)) / np.float64(nq - 1)
                var_a = ((q11 + q10) * (1 - x) ** 2 + (q01 + q00) * (0 - x) ** 2) / nqf
                var_b = np.float32(((q11 + q01) * np.float64(b1) ** 2 + (q10 + q00) * np.float64(b0) ** 2) / nqf)
                alpha = 4 * x * y32 * sxy
                beta = (x * x + y32 * y32) * (var_a + var_b)
                ssim = alpha / (beta + 1e-8) if alpha != 0 else (1 if beta == 0 else 0)
                reg += sizes[q] / n * ssim
        sm_v = 0.5 * obj + 0.5 * reg
    return dice_v, ea_v, float(sm_v)


def _gt_mask(ann, h, w):
    """Ground-truth mask of one annotation: a binary 'mask' array (synthetic data) or COCO polygons rasterised like
    pycocotools (dice_metric.py:94-108 -> adapteacher/data/build.py:polygon_to_mask)."""
    if "mask" in ann:
        return np.asarray(ann["mask"]).astype(bool)
    from adapteacher.data.build import segmentation_to_mask
    return segmentation_to_mask(ann["segmentation"], h, w)


class DiceEvaluator:
    def __init__(self, dataset_name, thres, dataset_dicts=None, on_device=None):
        self.dataset_name = dataset_name
        if dataset_dicts is None:                               # the reference: DatasetCatalog.get(dataset_name) (:16)
            from adapteacher.data.build import DatasetCatalog
            dataset_dicts = DatasetCatalog.get(dataset_name) if dataset_name in DatasetCatalog or str(dataset_name).startswith("synthetic_") else []
        self.dataset_dicts = dataset_dicts
        self._by_id = {d["image_id"]: d for d in self.dataset_dicts}     # the reference scans linearly (:29-32)
        self.score_threshold = thres
        # on_device: None = automatically when the predicted masks are CUDA tensors; the masks then stay on the GPU and only
        # 16 counts per (prediction, ground truth) pair come back (ttdg_b200.ops.mask_pair_counts)
        self.on_device = on_device
        self._gt_cache = {}
        self.reset()

    def _process_on_device(self, image_id, anns, inst, hw=None):
        from ttdg_b200 import ops
        import torch
        masks = inst.pred_masks
        dev = masks.device
        if image_id not in self._gt_cache:                    # ground truth: uploaded once per image, with its centroid split
            h, w = hw if hw is not None else tuple(masks.shape[-2:])
            gt = np.stack([_gt_mask(a, h, w).astype(np.uint8) for a in anns]) if anns else np.zeros((0, h, w), np.uint8)
            gt_d = torch.from_numpy(gt).to(dev)
            self._gt_cache[image_id] = (gt_d, ops.mask_gt_stats(gt_d), [a["category_id"] for a in anns])
        gt_d, stats_d, gt_cls = self._gt_cache[image_id]
        keep = (inst.scores >= self.score_threshold).nonzero().flatten().cpu().tolist()
        classes = inst.pred_classes.cpu().tolist()
        pairs = [(p, g) for p in keep for g, gc in enumerate(gt_cls) if classes[p] == gc]
        counts = ops.mask_pair_counts(masks, gt_d, pairs, stats_d).cpu().numpy() if pairs else np.zeros((0, 16), np.int64)
        stats = stats_d.cpu().numpy()
        best = {p: [0.0, 0.0, 0.0] for p in keep}
        for (p, g), cnt in zip(pairs, counts):
            v = metrics_from_counts(cnt, (int(stats[g, 3]), int(stats[g, 4])), tuple(masks.shape[-2:]))
            best[p] = [max(x, y) for x, y in zip(best[p], v)]
        for p in keep:
            self.dice_scores.append(best[p][0] * 100)
            self.ea_scores.append(best[p][1] * 100)
            self.sm_scores.append(best[p][2] * 100)

    def reset(self):
        self.dice_scores, self.ea_scores, self.sm_scores = [], [], []

    def process(self, inputs, outputs):
        for inp, out in zip(inputs, outputs):
            anns = self._by_id[inp["image_id"]]["annotations"]
            inst = out["instances"]
            if self.on_device or (self.on_device is None and inst.pred_masks.is_cuda):
                self._process_on_device(inp["image_id"], anns, inst, (inp["height"], inp["width"]) if "height" in inp else None)
                continue
            masks = inst.pred_masks.cpu().numpy()                       # device -> host, as in the reference (:34-36)
            classes = inst.pred_classes.cpu().numpy()
            scores = inst.scores.cpu().numpy()
            keep = scores >= self.score_threshold
            h, w = (inp["height"], inp["width"]) if "height" in inp else masks.shape[-2:]
            gts = [(a["category_id"], _gt_mask(a, h, w)) for a in anns]
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
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:      # image shards -> all ranks' scores
            parts = [None] * dist.get_world_size()
            dist.all_gather_object(parts, (self.dice_scores, self.ea_scores, self.sm_scores))
            self.dice_scores, self.ea_scores, self.sm_scores = ([v for p in parts for v in p[k]] for k in range(3))
        return {"Dice Coefficient": np.mean(self.dice_scores), "Enhanced Alignment Metric": np.mean(self.ea_scores),
                "Structural Similarity Metric": np.mean(self.sm_scores)}
