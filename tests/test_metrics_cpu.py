"""DiceEvaluator metric definitions against golden values computed by the reference's own functions
(oracle/gen_metric_golden.py), and the TTT driver's control flow on a stub model (CPU only)."""
import numpy as np
import torch

from adapteacher.evaluation.dice_metric import DiceEvaluator, Structure_measure, dice, enhanced_align
from adapteacher.engine.trainer import BaselineTrainer
from adapteacher.config import add_ateacher_config
from oracle.gen_metric_golden import cases
from ttdg_b200.structures import Boxes, Instances


def test_metric_definitions_match_reference(golden_dir):
    g = np.load(f"{golden_dir}/metrics.npz")
    for i, (pred, gt) in enumerate(cases()):
        np.testing.assert_allclose(dice(pred, gt), float(g[f"dice_{i}"]), rtol=1e-12)
        np.testing.assert_allclose(enhanced_align(pred, gt), float(g[f"ea_{i}"]), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(Structure_measure().get_score(pred, gt), float(g[f"sm_{i}"]), rtol=1e-6, atol=1e-9)


class _StubModel(torch.nn.Module):
    """Behaves like the meta-arch at its boundary: branch='TTT' -> (loss | None, [], [], feats); eval -> instances."""

    def __init__(self, gts):
        super().__init__()
        self.w = torch.nn.Parameter(torch.ones(1))
        self.gts = gts
        self.ttt_calls = 0

    def forward(self, inputs, branch="supervised"):
        if self.training:
            self.ttt_calls += 1
            if len(inputs) < 2:
                return None, [], [], []                     # single graph: skipped (trainer.py:477-478)
            return (self.w * 2).sum(), [], [], []
        out = []
        for d in inputs:
            m = torch.from_numpy(self.gts[d["image_id"]])[None]
            out.append({"instances": Instances(m.shape[-2:], pred_boxes=Boxes(torch.zeros(1, 4)), scores=torch.tensor([0.95]),
                                               pred_classes=torch.tensor([0]), pred_masks=m)})
        return out


def test_baseline_trainer_test_loop():
    pairs = cases()[:4]
    gts = {i: gt for i, (_, gt) in enumerate(pairs)}
    dicts = [{"image_id": i, "annotations": [{"category_id": 0, "mask": gt}]} for i, gt in gts.items()]
    batches = [[{"image_id": 0}, {"image_id": 1}], [{"image_id": 2}], [{"image_id": 3}]]
    cfg = add_ateacher_config()
    cfg.DATASETS.TEST = ("Fundus_a", "Fundus_b")
    model = _StubModel(gts)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    res = BaselineTrainer.test(cfg, model, opt, data_loaders={"Fundus_a": batches, "Fundus_b": batches},
                               dataset_dicts={"Fundus_a": dicts, "Fundus_b": dicts})
    assert model.ttt_calls == 6                              # every batch of both datasets is visited in pass 1
    np.testing.assert_allclose(float(model.w), 1.0 - 2 * 0.1 * 2)        # two non-skipped steps; weights carry over datasets
    for name in ("Fundus_a", "Fundus_b", "Fundus_mean"):
        np.testing.assert_allclose(res[name]["Dice Coefficient"], 100.0, rtol=1e-6)    # predictions == ground truth
    cfg.TEST.MIN_BATCH_NUM = 1
    model.ttt_calls = 0
    BaselineTrainer.test(cfg, model, opt, data_loaders={"Fundus_a": batches, "Fundus_b": batches},
                         dataset_dicts={"Fundus_a": dicts, "Fundus_b": dicts})
    assert model.ttt_calls == 2


def _counts_numpy(pred, gt):
    """What ttdg_mask_gt_stats / ttdg_mask_pair_counts compute, in numpy (the device kernels are checked against this)."""
    from scipy import ndimage
    h, w = gt.shape
    if gt.sum() > 0:
        cy, cx = ndimage.center_of_mass(gt)
        ys, xs = int(round(cy)) + 1, int(round(cx)) + 1
    else:
        ys = xs = 0
    c = np.zeros((4, 4), np.int64)
    for q, (sy, sx) in enumerate(((slice(0, ys), slice(0, xs)), (slice(0, ys), slice(xs, w)), (slice(ys, h), slice(0, xs)),
                                  (slice(ys, h), slice(xs, w)))):
        p, g = pred[sy, sx].astype(bool), gt[sy, sx].astype(bool)
        c[q] = [(p & g).sum(), (p & ~g).sum(), (~p & g).sum(), (~p & ~g).sum()]
    return c, (ys, xs)


def test_metrics_from_counts_match_reference(golden_dir):
    """The closed forms the on-device evaluator uses (pixel counts -> Dice / E-measure / S-measure) against the golden
    values of the reference's own array code, and against the array mirror on extra random shapes."""
    from adapteacher.evaluation.dice_metric import metrics_from_counts
    g = np.load(f"{golden_dir}/metrics.npz")
    for i, (pred, gt) in enumerate(cases()):
        c, split = _counts_numpy(pred, gt)
        d, e, s = metrics_from_counts(c, split, gt.shape)
        np.testing.assert_allclose(d, float(g[f"dice_{i}"]), rtol=1e-12)
        np.testing.assert_allclose(e, float(g[f"ea_{i}"]), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(s, float(g[f"sm_{i}"]), rtol=2e-6, atol=1e-9)
    rng = np.random.default_rng(5)
    for k in range(12):
        h, w = int(rng.integers(17, 90)), int(rng.integers(17, 90))
        yy, xx = np.mgrid[0:h, 0:w]
        gt = ((yy - rng.uniform(0, h)) ** 2 / rng.uniform(9, 400) + (xx - rng.uniform(0, w)) ** 2 / rng.uniform(9, 400)) <= 1
        pred = np.roll(gt, (int(rng.integers(-4, 5)), int(rng.integers(-4, 5))), (0, 1)) ^ (rng.random((h, w)) < 0.02)
        if gt.sum() == 0:
            continue
        c, split = _counts_numpy(pred, gt)
        d, e, s = metrics_from_counts(c, split, gt.shape)
        np.testing.assert_allclose(d, dice(pred, gt), rtol=1e-12)
        np.testing.assert_allclose(e, enhanced_align(pred, gt), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(s, Structure_measure().get_score(pred, gt), rtol=2e-6, atol=1e-9)


def test_metrics_from_counts_border_centroid_does_not_raise():
    """Ground truth touching the bottom-right border: the S-measure's quadrant split lands on the last row / column, a quadrant
    is empty (or one pixel), and the array code divides by zero the numpy way (nan with a warning).  The closed forms must do
    the same instead of raising ZeroDivisionError and aborting the evaluation."""
    import warnings
    from adapteacher.evaluation.dice_metric import metrics_from_counts
    h, w = 24, 31
    for gt_pix in ([(h - 1, w - 1)], [(h - 1, w - 2), (h - 1, w - 1)], [(0, 0)]):
        gt = np.zeros((h, w), bool)
        for p in gt_pix:
            gt[p] = True
        pred = np.zeros((h, w), bool)
        pred[h - 3:, w - 4:] = True
        c, split = _counts_numpy(pred, gt)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            d, e, s = metrics_from_counts(c, split, gt.shape)          # must not raise
            ref = Structure_measure().get_score(pred, gt)
        np.testing.assert_allclose(d, dice(pred, gt), rtol=1e-12)
        assert (np.isnan(s) and np.isnan(ref)) or abs(s - ref) <= 2e-6 * max(1.0, abs(ref)), (s, ref)


def test_overlapped_eval_uses_the_snapshot_and_keeps_the_order():
    """OverlappedEval (adapteacher/engine/trainer.py) on the host: batches queued at begin() are evaluated by window() / drain()
    with the weights of begin(), in order, whatever happens to the adapting model in between (no GPU: no second stream, no SM cap)."""
    from adapteacher.engine.trainer import OverlappedEval

    class Lin(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.ones(()))

        def forward(self, inputs):
            return [float(self.w) * d["x"] for d in inputs]

    class Rec:
        def reset(self):
            self.seen = []

        def process(self, inputs, outputs):
            self.seen += list(zip([d["x"] for d in inputs], outputs))

        def evaluate(self):
            return {"sum": sum(o for _, o in self.seen)}

    model, rec = Lin(), Rec()
    pipe = OverlappedEval(model)
    assert not pipe.active
    pipe.window()                                            # nothing queued: a no-op
    pipe.begin("a", [[{"x": 1.0}, {"x": 2.0}], [{"x": 3.0}], [{"x": 4.0}]], rec)
    with torch.no_grad():
        model.w.mul_(10.0)                                   # the next dataset's adaptation changes the model ...
    pipe.window()
    pipe.window()
    assert pipe.windows == 2 and rec.seen == [(1.0, 1.0), (2.0, 2.0), (3.0, 3.0)]       # ... the replica keeps the snapshot
    res, ev = pipe.drain()
    assert ev is rec and res == {"sum": 10.0} and rec.seen[-1] == (4.0, 4.0) and not pipe.active
    pipe.begin("b", [[{"x": 1.0}]], rec)                     # second snapshot: the adapted weights
    res, _ = pipe.drain()
    assert res == {"sum": 10.0} and rec.seen == [(1.0, 10.0)]
