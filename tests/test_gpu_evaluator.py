"""On-device half of DiceEvaluator (csrc/evalmetric.cu): the contingency counts are integer work and must be bit-exact
against numpy; the evaluator fed CUDA masks must report what the host (reference-style numpy) path reports."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from adapteacher.evaluation.dice_metric import DiceEvaluator  # noqa: E402
from ttdg_b200 import ops  # noqa: E402
from ttdg_b200.structures import Boxes, Instances  # noqa: E402
from test_metrics_cpu import _counts_numpy  # noqa: E402


def _blobs(rng, n, h, w):
    yy, xx = np.mgrid[0:h, 0:w]
    out = np.zeros((n, h, w), bool)
    for k in range(n):
        out[k] = ((yy - rng.uniform(0, h)) ** 2 / rng.uniform(2, max(3.0, (h / 3) ** 2)) + (xx - rng.uniform(0, w)) ** 2 / rng.uniform(2, max(3.0, (w / 3) ** 2))) <= 1
    return out


@pytest.mark.parametrize("h,w", [(64, 64), (37, 53), (512, 512), (384, 384), (5, 7)])
def test_pair_counts_bit_exact(h, w):
    rng = np.random.default_rng(h * 1000 + w)
    gt = _blobs(rng, 3, h, w)
    gt[2] = False                                                    # an empty ground truth (split = 0, 0)
    pred = _blobs(rng, 5, h, w) ^ (rng.random((5, h, w)) < 0.01)
    pred[4] = False
    pairs = [(p, g) for p in range(5) for g in range(3)]
    gd, pd = torch.from_numpy(gt).cuda(), torch.from_numpy(pred).cuda()
    stats = ops.mask_gt_stats(gd)
    counts = ops.mask_pair_counts(pd, gd, pairs, stats).cpu().numpy()
    st = stats.cpu().numpy()
    for g in range(3):
        ys, xs = np.nonzero(gt[g])
        assert st[g, 0] == gt[g].sum() and st[g, 1] == ys.sum() and st[g, 2] == xs.sum()
    for (p, g), c in zip(pairs, counts):
        ref, split = _counts_numpy(pred[p], gt[g])
        assert (int(st[g, 3]), int(st[g, 4])) == split
        assert np.array_equal(c.reshape(4, 4), ref), (p, g)
    with pytest.raises(ValueError):
        ops.mask_pair_counts(pd, gd, [(5, 0)], stats)


def test_evaluator_on_device_matches_host_path():
    rng = np.random.default_rng(11)
    h = w = 128
    dicts, inputs, outs_dev, outs_host = [], [], [], []
    for i in range(3):
        gt = _blobs(rng, 2, h, w)
        dicts.append({"image_id": i, "annotations": [{"category_id": 0, "mask": gt[0]}, {"category_id": 1, "mask": gt[1]}]})
        n = 6
        masks = np.concatenate([np.roll(gt, (k - 2, 2 - k), (1, 2)) for k in range(3)])[:n] ^ (rng.random((n, h, w)) < 0.01)
        scores = torch.tensor([0.99, 0.95, 0.5, 0.93, 0.97, 0.2])
        classes = torch.tensor([0, 1, 0, 1, 0, 1])
        mk = lambda dev: Instances((h, w), pred_boxes=Boxes(torch.zeros(n, 4, device=dev)), scores=scores.to(dev),  # noqa: E731
                                   pred_classes=classes.to(dev), pred_masks=torch.from_numpy(masks).to(dev))
        inputs.append({"image_id": i})
        outs_dev.append({"instances": mk("cuda")})
        outs_host.append({"instances": mk("cpu")})
    ev_d, ev_h = DiceEvaluator("synthetic", 0.9, dicts), DiceEvaluator("synthetic", 0.9, dicts, on_device=False)
    ev_d.process(inputs, outs_dev)
    ev_h.process(inputs, outs_host)
    assert len(ev_d.dice_scores) == len(ev_h.dice_scores) == 12
    np.testing.assert_allclose(ev_d.dice_scores, ev_h.dice_scores, rtol=1e-12)
    np.testing.assert_allclose(ev_d.ea_scores, ev_h.ea_scores, rtol=1e-9)
    np.testing.assert_allclose(ev_d.sm_scores, ev_h.sm_scores, rtol=2e-6)
    rd, rh = ev_d.evaluate(), ev_h.evaluate()
    for k in rd:
        np.testing.assert_allclose(rd[k], rh[k], rtol=2e-6)
