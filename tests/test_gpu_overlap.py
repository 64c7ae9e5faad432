"""BaselineTrainer.test with the evaluation pass of a dataset running inside the GA-GM solver windows of the next dataset's
adaptation pass (adapteacher/engine/trainer.py OverlappedEval) - same results and same adapted weights as the strictly
sequential order of reference trainer.py:469-485."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _datasets(size=128):
    from ttdg_b200 import synth
    names = ("REFUGE_a", "REFUGE_b", "ORIGA_c")
    loaders, dicts = {}, {}
    for k, name in enumerate(names):
        ims = [synth.fundus_like_image(900 + 10 * k + i, size) for i in range(5)]
        dicts[name] = [{"image_id": i, "annotations": [{"category_id": int(c), "mask": mk.numpy()}
                                                        for c, mk in zip(im["gt_classes"], im["gt_masks"])]} for i, im in enumerate(ims)]
        inputs = [{"image": im["image"], "height": size, "width": size, "image_id": i} for i, im in enumerate(ims)]
        loaders[name] = [inputs[:3], inputs[3:]]            # TEST.BATCH 3, drop_last False
    return names, loaders, dicts


def test_overlapped_evaluation_gives_the_sequential_results():
    """The weight-gradient kernels accumulate with floating-point reductions whose order is not fixed, so two adaptation runs
    are not bit-identical; the check therefore replays every dataset's evaluation SEQUENTIALLY with exactly the weights the
    overlapped run snapshotted for it and asks for identical metrics."""
    import _parity
    from adapteacher.config import add_ateacher_config
    from adapteacher.engine import trainer as T
    from adapteacher.evaluation.dice_metric import DiceEvaluator
    from ttdg_b200 import _C, ops
    from ttdg_b200.optim import FlatSGD
    m, _, _, _ = _parity.build_model(2)
    opt = FlatSGD(m.adapted_parameters(), lr=0.005, momentum=0.9, weight_decay=1e-4, on_step=m.refresh_weight_copies)
    names, loaders, dicts = _datasets()
    cfg = add_ateacher_config()
    cfg.DATASETS.TEST = names
    cfg.TEST.DICE_THRES = 0.0
    assert cfg.TEST.OVERLAP_EVAL is True                    # the default schedule
    snapshots, in_windows = {}, []
    orig_begin, orig_window = T.OverlappedEval.begin, T.OverlappedEval.window

    def begin(self, name, loader, evaluator):
        snapshots[name] = opt.flat_p.clone()
        orig_begin(self, name, loader, evaluator)

    def window(self):
        before = self.windows
        orig_window(self)
        in_windows.append(self.windows - before)

    T.OverlappedEval.begin, T.OverlappedEval.window = begin, window
    try:
        res = T.BaselineTrainer.test(cfg, m, opt, data_loaders=loaders, dataset_dicts=dicts)
    finally:
        T.OverlappedEval.begin, T.OverlappedEval.window = orig_begin, orig_window
    torch.cuda.synchronize()
    assert opt.steps == 6 and ops.SOLVER_WINDOW_HOOK[0] is None
    assert _C.lib().ttdg_set_sm_limit(0) == 0                # the SM cap is restored after every window
    assert list(res.keys())[:3] == list(names)
    assert sum(in_windows) == 4                              # datasets a and b (2 batches each) ran beside the solvers of b and c
    for name in names:                                       # sequential replay with the snapshotted weights
        with torch.no_grad():
            opt.flat_p.copy_(snapshots[name])
        m.refresh_weight_copies()
        seq, _ = T.inference_on_dataset(m, loaders[name], DiceEvaluator(name, cfg.TEST.DICE_THRES, dicts[name]), cfg)
        assert set(seq) == set(res[name])
        for k, v in seq.items():
            assert np.array_equal(np.asarray(v), np.asarray(res[name][k]), equal_nan=True), (name, k, v, res[name][k])
