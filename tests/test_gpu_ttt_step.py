"""The assembled hot path on the GPU: DAobjTwoStagePseudoLabGeneralizedRCNN.forward(branch='TTT') -> backward ->
fused SGD, then eval-mode inference with masks (adapteacher/engine/trainer.py:469-485)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import mgm_port  # noqa: E402  (checker only)
from ttdg_b200 import ops, synth  # noqa: E402
from ttdg_b200.optim import FlatSGD  # noqa: E402
from _traj import verify_trajectory  # noqa: E402


def build(num_classes=2):
    from adapteacher.modeling.meta_arch.rcnn import DAobjTwoStagePseudoLabGeneralizedRCNN
    m = DAobjTwoStagePseudoLabGeneralizedRCNN(num_classes).cuda()
    sd = dict(synth.detector_state_calibrated(0, num_classes))
    sd.update({"multi_matching_unsup." + k: v for k, v in synth.perturb_affinity_state(synth.mgm_unsup_state(0), 0).items()})
    sd["multi_matching_sup.U"] = synth.universe(0)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("D_img.") for k in missing), (missing, unexpected)
    return m


@pytest.mark.parametrize("size,batch", [(128, 3), (256, 4)])
def test_ttt_step_and_eval(size, batch):
    m = build()
    m.train()
    opt = FlatSGD(m.adapted_parameters(), lr=0.005, momentum=0.9, weight_decay=1e-4)
    assert opt.numel >= 26_971_137                                     # parameters that receive gradients
    inputs = [{"image": synth.fundus_like_image(200 + i, size)["image"], "height": size, "width": size, "image_id": i} for i in range(batch)]
    frozen_before = m.backbone.bottom_up.res2[1].conv2.weight.detach().clone()
    rpn_before = m.proposal_generator.rpn_head.conv.weight.detach().clone()
    w_before = m.backbone.bottom_up.res4[2].conv2.weight.detach().clone()
    loss, _, _, features = m(inputs, branch="TTT")
    assert loss is not None and torch.isfinite(loss)
    assert [tuple(f.shape) for f in features] == [(batch, 256, size // s, size // s) for s in (4, 8, 16, 32, 64)]
    opt.zero_grad()
    loss.backward()
    aux = m.multi_matching_unsup.last_aux
    # gradients reach exactly the adapted set
    for p in m.adapted_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all()
    assert float(opt.flat_g.abs().sum()) > 0
    assert float(m.backbone.bottom_up.res3[0].conv1.weight.grad.abs().sum()) > 0
    assert float(m.backbone.fpn_lateral5.weight.grad.abs().sum()) > 0
    assert m.proposal_generator.rpn_head.conv.weight.grad is None and m.multi_matching_sup.U.grad is None
    opt.step()
    assert torch.equal(m.backbone.bottom_up.res2[1].conv2.weight, frozen_before)
    assert torch.equal(m.proposal_generator.rpn_head.conv.weight, rpn_before)
    assert not torch.equal(m.backbone.bottom_up.res4[2].conv2.weight, w_before)
    # the node sampler picked exactly what the reference's sampler picks from OUR pyramid and detections
    sizes = aux["sizes"]
    assert len(sizes) == batch and all(1 <= n <= 95 for n in sizes)
    # every GA-GM iteration of the step is the oracle's step (chaotic system: tests/_traj.py)
    U2, info, trace, meta = ops.gagm_solve(aux["A"], aux["Wds"], aux["U0"], sizes, trace_cap=1300)
    assert torch.equal(U2, aux["U"])
    verify_trajectory(aux["A"].cpu(), aux["Wds"].cpu(), aux["U0"].cpu(), sizes, trace, meta, info.cpu().tolist())
    # eval pass with the adapted weights
    m.eval()
    out = m(inputs)
    assert len(out) == batch
    for o in out:
        inst = o["instances"]
        n = len(inst)
        assert 0 < n <= 100
        assert inst.pred_masks.shape == (n, size, size) and inst.pred_masks.dtype == torch.bool
        assert inst.pred_boxes.tensor.shape == (n, 4) and (inst.scores[:-1] >= inst.scores[1:]).all()
        assert int(inst.pred_classes.min()) >= 0 and int(inst.pred_classes.max()) <= 1


def test_sampler_on_detector_outputs_matches_oracle_sampler():
    m = build()
    m.train()
    size, batch = 256, 3
    images = [synth.fundus_like_image(300 + i, size)["image"] for i in range(batch)]
    with torch.no_grad():
        feats, props, dets = m._det[0].detect_ttt(images)
    feats_nchw = [f.permute(0, 3, 1, 2) for f in feats]
    nodes, labels = ops.sample_nodes(feats_nchw, [d[0] for d in dets], [d[2] for d in dets])
    rn, rl = mgm_port.sample_nodes([f.cpu().contiguous() for f in feats_nchw], [d[0].cpu() for d in dets], [d[2].cpu() for d in dets])
    assert len(nodes) == len(rn)
    for a, b, la, lb in zip(nodes, rn, labels, rl):
        assert torch.equal(a.cpu(), b) and torch.equal(la.cpu(), lb)


def test_ttt_single_image_is_skipped():
    """TEST.BATCH = 1: MGM3_unsup returns None for one graph (mgm:489-490) and the trainer skips the step (trainer.py:477-478)."""
    m = build()
    m.train()
    loss, _, _, _ = m([{"image": synth.fundus_like_image(1, 128)["image"]}], branch="TTT")
    assert loss is None


def test_baseline_trainer_test_end_to_end():
    """BaselineTrainer.test (trainer.py:430-529) with the real model: adapt on every batch, then Dice / E / S evaluation
    with the adapted weights; DICE_THRES 0 because a random-init detector has no confident detections."""
    from adapteacher.config import add_ateacher_config
    from adapteacher.engine.trainer import BaselineTrainer
    m = build()
    opt = FlatSGD(m.adapted_parameters(), lr=0.005, momentum=0.9, weight_decay=1e-4)
    size = 128
    ims = [synth.fundus_like_image(400 + i, size) for i in range(5)]
    dicts = [{"image_id": i, "annotations": [{"category_id": int(c), "mask": mk.numpy()} for c, mk in zip(im["gt_classes"], im["gt_masks"])]}
             for i, im in enumerate(ims)]
    inputs = [{"image": im["image"], "height": size, "width": size, "image_id": i} for i, im in enumerate(ims)]
    loader = [inputs[:3], inputs[3:]]                       # TEST.BATCH 3, drop_last False
    cfg = add_ateacher_config()
    cfg.DATASETS.TEST = ("REFUGE_test",)
    cfg.TEST.DICE_THRES = 0.0
    w0 = opt.flat_p.clone()
    res = BaselineTrainer.test(cfg, m, opt, data_loaders={"REFUGE_test": loader}, dataset_dicts={"REFUGE_test": dicts})
    assert opt.steps == 2 and not torch.equal(w0, opt.flat_p)
    for k in ("Dice Coefficient", "Enhanced Alignment Metric", "Structural Similarity Metric"):
        assert 0.0 <= res["REFUGE_test"][k] <= 100.0 and np.isfinite(res["REFUGE_mean"][k])


def test_config3_polyp_384_test_batch_5():
    """BASELINE.json configs[3] on one rank: 8 synthetic 384x384 1-class polyp-like images with TEST.BATCH = 5
    (drop_last False, data/build.py:146) -> one 5-graph and one 3-graph matching problem per pass over the shard, then the
    eval pass with the adapted weights.  The matching stage of the 5-graph problem is checked against the oracle port
    (differentiable half with the CUDA path's U; solver trajectory step by step)."""
    from adapteacher.config import add_ateacher_config
    from adapteacher.engine.trainer import BaselineTrainer
    m = build(num_classes=1)
    opt = FlatSGD(m.adapted_parameters(), lr=0.005, momentum=0.9, weight_decay=1e-4)
    size = 384
    ims = [synth.fundus_like_image(500 + i, size, polyp=True) for i in range(8)]
    inputs = [{"image": im["image"], "height": size, "width": size, "image_id": i} for i, im in enumerate(ims)]
    # the 5-graph problem on its own first (same weights as the trainer run will see)
    m.train()
    loss, _, _, feats = m(inputs[:5], branch="TTT")
    assert loss is not None and torch.isfinite(loss)
    assert [tuple(f.shape[-2:]) for f in feats] == [(size // s, size // s) for s in (4, 8, 16, 32, 64)]
    aux = m.multi_matching_unsup.last_aux
    sizes = aux["sizes"]
    assert len(sizes) == 5 and all(1 <= n <= 95 for n in sizes)
    U2, info, trace, meta = ops.gagm_solve(aux["A"], aux["Wds"], aux["U0"], sizes, trace_cap=1300)
    assert torch.equal(U2, aux["U"])
    verify_trajectory(aux["A"].cpu(), aux["Wds"].cpu(), aux["U0"].cpu(), sizes, trace, meta, info.cpu().tolist())
    dicts = [{"image_id": i, "annotations": [{"category_id": int(c), "mask": mk.numpy()} for c, mk in zip(im["gt_classes"], im["gt_masks"])]}
             for i, im in enumerate(ims)]
    cfg = add_ateacher_config()
    cfg.DATASETS.TEST = ("Polyp_test",)
    cfg.TEST.BATCH, cfg.TEST.DICE_THRES = 5, 0.0
    seen = []
    hook = m.multi_matching_unsup.register_forward_hook(lambda mod, a, o: seen.append(len(mod.last_aux["sizes"])))
    res = BaselineTrainer.test(cfg, m, opt, data_loaders={"Polyp_test": [inputs[:5], inputs[5:]]}, dataset_dicts={"Polyp_test": dicts})
    hook.remove()
    assert seen == [5, 3] and opt.steps == 2
    for k in ("Dice Coefficient", "Enhanced Alignment Metric", "Structural Similarity Metric"):
        assert 0.0 <= res["Polyp_test"][k] <= 100.0
