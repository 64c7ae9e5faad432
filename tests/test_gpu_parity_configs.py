"""Parity of the assembled hot path ON BASELINE.json's OWN SHAPES at the benched precision (tf32x3 tensor-core mode):
configs[1] = 8 x 512 x 512, 2 classes; configs[3] = 384 x 384, 1 class, TEST.BATCH 5 (adapteacher/engine/trainer.py:469-485).
See tests/_parity.py for what is measured and why every quantity is stated free-running AND teacher-forced.  The measured
report of each run is written to gpurun_out/parity_<config>.json (copied to profiles/ per round)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import _parity  # noqa: E402
from ttdg_b200 import detector as det  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _dump(name, rep):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_%s.json" % name), "w") as f:
        json.dump(rep, f, indent=1)
    short = {k: v for k, v in rep.items() if k != "per_tensor"}
    print(json.dumps(short, indent=1))


@pytest.mark.parametrize("name", ["configs1", "configs3"])
def test_eval_pass_parity_on_baseline_shapes(name):
    assert det.CONV_MODE[0] == "tf32x3"                        # the benched conv math
    cfg = _parity.CONFIGS[name]
    m, sd_det, sd_mgm, U = _parity.build_model(cfg["num_classes"])
    rep = _parity.eval_parity(m, sd_det, cfg)
    _dump(name + "_eval", rep)
    # continuous stages, teacher-forced: tight
    assert rep["pyramid_rel_max"] < 5e-5
    bh = rep["box_head_forced_proposals"]
    assert bh["matched_frac"] >= 0.99 and bh["score_max_abs"] < 1e-4 and bh["box_max_abs_px"] < 1e-2, bh
    mb = rep["mask_branch_forced_detections"]
    assert mb["miou_delta"] < 1e-4, mb                          # north_star: segmentation mIoU within 1e-4 on the same detections
    assert mb["mask_iou_mean"] > 0.999, mb
    # free-running: each side follows its own top-k / NMS decisions.  The CUDA path must be no further from the float64
    # limit than ~ the fp32 restatement itself is (x2 + a floor for the small-sample noise of a handful of flips)
    fg, f32 = rep["free_running_gpu_vs_f64"], rep["free_running_fp32_vs_f64"]
    assert fg["matched_frac"] >= min(0.95, f32["matched_frac"] - 0.03), (fg, f32)
    assert fg["miou_delta_matched"] <= max(2.0 * f32["miou_delta_matched"], 2e-4) or fg["miou_delta_matched"] < 1e-4, (fg, f32)


@pytest.mark.parametrize("name", ["configs1", "configs3"])
def test_ttt_step_parity_on_baseline_shapes(name):
    assert det.CONV_MODE[0] == "tf32x3"
    cfg = _parity.CONFIGS[name]
    m, sd_det, sd_mgm, U = _parity.build_model(cfg["num_classes"])
    rep = _parity.ttt_parity(m, sd_det, sd_mgm, U, cfg)
    _dump(name + "_ttt", rep)
    assert len(rep["sizes"]) == cfg["batch"]
    # (a) loss with the matching result and the detections forced: continuous arithmetic only
    assert rep["loss_rel_gpu_vs_f64"] < max(3.0 * rep["loss_rel_fp32_vs_f64"], 2e-5), rep
    assert rep["A_max_abs_gpu_vs_fp32"] < 1e-5 and rep["Wds_max_abs_gpu_vs_fp32"] < 1e-4 and rep["U0_rel_max_gpu_vs_fp32"] < 1e-4, rep
    # (b) gradients of all 58 + 6 adapted tensors and the post-step weights vs the float64 limit: ReLU masks in the backward
    # flip under 1e-6 forward noise on EITHER side, so the yardstick is the fp32 restatement's own distance to that limit
    g, g32 = rep["grad_rel_l2_gpu_vs_f64"], rep["grad_rel_l2_fp32_vs_f64"]
    assert g["tensors"] >= 58
    assert rep["grad_bucket_rel_l2_gpu"] <= max(3.0 * rep["grad_bucket_rel_l2_fp32"], 1e-3), rep
    assert g["median"] <= max(3.0 * g32["median"], 1e-3) and g["max"] <= max(3.0 * g32["max"], 2e-2), (g, g32)
    u = rep["update_rel_l2_gpu"]
    assert u["median"] <= max(3.0 * g32["median"], 1e-3), u
    assert rep["weight_max_abs_gpu"] < 1e-5, rep["weight_max_abs_gpu"]
