"""train_net.py --eval-only end to end on the GPU: yaml config -> model from a Detectron2-format .pth -> test-time
adaptation pass + evaluation pass over the synthetic dataset through the real loader / on-device evaluator."""
import json
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_train_net_eval_only_synthetic(tmp_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "ttdg-mgm_b200"))
    import bench
    import train_net
    from adapteacher.checkpoint import DetectionCheckpointer
    torch.save({"model": {"module." + k: v for k, v in bench.full_state(bench.CONFIGS[1]).items()}}, tmp_path / "model_final.pth")
    out_dir = tmp_path / "out"
    args = train_net.default_argument_parser().parse_args(
        ["--eval-only", "--config", os.path.join(ROOT, "ttdg-mgm_b200", "configs", "test_segment_synthetic.yaml"),
         "MODEL.WEIGHTS", str(tmp_path / "model_final.pth"), "OUTPUT_DIR", str(out_dir), "DATASETS.TEST", '("synthetic_fundus_6_256",)',
         "TEST.BATCH", "3"])
    res = train_net.main(args)
    assert set(res) == {"synthetic_fundus_6_256", "synthetic_mean"}
    for k in ("Dice Coefficient", "Enhanced Alignment Metric", "Structural Similarity Metric"):
        v = res["synthetic_fundus_6_256"][k]
        assert np.isfinite(v) and 0.0 <= v <= 100.0
        assert res["synthetic_mean"][k] == v
    lines = open(out_dir / "result_ap.txt").read().splitlines()
    assert lines[0].startswith("loading data from: ") and "synthetic_mean" in json.loads(lines[1])
    # the adapted weights can be written back in the same format and differ from the loaded ones (two TTT steps ran)
    cfg = train_net.setup(args)
    model = train_net.Trainer.build_model(cfg)
    ck = DetectionCheckpointer(model, save_dir=str(out_dir))
    ck.load(str(tmp_path / "model_final.pth"))
    assert ck.last_incompatible.incorrect_shapes == [] and not [k for k in ck.last_incompatible.missing_keys if "multi_matching_sup" not in k and "D_img" not in k]
    assert os.path.exists(ck.save("model_adapted"))
