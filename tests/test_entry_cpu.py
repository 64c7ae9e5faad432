"""Host-side pieces around the hot path that run without a GPU: checkpoint I/O in Detectron2's format, the yaml / override
config loader of train_net.py, the test data path (sampler, batching, resize rule) and the polygon rasteriser."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tiny_model():
    from ttdg_b200 import detector
    m = torch.nn.Module()
    m.conv = detector.Conv2d(8, 16, 3, 1, 1, bias=False, norm=True)
    m.fc = detector.Conv2d(16, 5, kind="linear")
    return m


def test_checkpoint_roundtrip_d2_format(tmp_path):
    from adapteacher.checkpoint import DetectionCheckpointer
    torch.manual_seed(0)
    sd = {"conv.weight": torch.randn(16, 8, 3, 3), "conv.norm.weight": torch.rand(16) + 0.5, "conv.norm.bias": torch.randn(16),
          "conv.norm.running_mean": torch.randn(16), "conv.norm.running_var": torch.rand(16) + 0.5,
          "fc.weight": torch.randn(5, 16), "fc.bias": torch.randn(5)}
    m = _tiny_model()
    # a teacher / student ensemble file wrapped by DDP: the teacher half is what a single detector loads
    ens = {"module.modelTeacher." + k: v for k, v in sd.items()}
    ens.update({"module.modelStudent." + k: v + 1 for k, v in sd.items()})
    torch.save({"model": ens, "iteration": 9999}, tmp_path / "model_0009999.pth")
    ck = DetectionCheckpointer(m, save_dir=str(tmp_path / "out"))
    extra = ck.resume_or_load(str(tmp_path / "model_0009999.pth"), resume=True)
    assert extra.get("iteration") == 9999 and ck.last_incompatible.missing_keys == [] and ck.last_incompatible.incorrect_shapes == []
    out = m.state_dict()
    for k, v in sd.items():
        assert torch.equal(out[k], v), k                       # kernels' layout inside, Detectron2 names / shapes outside
    # wrong shapes are dropped and reported, not fatal (detection_checkpoint.py:80-87)
    bad = dict(sd)
    bad["fc.weight"] = torch.randn(7, 16)
    torch.save(bad, tmp_path / "bare.pth")                     # a bare state dict is accepted too
    m2 = _tiny_model()
    ck2 = DetectionCheckpointer(m2)
    ck2.load(str(tmp_path / "bare.pth"))
    assert ck2.last_incompatible.incorrect_shapes == [("fc.weight", (7, 16), (5, 16))] and "fc.weight" in ck2.last_incompatible.missing_keys
    # save -> last_checkpoint -> resume
    path = ck.save("model_adapted", iteration=3)
    assert os.path.exists(path) and ck.has_checkpoint()
    m3 = _tiny_model()
    assert DetectionCheckpointer(m3, save_dir=str(tmp_path / "out")).resume_or_load("", resume=True)["iteration"] == 3
    assert all(torch.equal(m3.state_dict()[k], v) for k, v in sd.items())
    with pytest.raises(NotImplementedError):
        ck.load(str(tmp_path / "x.pkl")) if (tmp_path / "x.pkl").write_bytes(b"") is not None else None


def test_config_yaml_base_and_overrides(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "ttdg-mgm_b200"))
    import train_net
    (tmp_path / "base.yaml").write_text("MODEL:\n  ROI_HEADS:\n    NUM_CLASSES: 80\n    NAME: StandardROIHeads\nSOLVER:\n  BASE_LR: 0.02\n")
    (tmp_path / "t.yaml").write_text('_BASE_: "./base.yaml"\nDATASETS:\n  TEST: ("REFUGE_test", "ORIGA_test")\nMODEL:\n  ROI_HEADS:\n'
                                     "    NUM_CLASSES: 2\nSOLVER:\n  BASE_LR: 0.005\nTEST:\n  BATCH: 4\n")
    args = train_net.default_argument_parser().parse_args(["--eval-only", "--config-file", str(tmp_path / "t.yaml"), "MODEL.WEIGHTS", "w.pth",
                                                            "TEST.BATCH", "8", "INPUT.MIN_SIZE_TEST", "512"])
    cfg = train_net.setup(args)
    assert cfg.DATASETS.TEST == ("REFUGE_test", "ORIGA_test") and cfg.MODEL.ROI_HEADS.NUM_CLASSES == 2
    assert cfg.MODEL.ROI_HEADS.NAME == "StandardROIHeads" and cfg.SOLVER.BASE_LR == 0.005 and cfg.TEST.BATCH == 8
    assert cfg.MODEL.WEIGHTS == "w.pth" and cfg.INPUT.MIN_SIZE_TEST == 512 and cfg.TEST.TTT is True and cfg.TEST.DICE_THRES == 0.9
    shipped = train_net.setup(train_net.default_argument_parser().parse_args(
        ["--eval-only", "--config", os.path.join(ROOT, "ttdg-mgm_b200", "configs", "test_segment_synthetic.yaml")]))
    assert shipped.DATASETS.TEST == ("synthetic_fundus_16",) and shipped.TEST.BATCH == 8
    with pytest.raises(NotImplementedError):
        train_net.main(train_net.default_argument_parser().parse_args([]))


def test_test_loader_sharding_and_batches():
    from adapteacher.config import add_ateacher_config
    from adapteacher.data import build_detection_test_loader, DatasetMapper, InferenceSampler
    cfg = add_ateacher_config()
    cfg.TEST.BATCH = 4
    seen = []
    for rank in range(3):
        loader = build_detection_test_loader(cfg, "synthetic_polyp_10_64", rank=rank, world_size=3)
        batches = list(loader)
        assert list(loader) and [len(b) for b in batches] == [[4], [4], [2]][rank]      # drop_last = False, re-iterable
        for b in batches:
            for d in b:
                assert d["image"].dtype == torch.uint8 and tuple(d["image"].shape) == (3, 64, 64) and "annotations" not in d
                seen.append(d["image_id"])
    assert seen == list(range(10))                               # contiguous shards, every image exactly once
    assert list(InferenceSampler(7, 1, 2)) == [4, 5, 6]
    cfg.TEST.TTT = False
    assert len(list(build_detection_test_loader(cfg, "synthetic_polyp_10_64", rank=0, world_size=1))) == 10
    m = DatasetMapper(cfg)
    assert m._resize_shape(480, 640) == (800, 1067) and m._resize_shape(1000, 3000) == (444, 1333)


def test_polygon_rasteriser_agrees_with_fill():
    """pycocotools is absent (parity unpinned); the restated rleFrPoly must at least agree with an independent polygon
    fill up to boundary pixels, and reproduce axis-aligned rectangles exactly."""
    import cv2
    from adapteacher.data.build import polygon_to_mask, segmentation_to_mask
    m = polygon_to_mask([10, 5, 30, 5, 30, 20, 10, 20], 40, 50)
    ref = np.zeros((40, 50), bool)
    ref[5:20, 10:30] = True
    assert np.array_equal(m, ref)
    rng = np.random.default_rng(3)
    for _ in range(10):
        h, w = int(rng.integers(40, 120)), int(rng.integers(40, 120))
        ang = np.sort(rng.uniform(0, 2 * np.pi, 9))
        r = rng.uniform(0.3, 0.45, 9) * min(h, w)
        pts = np.stack([w / 2 + r * np.cos(ang), h / 2 + r * np.sin(ang)], 1)
        m = segmentation_to_mask([pts.reshape(-1).tolist()], h, w)
        ref = np.zeros((h, w), np.uint8)
        cv2.fillPoly(ref, [np.round(pts).astype(np.int32)], 1)
        inter, union = (m & (ref > 0)).sum(), (m | (ref > 0)).sum()
        assert inter / union > 0.85, inter / union
