"""add_ateacher_config; mirrors the keys of reference adapteacher/config.py:5-64 the test-time path reads, on a plain
namespace tree (yacs / Detectron2's CfgNode are not dependencies here)."""
from types import SimpleNamespace


def add_ateacher_config(cfg=None):
    cfg = cfg or SimpleNamespace()
    cfg.TEST = getattr(cfg, "TEST", SimpleNamespace())
    cfg.TEST.TTT = True                      # config.py:15
    cfg.TEST.BATCH = 1
    cfg.TEST.MIN_BATCH_NUM = None
    cfg.TEST.DICE_THRES = 0.9
    cfg.TEST.DRAW = False
    cfg.DATASETS = getattr(cfg, "DATASETS", SimpleNamespace(TEST=()))
    cfg.SEMISUPNET = getattr(cfg, "SEMISUPNET", SimpleNamespace(Trainer="baseline", DIS_TYPE="p2", BBOX_THRESHOLD=0.8))
    cfg.SOLVER = getattr(cfg, "SOLVER", SimpleNamespace(BASE_LR=0.005, MOMENTUM=0.9, WEIGHT_DECAY=1e-4))
    return cfg
