"""Config objects of the entry point: a yacs-style ``CfgNode`` (attribute access, ``merge_from_file`` with ``_BASE_``,
``merge_from_list``, ``freeze`` / ``defrost``, ``clone``, ``dump``), ``get_cfg()`` with the Detectron2 defaults the test-time
path reads, and ``add_ateacher_config(cfg)`` (reference adapteacher/config.py:5-64) - so that ``train_net.setup`` is the
reference's own five lines (reference train_net.py:23-33).  yacs / Detectron2 are not dependencies here.

Difference from yacs: keys that have no default are ACCEPTED by ``merge_from_file`` / ``merge_from_list`` (the reference's yaml
files set many Detectron2 keys of the training path that this package has no use for); a frozen node still rejects writes."""
import ast
import copy
import os

import yaml


def _literal(v):
    if isinstance(v, str):
        try:
            return ast.literal_eval(v)                          # yacs: '("a", "b")' -> tuple, '0.5' -> float
        except (ValueError, SyntaxError):
            return v
    return v


class CfgNode(dict):
    _FROZEN = "__frozen__"

    def __init__(self, init_dict=None):
        super().__init__()
        object.__setattr__(self, CfgNode._FROZEN, False)
        for k, v in (init_dict or {}).items():
            dict.__setitem__(self, k, CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v)

    # ---- attribute access
    def __getattr__(self, name):
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if self.is_frozen():
            raise AttributeError("Attempted to set {} to {}, but CfgNode is immutable".format(name, value))
        self[name] = CfgNode(value) if isinstance(value, dict) and not isinstance(value, CfgNode) else value

    def __setitem__(self, name, value):
        if self.is_frozen():
            raise AttributeError("Attempted to set {} to {}, but CfgNode is immutable".format(name, value))
        dict.__setitem__(self, name, value)

    # ---- yacs API
    def is_frozen(self):
        return self.__dict__.get(CfgNode._FROZEN, False)

    def _set_frozen(self, flag):
        object.__setattr__(self, CfgNode._FROZEN, flag)
        for v in self.values():
            if isinstance(v, CfgNode):
                v._set_frozen(flag)

    def freeze(self):
        self._set_frozen(True)

    def defrost(self):
        self._set_frozen(False)

    def clone(self):
        out = CfgNode({k: (v.clone() if isinstance(v, CfgNode) else copy.deepcopy(v)) for k, v in self.items()})
        out._set_frozen(self.is_frozen())
        return out

    def _merge_dict(self, d):
        for k, v in d.items():
            if isinstance(v, dict):
                sub = self.get(k)
                if not isinstance(sub, CfgNode):
                    sub = CfgNode()
                    self[k] = sub
                sub._merge_dict(v)
            else:
                self[k] = _literal(v)

    @staticmethod
    def load_yaml_with_base(path):
        with open(path) as f:
            d = yaml.safe_load(f) or {}
        base = d.pop("_BASE_", None)
        out = CfgNode.load_yaml_with_base(os.path.join(os.path.dirname(path), base)) if base else {}

        def deep(a, b):
            for k, v in b.items():
                if isinstance(v, dict) and isinstance(a.get(k), dict):
                    deep(a[k], v)
                else:
                    a[k] = v
        deep(out, d)
        return out

    def merge_from_file(self, cfg_filename):
        self._merge_dict(CfgNode.load_yaml_with_base(cfg_filename))

    def merge_from_other_cfg(self, other):
        self._merge_dict(other)

    def merge_from_list(self, cfg_list):
        cfg_list = list(cfg_list)
        if len(cfg_list) % 2:
            raise ValueError("Override list has odd length: {}; it must be a list of pairs".format(cfg_list))
        for k, v in zip(cfg_list[::2], cfg_list[1::2]):
            node = self
            *path, leaf = k.split(".")
            for pth in path:
                if not isinstance(node.get(pth), CfgNode):
                    node[pth] = CfgNode()
                node = node[pth]
            node[leaf] = _literal(v)

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, CfgNode) else v) for k, v in self.items()}

    def dump(self, **kwargs):
        def plain(v):
            if isinstance(v, dict):
                return {k: plain(x) for k, x in v.items()}
            return list(v) if isinstance(v, tuple) else v
        return yaml.safe_dump(plain(self.to_dict()), **kwargs)


def get_cfg():
    """The slice of detectron2.config.get_cfg() this package reads (d2 config/defaults.py)."""
    return CfgNode({
        "MODEL": {"WEIGHTS": "", "META_ARCHITECTURE": "GeneralizedRCNN", "DEVICE": "cuda", "MASK_ON": False,
                  "BACKBONE": {"NAME": "build_resnet_backbone", "FREEZE_AT": 2},
                  "PROPOSAL_GENERATOR": {"NAME": "RPN"},
                  "RPN": {}, "ROI_HEADS": {"NAME": "Res5ROIHeads", "NUM_CLASSES": 80}},
        "INPUT": {"FORMAT": "BGR", "MIN_SIZE_TEST": 800, "MAX_SIZE_TEST": 1333},
        "DATASETS": {"TRAIN": (), "TEST": ()},
        "DATALOADER": {"NUM_WORKERS": 4},
        "SOLVER": {"BASE_LR": 0.001, "MOMENTUM": 0.9, "WEIGHT_DECAY": 0.0001},
        "TEST": {},
        "OUTPUT_DIR": "./output",
    })


def add_ateacher_config(cfg=None):
    """reference adapteacher/config.py:5-64 (the keys it ADDS to a Detectron2 config), applied to ``cfg`` in place."""
    cfg = cfg if cfg is not None else get_cfg()
    for section in ("TEST", "MODEL", "SOLVER", "DATASETS", "DATALOADER"):
        if not isinstance(cfg.get(section), CfgNode):
            cfg[section] = CfgNode()
    for sub in ("RPN", "ROI_HEADS"):
        if not isinstance(cfg.MODEL.get(sub), CfgNode):
            cfg.MODEL[sub] = CfgNode()
    cfg.TEST._merge_dict({"VAL_LOSS": True, "EVAL_STU": False, "DRAW": False, "DICE": False, "DICE_THRES": 0.9, "TTT": True,
                          "BATCH": 1, "MIN_BATCH_NUM": None, "EVALUATOR": "COCOeval",
                          # not a reference key: pass 2 of a dataset inside the solver windows of the next dataset's pass 1
                          # (adapteacher/engine/trainer.py OverlappedEval); same results, different schedule
                          "OVERLAP_EVAL": True})
    cfg.MODEL.RPN._merge_dict({"UNSUP_LOSS_WEIGHT": 1.0, "LOSS": "CrossEntropy"})
    cfg.MODEL.ROI_HEADS._merge_dict({"LOSS": "CrossEntropy"})
    cfg.SOLVER._merge_dict({"IMG_PER_BATCH_LABEL": 1, "IMG_PER_BATCH_UNLABEL": 1, "FACTOR_LIST": (1,)})
    cfg.DATASETS._merge_dict({"TRAIN_LABEL": ("coco_2017_train",), "TRAIN_UNLABEL": ("coco_2017_train",), "CROSS_DATASET": True,
                              "NUM_BOUNDARY": 10, "NUM_CENTROID": 10, "RADIUS_CENTROID": 10})
    cfg.SEMISUPNET = CfgNode({"MLP_DIM": 128, "Trainer": "ateacher", "BBOX_THRESHOLD": 0.7, "PSEUDO_BBOX_SAMPLE": "thresholding",
                              "TEACHER_UPDATE_ITER": 1, "BURN_UP_STEP": 12000, "EMA_KEEP_RATE": 0.0, "UNSUP_LOSS_WEIGHT": 4.0,
                              "SUP_LOSS_WEIGHT": 0.5, "LOSS_WEIGHT_TYPE": "standard", "DIS_TYPE": "res4", "DIS_LOSS_WEIGHT": 0.1,
                              "CONTRASTIVE": False, "CONTRASTIVE_LOSS_WEIGHT": 0.05})
    cfg.DATALOADER._merge_dict({"SUP_PERCENT": 100.0, "RANDOM_DATA_SEED": 0, "RANDOM_DATA_SEED_PATH": None})
    cfg.EMAMODEL = CfgNode({"SUP_CONSIST": True})
    return cfg
