"""The drop-in boundary without a GPU: registries resolve the names the reference's yaml files use, the adapters carry the
reference's call signatures (read from /root/reference when it is present, pinned copies otherwise), the entry point
dispatches through them."""
import ast
import inspect
import os
import sys
from types import SimpleNamespace

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

# (file, class, method) -> parameter names of the reference (adapteacher/...: rpn.py:16-23, roi_heads.py:65-74, rcnn.py:154-156,
# multi_graph_matching.py:487, build_graph.py:160)
PINNED = {
    ("adapteacher/modeling/proposal_generator/rpn.py", "PseudoLabRPN", "forward"):
        ["self", "images", "features", "gt_instances", "compute_loss", "compute_val_loss"],
    ("adapteacher/modeling/roi_heads/roi_heads.py", "StandardROIHeadsPseudoLab", "forward"):
        ["self", "images", "features", "proposals", "targets", "compute_loss", "branch", "compute_val_loss"],
    ("adapteacher/modeling/meta_arch/rcnn.py", "DAobjTwoStagePseudoLabGeneralizedRCNN", "forward"):
        ["self", "batched_inputs", "branch", "given_proposals", "val_mode"],
    ("adapteacher/modeling/GModule/multi_graph_matching.py", "MGM3_unsup", "forward"): ["self", "nodes", "labels", "U"],
    ("adapteacher/modeling/GModule/build_graph.py", "PrototypeComputation", "__call__"): ["self", "features", "targets"],
}


def _ref_args(path, cls, fn):
    tree = ast.parse(open(os.path.join(REF, path)).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for f in node.body:
                if isinstance(f, ast.FunctionDef) and f.name == fn:
                    return [a.arg for a in f.args.args]
    raise KeyError((path, cls, fn))


def _ours(path, cls):
    mod = __import__(path[:-3].replace("/", "."), fromlist=[cls])
    return getattr(mod, cls)


@pytest.mark.parametrize("key", sorted(PINNED))
def test_adapter_signatures_are_the_reference_ones(key):
    path, cls, fn = key
    want = PINNED[key]
    if os.path.isdir(REF):
        assert _ref_args(path, cls, fn) == want          # the pinned copy is what the reference really has
    got = list(inspect.signature(getattr(_ours(path, cls), fn)).parameters)
    assert got[:len(want)] == want, (got, want)


def test_registries_resolve_the_yaml_names():
    import train_net  # noqa: F401  (importing registers, reference train_net.py:14-20)
    from ttdg_b200.registry import BACKBONE_REGISTRY, META_ARCH_REGISTRY, PROPOSAL_GENERATOR_REGISTRY, ROI_HEADS_REGISTRY
    from ttdg_b200 import detector
    assert META_ARCH_REGISTRY.get("DAobjTwoStagePseudoLabGeneralizedRCNN").__name__ == "DAobjTwoStagePseudoLabGeneralizedRCNN"
    assert "TwoStagePseudoLabGeneralizedRCNN" in META_ARCH_REGISTRY
    assert issubclass(PROPOSAL_GENERATOR_REGISTRY.get("PseudoLabRPN"), detector.RPN)
    assert issubclass(ROI_HEADS_REGISTRY.get("StandardROIHeadsPseudoLab"), detector.ROIHeads)
    assert BACKBONE_REGISTRY.get("build_resnet_fpn_backbone") is detector.Backbone
    with pytest.raises(KeyError):
        ROI_HEADS_REGISTRY.get("Res5ROIHeads")


def test_from_config_dispatches_by_name():
    """MODEL.{BACKBONE, PROPOSAL_GENERATOR, ROI_HEADS}.NAME of the reference's own yaml pick the sub-module classes."""
    import train_net
    cfg_file = os.path.join(REF, "configs", "test_segment.yaml")
    if not os.path.isfile(cfg_file):
        cfg_file = os.path.join(ROOT, "ttdg-mgm_b200", "configs", "test_segment_synthetic.yaml")
    args = SimpleNamespace(config_file=cfg_file, opts=[])
    cfg = train_net.setup(args)
    assert cfg.MODEL.PROPOSAL_GENERATOR.NAME == "PseudoLabRPN" and cfg.MODEL.ROI_HEADS.NAME == "StandardROIHeadsPseudoLab"
    from ttdg_b200.registry import META_ARCH_REGISTRY
    arch = META_ARCH_REGISTRY.get(cfg.MODEL.META_ARCHITECTURE)
    m = arch.from_config(cfg)                              # CPU construction: parameters only, no kernel runs
    assert type(m.proposal_generator).__name__ == "PseudoLabRPN" and type(m.roi_heads).__name__ == "StandardROIHeadsPseudoLab"
    assert m.roi_heads.num_classes == cfg.MODEL.ROI_HEADS.NUM_CLASSES
    with pytest.raises(AttributeError):                    # setup() freezes the config, like the reference's (train_net.py:31)
        cfg.MODEL.ROI_HEADS.NAME = "NoSuchHeads"
    cfg = cfg.clone()
    cfg.defrost()
    cfg.MODEL.ROI_HEADS.NAME = "NoSuchHeads"
    with pytest.raises(KeyError):
        arch.from_config(cfg)
    # train() / eval() reach every sub-module (nn.Module contract), also after the cached fast path is in place
    m.train(); m.eval()
    assert not m.roi_heads.training and not m.backbone.bottom_up.res3[0].conv1.training
    m.train()
    assert m.roi_heads.training and m.proposal_generator.training and m.multi_matching_unsup.intra_domain_graph.training


def test_training_branches_refuse_loudly():
    from adapteacher.modeling.proposal_generator.rpn import PseudoLabRPN
    from adapteacher.modeling.roi_heads.roi_heads import StandardROIHeadsPseudoLab
    from ttdg_b200.structures import ImageList
    rpn, heads = PseudoLabRPN(), StandardROIHeadsPseudoLab(2)
    rpn.train(); heads.train()
    with pytest.raises(NotImplementedError):
        rpn(ImageList(None, [(32, 32)]), {}, None)                       # compute_loss defaults to True, as in the reference
    with pytest.raises(NotImplementedError):
        heads(None, {}, [], targets=None)


def test_image_list_keeps_per_image_sizes():
    from ttdg_b200.structures import ImageList
    il = ImageList(None, [(96, 128), (128, 160)])
    assert len(il) == 2 and il.image_sizes == [(96, 128), (128, 160)]


def test_cfgnode_mirrors_the_yacs_calls_of_the_reference_entry_point(tmp_path):
    """get_cfg / add_ateacher_config(cfg) / merge_from_file (with _BASE_) / merge_from_list / freeze / clone / defrost:
    the calls of reference train_net.py:23-33, on this package's CfgNode."""
    from adapteacher.config import CfgNode, add_ateacher_config, get_cfg
    base = tmp_path / "base.yaml"
    base.write_text("MODEL:\n  META_ARCHITECTURE: GeneralizedRCNN\n  ROI_HEADS:\n    NUM_CLASSES: 7\nSOLVER:\n  BASE_LR: 0.02\n")
    top = tmp_path / "top.yaml"
    top.write_text("_BASE_: base.yaml\nMODEL:\n  ROI_HEADS:\n    NAME: StandardROIHeadsPseudoLab\nTEST:\n  BATCH: 5\nDATASETS:\n  TEST: (\"a\", \"b\")\n")
    cfg = get_cfg()
    assert add_ateacher_config(cfg) is cfg
    assert cfg.TEST.TTT is True and cfg.TEST.BATCH == 1 and cfg.TEST.MIN_BATCH_NUM is None and cfg.SEMISUPNET.Trainer == "ateacher"
    cfg.merge_from_file(str(top))
    assert cfg.MODEL.ROI_HEADS.NUM_CLASSES == 7 and cfg.MODEL.ROI_HEADS.NAME == "StandardROIHeadsPseudoLab"     # _BASE_ then override
    assert cfg.SOLVER.BASE_LR == 0.02 and cfg.SOLVER.MOMENTUM == 0.9 and cfg.TEST.BATCH == 5 and cfg.DATASETS.TEST == ("a", "b")
    cfg.merge_from_list(["SOLVER.BASE_LR", "0.005", "SEMISUPNET.Trainer", "baseline", "OUTPUT_DIR", "out/x", "NEW.KEY", "(1, 2)"])
    assert cfg.SOLVER.BASE_LR == 0.005 and cfg.SEMISUPNET.Trainer == "baseline" and cfg.OUTPUT_DIR == "out/x" and cfg.NEW.KEY == (1, 2)
    with pytest.raises(ValueError):
        cfg.merge_from_list(["A"])
    cfg.freeze()
    assert cfg.is_frozen() and cfg.MODEL.is_frozen()
    with pytest.raises(AttributeError):
        cfg.TEST.BATCH = 3
    with pytest.raises(AttributeError):
        cfg["TEST"] = 1
    c2 = cfg.clone()
    c2.defrost()
    c2.TEST.BATCH = 3
    assert c2.TEST.BATCH == 3 and cfg.TEST.BATCH == 5
    assert isinstance(c2.MODEL, CfgNode) and hasattr(c2.MODEL, "WEIGHTS") and not hasattr(c2.MODEL, "NOPE")
    assert "BATCH: 5" in cfg.dump()
