#!/usr/bin/env python
"""Entry point; mirrors the ``--eval-only`` path of the reference's train_net.py:24-84:

    python ttdg-mgm_b200/train_net.py --eval-only --config-file <yaml> [--num-gpus N] MODEL.WEIGHTS <pth> [KEY VAL ...]

setup (yaml with ``_BASE_`` inheritance + ``add_ateacher_config`` defaults + command-line overrides) -> build the
meta-architecture named by ``MODEL.META_ARCHITECTURE`` -> ``DetectionCheckpointer(model).resume_or_load(MODEL.WEIGHTS)``
-> ``Trainer.test(cfg, model, optimizer)`` (test-time adaptation on every batch, then evaluation with the adapted
weights) -> results appended to ``OUTPUT_DIR/result_ap.txt``.  Training (no ``--eval-only``) is out of scope.
With ``--num-gpus N`` it re-launches itself under torch.distributed.run, one process per GPU; images shard over ranks
and the adaptation gradients are all-reduced (NCCL)."""
import argparse
import ast
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

import torch  # noqa: E402
import yaml  # noqa: E402

from adapteacher.checkpoint import DetectionCheckpointer  # noqa: E402
from adapteacher.config import add_ateacher_config, get_cfg  # noqa: E402
from adapteacher.data import build_detection_test_loader  # noqa: E402
from adapteacher.engine.trainer import BaselineTrainer  # noqa: E402

# hacky way to register (reference train_net.py:14-20): importing the modules fills the registries
from adapteacher.modeling.meta_arch.rcnn import DAobjTwoStagePseudoLabGeneralizedRCNN, TwoStagePseudoLabGeneralizedRCNN  # noqa: E402,F401
from adapteacher.modeling.proposal_generator.rpn import PseudoLabRPN  # noqa: E402,F401
from adapteacher.modeling.roi_heads.roi_heads import StandardROIHeadsPseudoLab  # noqa: E402,F401
from ttdg_b200.registry import META_ARCH_REGISTRY  # noqa: E402


def setup(args):
    """Create configs and perform basic setups - the reference's own sequence (reference train_net.py:23-33)."""
    cfg = get_cfg()
    add_ateacher_config(cfg)
    if args.config_file:
        cfg.merge_from_file(args.config_file)
    cfg.merge_from_list(args.opts)
    cfg.freeze()
    return cfg


class Trainer(BaselineTrainer):
    @classmethod
    def build_model(cls, cfg):
        arch = META_ARCH_REGISTRY.get(cfg.MODEL.META_ARCHITECTURE)      # KeyError = unknown name, like the d2 registry
        return arch.from_config(cfg).to("cuda")                         # sub-modules by MODEL.{BACKBONE,PROPOSAL_GENERATOR,ROI_HEADS}.NAME

    @classmethod
    def build_optimizer(cls, cfg, model):
        from ttdg_b200.optim import FlatSGD
        s = cfg.SOLVER
        opt = FlatSGD(model.adapted_parameters(), lr=s.BASE_LR, momentum=getattr(s, "MOMENTUM", 0.9),
                      weight_decay=getattr(s, "WEIGHT_DECAY", 1e-4), buckets=[len(g) for g in model.adapted_parameter_groups()],
                      on_step=model.refresh_weight_copies)
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world > 1:                                       # bucketed all-reduce overlapped with the backward pass
            from ttdg_b200 import detector
            opt.enable_overlap(world)
            detector.GRAD_READY_HOOK[0] = opt.grad_ready
        return opt

    @classmethod
    def build_test_loader(cls, cfg, dataset_name):
        return build_detection_test_loader(cfg, dataset_name)


def main(args):
    cfg = setup(args)
    if cfg.SEMISUPNET.Trainer == "ateacher":                   # reference train_net.py:38-55: mean-teacher ensemble, no test-time adaptation
        raise NotImplementedError("SEMISUPNET.Trainer 'ateacher' (teacher / student ensemble evaluation) is outside the test-time-"
                                  "adaptation path; use SEMISUPNET.Trainer baseline (configs/test_segment.yaml)")
    if cfg.SEMISUPNET.Trainer != "baseline":
        raise ValueError("Trainer Name is not found.")         # reference train_net.py:43
    if not args.eval_only:
        raise NotImplementedError("only --eval-only (test-time adaptation + evaluation) is implemented")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        torch.distributed.init_process_group("nccl")
    model = Trainer.build_model(cfg)
    DetectionCheckpointer(model, save_dir=cfg.OUTPUT_DIR).resume_or_load(cfg.MODEL.WEIGHTS, resume=args.resume, strict=True)
    loaders = {name: Trainer.build_test_loader(cfg, name) for name in cfg.DATASETS.TEST}
    res = Trainer.test(cfg, model, Trainer.build_optimizer(cfg, model), data_loaders=loaders, world_size=world)
    if world == 1 or torch.distributed.get_rank() == 0:
        print(res)
        os.makedirs(cfg.OUTPUT_DIR, exist_ok=True)
        with open(os.path.join(cfg.OUTPUT_DIR, "result_ap.txt"), "a") as f:
            f.write("loading data from: " + str(cfg.MODEL.WEIGHTS) + "\n")
            f.write(json.dumps({k: {m: float(v) for m, v in d.items()} for k, d in res.items()}) + "\n")
    if world > 1:
        torch.distributed.destroy_process_group()
    return res


def _free_port():
    import socket
    with socket.socket() as sck:
        sck.bind(("127.0.0.1", 0))
        return sck.getsockname()[1]


def default_argument_parser():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config-file", default="", metavar="FILE")
    ap.add_argument("--config", dest="config_file", help=argparse.SUPPRESS)       # the README's spelling (README.md:93-94)
    ap.add_argument("--resume", action="store_true", default=True,
                    help="a last_checkpoint in OUTPUT_DIR wins over MODEL.WEIGHTS (the reference forces this, train_net.py:92)")
    ap.add_argument("--no-resume", dest="resume", action="store_false")
    ap.add_argument("--eval-only", action="store_true")
    ap.add_argument("--num-gpus", type=int, default=1)
    ap.add_argument("opts", nargs=argparse.REMAINDER, default=[])
    return ap


if __name__ == "__main__":
    args = default_argument_parser().parse_args()
    if args.num_gpus > 1 and "WORLD_SIZE" not in os.environ:       # d2 launch(): one process per GPU
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.num_gpus}",
                                   "--master-addr", "127.0.0.1", "--master-port", str(_free_port())] + sys.argv)
    main(args)
