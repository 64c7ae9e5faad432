"""BaselineTrainer.test and inference_on_dataset; mirrors reference adapteacher/engine/trainer.py:430-529 (the
test-time-adaptation driver: adapt on every batch of a dataset, then evaluate the whole dataset with the adapted
weights, weights carrying over to the next dataset) and :1230-1360 (eval loop).  Detectron2's data loader / catalog /
logging are outside the hot path: ``data_loaders`` is a mapping dataset name -> iterable of batches (lists of dicts
with 'image', 'height', 'width', 'image_id') and ``dataset_dicts`` the matching ground truth."""
import time
from collections import OrderedDict, defaultdict

import torch

from adapteacher.evaluation.dice_metric import DiceEvaluator


def inference_on_dataset(model, data_loader, evaluator, cfg=None):
    """trainer.py:1230-1360: eval-mode pass over the loader, feeding the evaluator; returns (results, evaluator)."""
    was_training = model.training
    model.eval()                                           # inference_context (:1362-1374)
    evaluator.reset()
    t0, n = time.perf_counter(), 0
    with torch.no_grad():
        for inputs in data_loader:
            outputs = model(inputs)
            if torch.cuda.is_available():
                torch.cuda.synchronize()                   # :1310-1311
            evaluator.process(inputs, outputs)
            n += len(inputs)
    model.train(was_training)
    results = evaluator.evaluate()
    evaluator.seconds_per_image = (time.perf_counter() - t0) / max(n, 1)
    return (results if results is not None else {}), evaluator


class BaselineTrainer:
    @classmethod
    def _ttt_pass_sharded(cls, cfg, model, optimizer, loader, world_size):
        """The adaptation pass when images are sharded over ranks.  Shards can differ by one batch and a rank's
        batch can yield no loss (single graph, mgm:489-490), but the gradient all-reduce is collective: all ranks run
        max-over-ranks steps; a rank without a loss contributes a zero gradient; a step no rank has a loss for is skipped
        on all of them (decided by a one-element all-reduce)."""
        import torch.distributed as dist
        model.train()
        dev = next(model.parameters()).device
        n_local = len(loader) if cfg.TEST.MIN_BATCH_NUM is None else min(len(loader), cfg.TEST.MIN_BATCH_NUM)
        t = torch.tensor([n_local], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        it = iter(loader)
        for b in range(int(t.item())):
            inputs = next(it, None) if b < n_local else None
            loss = model(inputs, branch="TTT")[0] if inputs is not None else None
            have = torch.tensor([0 if loss is None else 1], dtype=torch.int64, device=dev)
            dist.all_reduce(have)
            if int(have.item()) == 0:
                continue
            optimizer.zero_grad()
            if loss is not None:
                loss.backward()
            optimizer.step(world_size)

    @classmethod
    def test(cls, cfg, model, optimizer=None, evaluators=None, data_loaders=None, dataset_dicts=None, world_size=1):
        """cfg needs DATASETS.TEST, TEST.TTT, TEST.MIN_BATCH_NUM, TEST.DICE_THRES (adapteacher/config.py:15-17)."""
        results = OrderedDict()
        for idx, name in enumerate(cfg.DATASETS.TEST):
            loader = data_loaders[name]
            if cfg.TEST.TTT and world_size > 1:            # pass 1 on image shards: every rank takes part in every all-reduce
                cls._ttt_pass_sharded(cfg, model, optimizer, loader, world_size)
            elif cfg.TEST.TTT:                             # pass 1: adaptation, model in train mode (:469-482)
                model.train()
                for b, inputs in enumerate(loader):
                    if cfg.TEST.MIN_BATCH_NUM is not None and b >= cfg.TEST.MIN_BATCH_NUM:
                        break
                    loss, _, _, _ = model(inputs, branch="TTT")
                    if loss is None:
                        continue
                    optimizer.zero_grad()
                    loss.backward()
                    if hasattr(optimizer, "flat_g"):
                        optimizer.step(world_size)         # fused flat-bucket step (+ gradient all-reduce)
                    else:
                        optimizer.step()
            evaluator = evaluators[idx] if evaluators is not None else DiceEvaluator(name, cfg.TEST.DICE_THRES, (dataset_dicts or {}).get(name))
            results_i, _ = inference_on_dataset(model, loader, evaluator, cfg)       # pass 2 (:484-485)
            results[name] = results_i
        groups = defaultdict(lambda: defaultdict(list))    # per-family means over datasets sharing a prefix (:509-527)
        for key, value in results.items():
            for metric in ("Dice Coefficient", "Enhanced Alignment Metric", "Structural Similarity Metric"):
                if metric in value:
                    groups[key.split("_")[0]][metric].append(value[metric])
        for fam, metrics in groups.items():
            results[f"{fam}_mean"] = {k: sum(v) / len(v) for k, v in metrics.items()}
        return results
