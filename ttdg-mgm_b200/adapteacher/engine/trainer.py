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


class OverlappedEval:
    """Pass 2 (evaluation) of dataset k inside the solver windows of pass 1 (adaptation) of dataset k + 1.

    The reference runs, per dataset, an adaptation pass over every batch and then ``inference_on_dataset`` with the adapted
    weights (trainer.py:469-485), and the weights carry over to the next dataset.  In an adaptation step the GA-GM solver keeps
    one thread-block cluster (<= 8 of 148 SMs) busy for ~10 ms and nothing else of that step can run beside it - the loss needs
    its result.  The evaluation pass of the PREVIOUS dataset is independent work: it only needs the weights that dataset ended
    with.  So ``begin`` snapshots the weights into an evaluation replica of the model and ``window`` - called by
    ``ttdg_b200.ops.SOLVER_WINDOW_HOOK`` right after every solver launch - evaluates one queued batch with the replica on a
    second stream, its persistent kernels capped at the SMs the solver leaves free (``ttdg_set_sm_limit``).  ``drain`` runs
    what is left at full width and returns ``evaluator.evaluate()``.  Same kernels, same inputs, same weights as the sequential
    order: the results are identical (tests/test_gpu_overlap.py); only the schedule differs."""

    def __init__(self, model):
        self.model = model
        self.replica = None
        self.stream = None
        self.batches = None
        self.evaluator = None
        self.name = None
        self.images = 0
        self.windows = 0            # batches evaluated inside a solver window
        self._params = None
        self.t0 = 0.0

    @property
    def active(self):
        return self.batches is not None

    def _make_replica(self):
        import copy
        held = []                                          # per-step records (non-leaf tensors) do not belong in the replica
        for mod in self.model.modules():
            for attr in ("last_ttt", "last_aux"):
                if getattr(mod, attr, None) is not None:
                    held.append((mod, attr, getattr(mod, attr)))
                    setattr(mod, attr, None)
        try:
            rep = copy.deepcopy(self.model)
        finally:
            for mod, attr, val in held:
                setattr(mod, attr, val)
        rep.eval()
        for p in rep.parameters():
            p.requires_grad_(False)
        return rep

    def begin(self, name, loader, evaluator):
        """Snapshot the adapted weights and queue ``loader``'s batches for evaluation."""
        assert not self.active, "drain() the previous dataset first"
        if self.replica is None:
            self.replica = self._make_replica()
            if torch.cuda.is_available():
                self.stream = torch.cuda.Stream()
        else:
            with torch.no_grad():                           # a few multi-tensor launches instead of one small copy per parameter
                if self._params is None:                    # (the two module trees are walked once)
                    self._params = (list(self.replica.parameters()), list(self.model.parameters()))
                dst, src = self._params
                if hasattr(torch, "_foreach_copy_"):
                    torch._foreach_copy_(dst, src)
                else:
                    for pr, p in zip(dst, src):
                        pr.copy_(p)
        if hasattr(self.replica, "refresh_weight_copies"):
            self.replica.refresh_weight_copies()
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream())     # the snapshot is complete before the replica reads it
        self.name, self.evaluator, self.batches = name, evaluator, iter(loader)
        self.images, self.t0 = 0, time.perf_counter()
        evaluator.reset()

    def _eval_one(self):
        inputs = next(self.batches, None)
        if inputs is None:
            return False
        with torch.no_grad():
            if self.stream is not None:
                with torch.cuda.stream(self.stream):
                    self.evaluator.process(inputs, self.replica(inputs))
            else:
                self.evaluator.process(inputs, self.replica(inputs))
        self.images += len(inputs)
        return True

    def window(self):
        """One queued batch beside the solver that was just launched on the current stream."""
        if not self.active:
            return
        from ttdg_b200 import _C, ops
        lib = _C.lib() if torch.cuda.is_available() else None
        prev = None
        if lib is not None:
            sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
            prev = lib.ttdg_set_sm_limit(max(sms - ops.GAGM_CLUSTER_SMS, 1))
        try:
            if self._eval_one():
                self.windows += 1
        finally:
            if lib is not None:
                lib.ttdg_set_sm_limit(prev)

    def drain(self):
        """Evaluate what the windows did not get to; returns (results, evaluator) like ``inference_on_dataset``."""
        assert self.active
        while self._eval_one():
            pass
        if self.stream is not None:
            self.stream.synchronize()
        results = self.evaluator.evaluate()
        evaluator = self.evaluator
        evaluator.seconds_per_image = (time.perf_counter() - self.t0) / max(self.images, 1)
        self.batches = self.evaluator = self.name = None
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
        # Pass 2 of a dataset runs inside the solver windows of the NEXT dataset's pass 1 (OverlappedEval; TEST.OVERLAP_EVAL,
        # default on for a single process with TTT on a GPU model).  Results are identical to the sequential order.
        from ttdg_b200 import ops
        overlap = (bool(getattr(cfg.TEST, "OVERLAP_EVAL", True)) and cfg.TEST.TTT and world_size == 1 and len(cfg.DATASETS.TEST) > 1
                   and torch.cuda.is_available() and next(model.parameters()).is_cuda)
        pipe = OverlappedEval(model) if overlap else None
        for idx, name in enumerate(cfg.DATASETS.TEST):
            loader = data_loaders[name]
            if pipe is not None and pipe.active:
                ops.SOLVER_WINDOW_HOOK[0] = pipe.window
            try:
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
            finally:
                ops.SOLVER_WINDOW_HOOK[0] = None
            evaluator = evaluators[idx] if evaluators is not None else DiceEvaluator(name, cfg.TEST.DICE_THRES, (dataset_dicts or {}).get(name))
            if pipe is not None:
                if pipe.active:                            # the previous dataset's pass 2: whatever the windows left over
                    prev_name = pipe.name
                    results[prev_name], _ = pipe.drain()
                pipe.begin(name, loader, evaluator)        # this dataset's pass 2 starts with the next pass 1 (or below)
                if idx + 1 == len(cfg.DATASETS.TEST):
                    results[name], _ = pipe.drain()
                continue
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
