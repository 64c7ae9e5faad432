"""Checkpoint I/O; mirrors reference adapteacher/checkpoint/detection_checkpoint.py:10-93 (``DetectionTSCheckpointer``)
and the Detectron2 ``DetectionCheckpointer`` it extends, as used by train_net.py:38-63: ``resume_or_load(path)`` on a
``.pth`` file whose ``"model"`` entry is a state dict in Detectron2's names.

Handled like the reference: a ``module.`` prefix from (Distributed)DataParallel is stripped (:74-77); keys whose shape
differs from the model's are dropped and reported (:80-87); ``pixel_mean`` / ``pixel_std`` are never reported missing
(:27-36).  A teacher/student ensemble checkpoint (keys ``modelTeacher.*`` / ``modelStudent.*``, written by the
mean-teacher trainers) loaded into a single detector takes the TEACHER weights - the model train_net.py:45-57 tests.
Caffe2 / ``.pkl`` model-zoo files need Detectron2's name-matching heuristics and are refused.  New relative to the
reference (which never saves after test-time adaptation): ``save`` writes the adapted weights back in the same format."""
import os
from collections import namedtuple

import torch

_IncompatibleKeys = namedtuple("_IncompatibleKeys", ["missing_keys", "unexpected_keys", "incorrect_shapes"])


def _strip_prefix_if_present(sd, prefix):
    if sd and all(k.startswith(prefix) for k in sd):
        for k in list(sd.keys()):
            sd[k[len(prefix):]] = sd.pop(k)


class DetectionCheckpointer:
    def __init__(self, model, save_dir="", *, save_to_disk=True, **checkpointables):
        self.model = model
        self.save_dir = save_dir
        self.save_to_disk = save_to_disk
        self.checkpointables = checkpointables

    # ---- files
    def _last_file(self):
        return os.path.join(self.save_dir, "last_checkpoint")

    def has_checkpoint(self):
        return bool(self.save_dir) and os.path.exists(self._last_file())

    def get_checkpoint_file(self):
        with open(self._last_file()) as f:
            return os.path.join(self.save_dir, f.read().strip())

    def _load_file(self, path):
        if path.endswith(".pkl"):
            raise NotImplementedError("Caffe2 / model-zoo .pkl checkpoints need Detectron2's name-matching heuristics; "
                                      "convert them to a .pth state dict first")
        loaded = torch.load(path, map_location="cpu", weights_only=False)
        if "model" not in loaded:
            loaded = {"model": loaded}
        return loaded

    # ---- loading
    def _select(self, sd):
        """Teacher / student ensemble -> the sub-model this checkpointer's model is."""
        if any(k.startswith("modelTeacher.") for k in sd):
            return {k[len("modelTeacher."):]: v for k, v in sd.items() if k.startswith("modelTeacher.")}
        return sd

    def _load_model(self, checkpoint):
        sd = dict(checkpoint.pop("model"))
        sd = {k: (torch.from_numpy(v) if not torch.is_tensor(v) else v) for k, v in sd.items()}
        _strip_prefix_if_present(sd, "module.")
        sd = self._select(sd)
        model_sd = self.model.state_dict()
        incorrect = []
        for k in list(sd.keys()):
            if k in model_sd and tuple(model_sd[k].shape) != tuple(sd[k].shape):
                incorrect.append((k, tuple(sd[k].shape), tuple(model_sd[k].shape)))
                sd.pop(k)
        inc = self.model.load_state_dict(sd, strict=False)
        missing = [k for k in inc.missing_keys if k not in ("pixel_mean", "pixel_std")]
        return _IncompatibleKeys(missing, list(inc.unexpected_keys), incorrect)

    def load(self, path, checkpointables=None):
        if not path:
            return {}
        if not os.path.isfile(path):
            raise FileNotFoundError(f"Checkpoint {path} not found!")
        checkpoint = self._load_file(path)
        self.last_incompatible = self._load_model(checkpoint)
        for key in self.checkpointables if checkpointables is None else checkpointables:
            if key in checkpoint:
                self.checkpointables[key].load_state_dict(checkpoint.pop(key))
        return checkpoint

    def resume_or_load(self, path, *, resume=True):
        if resume and self.has_checkpoint():
            return self.load(self.get_checkpoint_file())
        return self.load(path, checkpointables=[])

    # ---- saving (Detectron2 format: {"model": state_dict, <checkpointables>...} + last_checkpoint)
    def save(self, name, **kwargs):
        if not self.save_dir or not self.save_to_disk:
            return None
        os.makedirs(self.save_dir, exist_ok=True)
        data = {"model": {k: v.detach().cpu() for k, v in self.model.state_dict().items()}}
        for key, obj in self.checkpointables.items():
            data[key] = obj.state_dict()
        data.update(kwargs)
        basename = f"{name}.pth"
        torch.save(data, os.path.join(self.save_dir, basename))
        with open(self._last_file(), "w") as f:
            f.write(basename)
        return os.path.join(self.save_dir, basename)


class DetectionTSCheckpointer(DetectionCheckpointer):
    """For a teacher/student ensemble model (attributes ``modelTeacher`` / ``modelStudent``): whole-model checkpoints load
    both halves; a plain (pretrained) state dict updates the student only (detection_checkpoint.py:11-36, 66-93)."""

    def _load_model(self, checkpoint):
        sd = checkpoint["model"]
        if hasattr(self.model, "modelStudent") and not any(k.startswith(("modelTeacher.", "modelStudent.")) for k in sd):
            return DetectionCheckpointer(self.model.modelStudent)._load_model(checkpoint)
        if hasattr(self.model, "modelStudent"):
            model, self.model = self.model, self.model           # whole ensemble: keys carry their own prefixes
            sd = dict(checkpoint.pop("model"))
            _strip_prefix_if_present(sd, "module.")
            inc = model.load_state_dict(sd, strict=False)
            return _IncompatibleKeys(list(inc.missing_keys), list(inc.unexpected_keys), [])
        return super()._load_model(checkpoint)
