"""Minimal stand-ins for the Detectron2 structures the hot path touches (d2 is not a dependency here):
``Boxes`` (``.tensor``), ``Instances`` (attribute bag with ``_fields``, ``len``) - exactly the members
adapteacher/modeling/GModule/build_graph.py:78-85 and adapteacher/evaluation/dice_metric.py:34-36 read."""


class ImageList:
    """d2 ImageList: ``tensor`` = the padded batch (here: the preprocessed NHWC fp32 tensor the backbone kernels read),
    ``image_sizes`` = [(h, w)] of every image before padding (boxes are clipped to these, not to the canvas)."""

    def __init__(self, tensor, image_sizes):
        self.tensor = tensor
        self.image_sizes = [tuple(int(v) for v in s) for s in image_sizes]

    def __len__(self):
        return len(self.image_sizes)


class Boxes:
    def __init__(self, tensor):
        self.tensor = tensor

    def __len__(self):
        return self.tensor.shape[0]


class Instances:
    def __init__(self, image_size, **fields):
        object.__setattr__(self, "_image_size", tuple(image_size))
        object.__setattr__(self, "_fields", {})
        for k, v in fields.items():
            self.set(k, v)

    @property
    def image_size(self):
        return self._image_size

    def set(self, name, value):
        self._fields[name] = value

    def has(self, name):
        return name in self._fields

    def get_fields(self):
        return self._fields

    def __setattr__(self, name, value):
        if name.startswith("_"):
            object.__setattr__(self, name, value)
        else:
            self.set(name, value)

    def __getattr__(self, name):
        if name == "_fields" or name not in self._fields:
            raise AttributeError("Cannot find field '{}' in the given Instances!".format(name))
        return self._fields[name]

    def __len__(self):
        for v in self._fields.values():
            return len(v)
        return 0

    def num_valid(self):
        """Rows that are real when the fields are padded to a fixed capacity (PseudoLabRPN proposals); = len() otherwise."""
        pad = self.__dict__.get("_padded")
        return int(pad[2][pad[3]].item()) if pad is not None else len(self)

    def to(self, device):
        out = Instances(self._image_size)
        for k, v in self._fields.items():
            out.set(k, Boxes(v.tensor.to(device)) if isinstance(v, Boxes) else v.to(device))
        return out
