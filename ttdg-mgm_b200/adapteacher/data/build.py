"""Test data path; mirrors reference adapteacher/data/build.py:122-154 (``build_detection_test_loader``: dataset dicts ->
mapper -> InferenceSampler -> batches of TEST.BATCH, ``drop_last=False``, trivial collate) with Detectron2's pieces
restated: ``DatasetMapper(cfg, False)`` (read image, ``INPUT.FORMAT``, ResizeShortestEdge(MIN_SIZE_TEST, MAX_SIZE_TEST),
CHW uint8 tensor), ``InferenceSampler`` (contiguous shard per rank) and the polygon rasteriser of pycocotools
(``frPyObjects`` + ``decode``, used by dice_metric.py:94-108 for the ground-truth masks).  Detectron2 / pycocotools are
absent offline: these restatements are PARITY UNPINNED."""
import os

import numpy as np
import torch


class _Catalog(dict):
    def register(self, name, fn):
        self[name] = fn

    def get(self, name):                                       # noqa: A003 - Detectron2's spelling
        if name not in self:
            if name.startswith("synthetic_"):
                from .datasets.builtin import register_synthetic
                register_synthetic(name)
            else:
                raise KeyError(f"Dataset '{name}' is not registered")
        v = self[name]
        if callable(v):
            v = self[name] = v()
        return v


DatasetCatalog = _Catalog()


# ------------------------------------------------------------------------------------------------ polygons -> masks
def polygon_to_mask(xy, h, w):
    """pycocotools ``rleFrPoly`` (maskApi.c) restated: boundary upsampled x5, points along the y-boundaries, column-major
    run lengths.  xy: flat [x0, y0, x1, y1, ...]."""
    xy = np.asarray(xy, dtype=np.float64).reshape(-1, 2)
    k = len(xy)
    scale = 5.0
    x = np.floor(scale * xy[:, 0] + 0.5).astype(np.int64)
    y = np.floor(scale * xy[:, 1] + 0.5).astype(np.int64)
    x, y = np.append(x, x[0]), np.append(y, y[0])
    us, vs = [], []
    for j in range(k):
        xs, xe, ys, ye = int(x[j]), int(x[j + 1]), int(y[j]), int(y[j + 1])
        dx, dy = abs(xe - xs), abs(ys - ye)
        flip = (dx >= dy and xs > xe) or (dx < dy and ys > ye)
        if flip:
            xs, xe, ys, ye = xe, xs, ye, ys
        if dx >= dy:
            s = (ye - ys) / dx if dx else 0.0
            d = np.arange(dx + 1)
            t = dx - d if flip else d
            us.append(t + xs)
            vs.append(np.floor(ys + s * t + 0.5).astype(np.int64))
        else:
            s = (xe - xs) / dy
            d = np.arange(dy + 1)
            t = dy - d if flip else d
            vs.append(t + ys)
            us.append(np.floor(xs + s * t + 0.5).astype(np.int64))
    u, v = np.concatenate(us), np.concatenate(vs)
    chg = np.nonzero(u[1:] != u[:-1])[0] + 1
    xd = np.where(u[chg] < u[chg - 1], u[chg], u[chg] - 1).astype(np.float64)
    xd = (xd + 0.5) / scale - 0.5
    ok = (np.floor(xd) == xd) & (xd >= 0) & (xd <= w - 1)
    yd = np.minimum(v[chg], v[chg - 1]).astype(np.float64)
    yd = np.ceil(np.clip((yd + 0.5) / scale - 0.5, 0, h))
    a = np.sort((xd[ok] * h + yd[ok]).astype(np.int64))
    # each boundary crossing toggles the (column-major) fill state from that position on
    flat = np.zeros(h * w + 1, dtype=np.int64)
    np.add.at(flat, a, 1)
    mask = (np.cumsum(flat)[:h * w] % 2).astype(bool)
    return mask.reshape(w, h).T


def segmentation_to_mask(seg, h, w):
    """COCO ``segmentation`` (list of polygons, or an already binary array) -> bool H x W (polygons are OR-ed)."""
    if isinstance(seg, np.ndarray):
        return seg.astype(bool)
    if isinstance(seg, dict):
        raise NotImplementedError("RLE segmentations need pycocotools' decoder")
    m = np.zeros((h, w), bool)
    for poly in seg:
        if len(poly) >= 6:
            m |= polygon_to_mask(poly, h, w)
    return m


# ------------------------------------------------------------------------------------------------ mapper / sampler / loader
class DatasetMapper:
    """d2 DatasetMapper(cfg, is_train=False): image in ``INPUT.FORMAT``, shortest edge resized to MIN_SIZE_TEST (max
    MAX_SIZE_TEST), uint8 CHW tensor; 'height' / 'width' keep the ORIGINAL size (the masks are pasted back to it)."""

    def __init__(self, cfg, is_train=False, device_resize=None):
        inp = getattr(cfg, "INPUT", None)
        self.format = getattr(inp, "FORMAT", "BGR")
        self.min_size = getattr(inp, "MIN_SIZE_TEST", 800)
        self.max_size = getattr(inp, "MAX_SIZE_TEST", 1333)
        # device_resize: the decoded image goes to the GPU at its ORIGINAL size and is resized there (csrc/resize.cu, bit-exact with
        # PIL's bilinear resampler) - the host does the JPEG / PNG decode only.  Default: on when CUDA is available
        # (TTDG_DEVICE_RESIZE=0 keeps PIL's resize on the host).
        if device_resize is None:
            device_resize = torch.cuda.is_available() and os.environ.get("TTDG_DEVICE_RESIZE", "1") != "0"
        self.device_resize = bool(device_resize)

    def _resize_shape(self, h, w):
        if not self.min_size:
            return h, w
        scale = self.min_size / min(h, w)
        nh, nw = (self.min_size, scale * w) if h < w else (scale * h, self.min_size)
        if max(nh, nw) > self.max_size:
            s = self.max_size / max(nh, nw)
            nh, nw = nh * s, nw * s
        return int(nh + 0.5), int(nw + 0.5)

    def __call__(self, d):
        d = dict(d)
        if "image" in d:                                       # synthetic datasets carry their tensor
            img = d["image"]
        else:
            from PIL import Image
            with Image.open(d["file_name"]) as im:
                im = im.convert("RGB")
                nh, nw = self._resize_shape(im.height, im.width)
                if self.device_resize:
                    from ttdg_b200 import ops
                    raw = torch.from_numpy(np.ascontiguousarray(np.asarray(im)))                  # H x W x 3, original size
                    raw = (raw.pin_memory() if torch.cuda.is_available() else raw).to("cuda", non_blocking=True)
                    img = ops.resize_bilinear_u8(raw, nh, nw, planar=True, flip=self.format == "BGR")     # 3 x nh x nw on the device
                    d["image"] = img
                    d.pop("annotations", None)
                    return d
                if (nh, nw) != (im.height, im.width):
                    im = im.resize((nw, nh), Image.BILINEAR)   # d2 ResizeTransform: PIL bilinear on uint8 images
                arr = np.asarray(im)
            if self.format == "BGR":
                arr = arr[:, :, ::-1]
            img = torch.from_numpy(np.ascontiguousarray(arr.transpose(2, 0, 1)))
        d["image"] = img
        d.pop("annotations", None)                             # is_train = False
        return d


class InferenceSampler:
    """d2 InferenceSampler: rank r gets the contiguous slice [r * ceil(n / world), ...)."""

    def __init__(self, size, rank=None, world_size=None):
        if rank is None:
            import torch.distributed as dist
            on = dist.is_available() and dist.is_initialized()
            rank, world_size = (dist.get_rank(), dist.get_world_size()) if on else (0, 1)
        shard = (size + world_size - 1) // world_size
        self.indices = range(min(size, shard * rank), min(size, shard * (rank + 1)))

    def __iter__(self):
        return iter(self.indices)

    def __len__(self):
        return len(self.indices)


class _TestLoader:
    """Re-iterable (BaselineTrainer.test walks it twice: adaptation pass, then evaluation pass)."""

    def __init__(self, dicts, mapper, sampler, batch, pin):
        self.dicts, self.mapper, self.sampler, self.batch, self.pin = dicts, mapper, sampler, batch, pin

    def __len__(self):
        return (len(self.sampler) + self.batch - 1) // self.batch

    def __iter__(self):
        cur = []
        for i in self.sampler:
            d = self.mapper(self.dicts[i])
            if self.pin and torch.cuda.is_available() and not d["image"].is_cuda:
                d["image"] = d["image"].pin_memory()
            cur.append(d)
            if len(cur) == self.batch:
                yield cur
                cur = []
        if cur:                                                # drop_last = False (build.py:146)
            yield cur


def build_detection_test_loader(cfg, dataset_name, mapper=None, rank=None, world_size=None):
    dicts = DatasetCatalog.get(dataset_name)
    mapper = mapper if mapper is not None else DatasetMapper(cfg, False)
    batch = cfg.TEST.BATCH if cfg.TEST.TTT else 1             # build.py:141-144
    return _TestLoader(dicts, mapper, InferenceSampler(len(dicts), rank, world_size), batch, pin=True)
