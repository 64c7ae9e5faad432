"""CPU restatement of the image resize of the reference's test data path: Detectron2 ``ResizeShortestEdge`` ->
``ResizeTransform.apply_image`` on uint8 images = ``PIL.Image.resize((w, h), Image.BILINEAR)`` (reference
adapteacher/data/build.py:122-154 -> d2 DatasetMapper(cfg, False)).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pillow is a third-party dependency (requirements.txt); the algorithm restated here
is Pillow's ``ImagingResample`` for 8-bit images (src/libImaging/Resample.c): a separable triangle filter whose support grows with
the down-scaling factor (antialiasing), coefficients computed in double precision, normalised, converted to 22-bit fixed point
(round half away from zero), a horizontal pass then a vertical pass, each accumulating in int32 from 2^21 and clipping
``>> 22`` to [0, 255].  **Pinned**: tests/test_resize.py checks this restatement bit for bit against the installed Pillow itself
(up- and down-scaling, odd sizes, one-axis resizes)."""
import numpy as np

PRECISION_BITS = 32 - 8 - 2


def coefficients(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter (support 1.0) over the full axis.
    Returns (bounds int32 [out, 2] = (first input index, count), kk int32 [out, ksize], ksize)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.float64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        ww = 0.0
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            w = 1.0 - a if a < 1.0 else 0.0
            kk[xx, x] = w
            ww += w
        if ww != 0.0:
            kk[xx, :xmax] /= ww
        bounds[xx] = (xmin, xmax)
    fixed = np.where(kk < 0, (-0.5 + kk * (1 << PRECISION_BITS)).astype(np.int64), (0.5 + kk * (1 << PRECISION_BITS)).astype(np.int64))
    return bounds, fixed.astype(np.int32), ksize


def _pass(img, out_size, axis):
    bounds, kk, _ = coefficients(img.shape[axis], out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for o in range(out_size):
        lo, cnt = bounds[o]
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(cnt):
            acc += src[lo + x] * int(kk[o, x])
        out[o] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bilinear_u8(img_hwc, new_h, new_w):
    """PIL.Image.fromarray(img).resize((new_w, new_h), Image.BILINEAR) for a uint8 H x W x C array."""
    out = img_hwc
    if new_w != out.shape[1]:
        out = _pass(out, new_w, 1)                      # horizontal pass first (ImagingResample)
    if new_h != out.shape[0]:
        out = _pass(out, new_h, 0)
    return np.ascontiguousarray(out)
