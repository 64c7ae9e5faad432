"""Image resize of the test data path: d2 ResizeShortestEdge -> PIL.Image.resize(..., BILINEAR) on uint8 images.
CPU: the oracle restatement (oracle/pil_resize.py) is pinned bit for bit to the installed Pillow itself, and the library's host-side
coefficient builder to the oracle.  GPU: csrc/resize.cu through the C ABI is bit-exact with Pillow, planar / BGR output included,
and DatasetMapper's device path gives the tensor of the host path."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ttdg-mgm_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
from oracle import pil_resize  # noqa: E402

Image = pytest.importorskip("PIL.Image")

CASES = [(37, 53, 80, 115), (200, 300, 128, 192), (64, 64, 64, 100), (101, 77, 50, 77), (480, 640, 800, 1067), (600, 900, 333, 500),
         (31, 1000, 31, 17), (5, 7, 1, 1), (1, 1, 9, 4), (123, 45, 123, 45)]


def _img(H, W, seed, C=3):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, (H, W, C), dtype=np.uint8)
    a[: H // 3, : W // 2] = 255                                            # flat saturated / zero areas: the clip paths
    a[H // 2:, W // 2:] = 0
    return a


def _pil(a, nh, nw):
    return np.asarray(Image.fromarray(a).resize((nw, nh), Image.BILINEAR))


@pytest.mark.parametrize("H,W,nh,nw", CASES)
def test_oracle_is_pillow(H, W, nh, nw):
    a = _img(H, W, H * 1000 + W)
    assert np.array_equal(pil_resize.resize_bilinear_u8(a, nh, nw), _pil(a, nh, nw))


@pytest.mark.parametrize("n_in,n_out", [(53, 115), (300, 192), (640, 1067), (900, 500), (1000, 17), (7, 1), (1, 4), (1333, 800), (2048, 800)])
def test_library_coefficient_tables_are_the_oracles(n_in, n_out):
    """ttdg_resize_coeffs_u8 is a HOST function of the C-ABI library (no GPU needed): same bounds and fixed-point weights."""
    from ttdg_b200 import _C
    L = _C.lib()
    ks = L.ttdg_resize_ksize(n_in, n_out)
    bounds, kk, ks_ref = pil_resize.coefficients(n_in, n_out)
    assert ks == ks_ref
    b = np.zeros((n_out, 2), np.int32)
    k = np.zeros((n_out, ks), np.int32)
    assert L.ttdg_resize_coeffs_u8(n_in, n_out, b.ctypes.data_as(ctypes.c_void_p), k.ctypes.data_as(ctypes.c_void_p)) == ks
    assert np.array_equal(b, bounds) and np.array_equal(k, kk)
    assert L.ttdg_resize_ksize(0, 5) == -1


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,nh,nw", CASES)
def test_device_resize_is_pillow(H, W, nh, nw):
    from ttdg_b200 import ops
    a = _img(H, W, H * 7 + W)
    ref = _pil(a, nh, nw)
    x = torch.from_numpy(a).cuda()
    out = ops.resize_bilinear_u8(x, nh, nw, planar=False)
    assert out.shape == (nh, nw, 3) and np.array_equal(out.cpu().numpy(), ref)
    chw_bgr = ops.resize_bilinear_u8(x, nh, nw, planar=True, flip=True)
    assert np.array_equal(chw_bgr.cpu().numpy(), np.ascontiguousarray(ref[:, :, ::-1].transpose(2, 0, 1)))
    assert np.array_equal(out.cpu().numpy(), pil_resize.resize_bilinear_u8(a, nh, nw))


@pytest.mark.gpu
def test_device_resize_other_channel_counts_and_errors():
    from ttdg_b200 import _C, ops
    for C in (1, 4):
        a = _img(90, 130, 5 + C, C)
        ref = np.stack([_pil(np.repeat(a[:, :, c:c + 1], 3, axis=2), 40, 200)[:, :, 0] for c in range(C)], axis=2)
        out = ops.resize_bilinear_u8(torch.from_numpy(a).cuda(), 40, 200, planar=False)
        assert np.array_equal(out.cpu().numpy(), ref)
    with pytest.raises(_C.TTDGError):
        ops.resize_bilinear_u8(torch.zeros(4, 4, 3, dtype=torch.uint8), 8, 8)          # CPU tensor: no fallback
    with pytest.raises(ValueError):
        ops.resize_bilinear_u8(torch.zeros(4, 4, 3, device="cuda"), 8, 8)              # not uint8
    with pytest.raises(ValueError):
        ops.resize_bilinear_u8(torch.zeros(4, 4, 2, dtype=torch.uint8, device="cuda"), 8, 8)      # 2 channels


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["BGR", "RGB"])
def test_dataset_mapper_device_path_equals_host_path(tmp_path, fmt):
    """DatasetMapper(cfg, False) with the resize on the device hands the detector the same uint8 C x H x W tensor as the host
    (PIL) path: the reference's ResizeShortestEdge(MIN_SIZE_TEST, MAX_SIZE_TEST) result."""
    from adapteacher.config import add_ateacher_config, get_cfg
    from adapteacher.data.build import DatasetMapper
    cfg = add_ateacher_config(get_cfg())
    cfg.INPUT.FORMAT = fmt
    cfg.INPUT.MIN_SIZE_TEST = 160
    cfg.INPUT.MAX_SIZE_TEST = 260
    for i, (H, W) in enumerate([(97, 211), (300, 120), (160, 200)]):
        path = str(tmp_path / f"im{i}.png")
        Image.fromarray(_img(H, W, 50 + i)).save(path)
        rec = {"file_name": path, "height": H, "width": W, "image_id": i, "annotations": []}
        host = DatasetMapper(cfg, False, device_resize=False)(rec)
        dev = DatasetMapper(cfg, False, device_resize=True)(rec)
        assert dev["image"].is_cuda and dev["image"].dtype == torch.uint8
        assert torch.equal(dev["image"].cpu(), host["image"])
        assert dev["height"] == H and dev["width"] == W and "annotations" not in dev
