"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/ttdg_b200.h declares (no compute calls - there is no GPU here), the ctypes table mirrors the header, and
the product path refuses to run without CUDA (no fallback)."""
import os
import re

import pytest
import torch

from ttdg_b200 import _C, _build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    _build.build()
    return _C.lib()


def _declared():
    hdr = open(os.path.join(ROOT, "include", "ttdg_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return set(re.findall(r"\b(ttdg_[a-z0-9_]+)\s*\(", hdr))


def test_library_exports_every_declared_symbol(lib):
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    assert names == set(_C.SIGNATURES)


def test_host_only_entry_points(lib):
    assert lib.ttdg_version() >= 100
    assert b"sm_100a" in lib.ttdg_build_info()
    assert _C.limit("small_max_dim") == 96 and _C.limit("univ") == 32 and _C.limit("nope") == -1
    assert lib.ttdg_gagm_scratch_bytes(160, 4) >= 3 * 160 * 32 * 8
    assert lib.ttdg_sinkhorn_stream_scratch_bytes(8, 1024, 1024) == 0


def test_argument_errors_do_not_touch_the_gpu(lib):
    assert lib.ttdg_sinkhorn_small_fwd(None, None, None, 1, 8, 0.05, 20, 1, None) == -1
    assert lib.ttdg_lap_solve(None, None, None, 1, None) == -1
    assert lib.ttdg_sinkhorn_stream_fwd(1, 1, 1, 8, 4, 0.05, 5, 0, None, None) == -2     # n1 > n2


def test_no_cpu_fallback():
    from ttdg_b200 import ops
    with pytest.raises(_C.TTDGError):
        ops.sinkhorn(torch.randn(4, 4))
    with pytest.raises(_C.TTDGError):
        ops.hungarian(torch.randn(4, 4))
    with pytest.raises(_C.TTDGError):
        ops.linear(torch.randn(4, 4), torch.randn(4, 4))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ttdg-mgm_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dp, f)


def test_mirror_state_dict_keys_match_reference_names():
    from adapteacher.modeling.GModule.multi_graph_matching import MGM3_unsup, U_sup
    from ttdg_b200 import synth
    m = MGM3_unsup(2, 32)
    assert m.load_state_dict(synth.mgm_unsup_state(0), strict=True)
    assert tuple(U_sup(2, 32).U.shape) == (32, 256)
