"""Import the reference's own MGM Python, unmodified, from /root/reference.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Only usable in the build container
(/root/reference does not exist on the GPU box): used by ``gen_golden.py`` to make
``tests/golden/*.npz`` and by ``tests/test_oracle_vs_reference.py`` (skipped when the
reference tree is absent) to pin ``mgm_port.py``.

Three shims: (i) bare ``adapteacher`` / ``adapteacher.modeling``
packages so ``adapteacher/__init__.py:2`` (imports Detectron2) is bypassed; (ii) stub
``matplotlib`` (imported, unused, at ``multi_graph_matching.py:6-7``); (iii) a
``pygmtools`` module exposing the restated ``sinkhorn``.
"""
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("TTDG_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "adapteacher", "modeling", "GModule"))


def load():
    """Returns a namespace with the reference classes.  Must run in a process that has NOT
    put ``ttdg-mgm_b200/`` (our own ``adapteacher`` mirror) on sys.path before."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if "adapteacher" in sys.modules and not getattr(sys.modules["adapteacher"], "_ttdg_ref_shim", False):
        raise RuntimeError("another 'adapteacher' package is already imported in this process")
    from oracle import pygm_sinkhorn

    pg = types.ModuleType("pygmtools")
    pg.sinkhorn = pygm_sinkhorn.sinkhorn
    sys.modules["pygmtools"] = pg
    if "matplotlib" not in sys.modules:
        mpl, plt, col = (types.ModuleType(n) for n in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors"))
        col.ListedColormap = object
        mpl.pyplot, mpl.colors = plt, col
        sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt, "matplotlib.colors": col})
    for name, rel in (("adapteacher", "adapteacher"), ("adapteacher.modeling", "adapteacher/modeling")):
        if name not in sys.modules:
            pkg = types.ModuleType(name)
            pkg.__path__ = [os.path.join(REFERENCE_ROOT, rel)]
            pkg._ttdg_ref_shim = True
            sys.modules[name] = pkg
    import warnings
    warnings.filterwarnings("ignore", message=".*chain_matmul.*")
    warnings.filterwarnings("ignore", message=".*meshgrid.*")
    from adapteacher.modeling.GModule import multi_graph_matching as mgm
    from adapteacher.modeling.GModule import build_graph
    from adapteacher.modeling.GModule.utils import sinkhorn as sk, hungarian as hg, affinity as af
    from adapteacher.modeling.GModule.utils import attentions as at, losses as ls, pad_tensor as pt

    ns = types.SimpleNamespace(
        mgm=mgm, build_graph=build_graph, sinkhorn=sk, hungarian=hg, affinity=af, attentions=at,
        losses=ls, pad_tensor=pt,
        MGM3_unsup=mgm.MGM3_unsup, U_sup=mgm.U_sup, GA_GM=mgm.GA_GM, HiPPI=mgm.HiPPI,
        PrototypeComputation=build_graph.PrototypeComputation, Sinkhorn=sk.Sinkhorn,
        hungarian_fn=hg.hungarian, Affinity=af.Affinity, MultiHeadAttention=at.MultiHeadAttention,
        PermutationLoss=ls.PermutationLoss, BCEFocalLoss=ls.BCEFocalLoss,
    )
    return ns


class MaskedDropout(torch.nn.Module):
    """Drop-in for the ``nn.Dropout`` inside the reference's ``dot_attention``
    (``utils/attentions.py:29,40``): applies pre-drawn keep masks (0/1) scaled by 1/(1-p), in
    call order, so the stochastic adjacency can be pinned on both sides."""

    def __init__(self, masks, p=0.1):
        super().__init__()
        self.masks = list(masks)
        self.p = p
        self.i = 0

    def forward(self, x):
        m = self.masks[self.i]
        self.i += 1
        return x * (m.reshape(x.shape).to(x.dtype) / (1.0 - self.p))  # torch: mask.div_(1-p) then mul
