"""oracle/lap_ref.c (restated SciPy rectangular LSAP) vs the installed SciPy and
vs the reference's hungarian() golden outputs."""
import numpy as np
import pytest
import scipy.optimize

from oracle import clib


def _cases():
    rng = np.random.default_rng(0)
    for nr in range(1, 14):
        for nc in range(1, 14):
            yield rng.standard_normal((nr, nc)).astype(np.float32).astype(np.float64)
            yield rng.integers(0, 3, (nr, nc)).astype(np.float64)          # tie-heavy
            yield np.full((nr, nc), 1.5)                                    # constant
    for n in (23, 32, 40, 57, 90):                                          # reference shapes n x 32
        yield rng.standard_normal((n, 32)).astype(np.float32).astype(np.float64)
        yield rng.standard_normal((32, n)).astype(np.float32).astype(np.float64)


def test_lsap_matches_scipy():
    n = 0
    for cost in _cases():
        r0, c0 = scipy.optimize.linear_sum_assignment(cost)
        r1, c1 = clib.lsap(cost)
        assert np.array_equal(r0, r1) and np.array_equal(c0, c1), cost.shape
        n += 1
    assert n > 500


def test_hungarian_wrapper_matches_reference_golden(golden_dir):
    g = np.load(f"{golden_dir}/ops.npz")
    for name in ("hung_40x32", "hung_20x32", "hung_32x32", "hung_23x57", "hung_ties"):
        out = clib.hungarian(g[name + "_in"])
        assert np.array_equal(out.astype(np.uint8), g[name + "_out"]), name
