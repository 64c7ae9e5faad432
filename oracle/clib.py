"""ctypes loader for oracle/_build/liblap_ref.so (TEST INFRASTRUCTURE, see oracle/__init__.py)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblap_ref.so")
_lib = None


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.ttdg_oracle_lsap.restype = ctypes.c_int
        _lib.ttdg_oracle_hungarian.restype = ctypes.c_int
    return _lib


def lsap(cost):
    cost = np.ascontiguousarray(cost, dtype=np.float64)
    nr, nc = cost.shape
    m = min(nr, nc)
    rows = np.zeros(m, dtype=np.int64)
    cols = np.zeros(m, dtype=np.int64)
    rc = lib().ttdg_oracle_lsap(ctypes.c_int(nr), ctypes.c_int(nc), cost.ctypes.data_as(ctypes.c_void_p),
                                rows.ctypes.data_as(ctypes.c_void_p), cols.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise ValueError("cost matrix is infeasible")
    return rows, cols


def hungarian(s):
    s = np.ascontiguousarray(s, dtype=np.float32)
    perm = np.zeros_like(s)
    rc = lib().ttdg_oracle_hungarian(ctypes.c_int(s.shape[0]), ctypes.c_int(s.shape[1]),
                                     s.ctypes.data_as(ctypes.c_void_p), perm.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise ValueError("cost matrix is infeasible")
    return perm
