"""ctypes binding of oracle/lsap.c (TEST INFRASTRUCTURE ONLY)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblsap_oracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.spe_oracle_lsap.restype = ctypes.c_int
        _lib.spe_oracle_lsap.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return _lib


def lsap_c(cost):
    """cost: array-like [nr, nc] (float32) -> (row_ind int64[k], col_ind int64[k]) like scipy."""
    c = np.ascontiguousarray(np.asarray(cost, dtype=np.float32))
    nr, nc = c.shape
    k = min(nr, nc)
    ri, ci = np.empty(k, np.int64), np.empty(k, np.int64)
    got = _load().spe_oracle_lsap(nr, nc, c.ctypes.data, ri.ctypes.data, ci.ctypes.data)
    if got < 0:
        raise ValueError("cost matrix is infeasible")
    return ri, ci
