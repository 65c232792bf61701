"""ctypes front end of oracle/csrc/libref_loops.so (test / CPU-baseline infrastructure only)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "csrc", "libref_loops.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], check=True)
        _lib = C.CDLL(path)
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class CRefModel:
    """B, A1, A2 dense + N COO; f(x,R) and the ensemble form through the C loops."""

    def __init__(self, B, A1, A2, ijk, val):
        self.m = B.shape[0]
        self.A1 = np.asfortranarray(A1, dtype=np.float64)
        self.A2 = np.asfortranarray(A2, dtype=np.float64)
        self.ijk = np.asfortranarray(ijk, dtype=np.int64)
        self.val = np.ascontiguousarray(val, dtype=np.float64)
        self.LU = np.asfortranarray(B, dtype=np.float64).copy(order="F")
        self.piv = np.zeros(self.m, dtype=np.int64)
        if lib().ref_lu_factor(_p(self.LU, C.c_double), C.c_int64(self.m), _p(self.piv, C.c_int64)) != 0:
            raise ValueError("B is singular")

    def N(self, x, y=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = x if y is None else np.ascontiguousarray(y, dtype=np.float64)
        out = np.empty(self.m)
        lib().ref_bilinear_eval(_p(self.ijk, C.c_int64), _p(self.val, C.c_double), C.c_int64(len(self.val)),
                                C.c_int64(self.m), _p(x, C.c_double), _p(y, C.c_double), _p(out, C.c_double))
        return out

    def derivative(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        DN = np.empty((self.m, self.m), order="F")
        lib().ref_jacobian(_p(self.ijk, C.c_int64), _p(self.val, C.c_double), C.c_int64(len(self.val)),
                           C.c_int64(self.m), _p(x, C.c_double), _p(DN, C.c_double))
        return DN

    def f_ensemble(self, X, R, threads=1):
        X = np.asfortranarray(X, dtype=np.float64)
        OUT = np.empty(X.shape, order="F")
        lib().ref_rhs_ensemble(_p(self.LU, C.c_double), _p(self.piv, C.c_int64), _p(self.A1, C.c_double),
                               _p(self.A2, C.c_double), _p(self.ijk, C.c_int64), _p(self.val, C.c_double),
                               C.c_int64(len(self.val)), C.c_int64(self.m), _p(X, C.c_double),
                               C.c_int64(X.shape[0]), C.c_double(R), _p(OUT, C.c_double), C.c_int(threads))
        return OUT

    def f(self, x, R):
        return self.f_ensemble(np.asarray(x, dtype=np.float64).reshape(1, -1), R)[0]
