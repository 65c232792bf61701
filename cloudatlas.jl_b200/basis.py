"""Symmetry, basisIndices, basis component table -- mirror of src/Symmetries.jl and the index/basis
part of src/BasisFunctions.jl, over the C ABI (host-side integer work done by the library)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from fractions import Fraction

import numpy as np

from ._lib import CloudAtlasError, _sig, c_f64p, c_i32p, c_i64p, check


class _CSym(C.Structure):
    _fields_ = [("sx", C.c_int32), ("sy", C.c_int32), ("sz", C.c_int32),
                ("ax2", C.c_int32), ("az2", C.c_int32), ("s", C.c_int32)]


_indices = _sig("ca_basis_indices", C.c_int32, C.c_int32, C.c_int32, C.POINTER(_CSym), C.c_int32, c_i64p, c_i64p)
_components = _sig("ca_basis_components", C.c_double, C.c_double, c_i64p, C.c_int64, C.c_int32,
                   c_f64p, c_i32p, c_i32p, c_i32p, c_i32p)


@dataclass(frozen=True)
class Symmetry:
    """Symmetry(sx, sy, sz, ax, az, s) -- Symmetries.jl:3-14."""
    sx: int = 1
    sy: int = 1
    sz: int = 1
    ax: Fraction = Fraction(0)
    az: Fraction = Fraction(0)
    s: int = 1

    def __mul__(self, o: "Symmetry") -> "Symmetry":
        # group product with shifts mod 1 -- Symmetries.jl:16-17
        return Symmetry(self.sx * o.sx, self.sy * o.sy, self.sz * o.sz,
                        (Fraction(self.ax) + Fraction(o.ax)) % 1, (Fraction(self.az) + Fraction(o.az)) % 1,
                        self.s * o.s)

    def _c(self) -> _CSym:
        ax2, az2 = Fraction(self.ax) * 2, Fraction(self.az) * 2
        if ax2.denominator != 1 or az2.denominator != 1:
            raise CloudAtlasError(1, "symmetries only implemented for half-box shifts")  # Symmetries.jl:36,42
        return _CSym(self.sx, self.sy, self.sz, int(ax2), int(az2), self.s)


def halfbox_symmetries():
    """sx, sy, sz, tx, tz -- Symmetries.jl:68-75."""
    return (Symmetry(-1, 1, 1), Symmetry(1, -1, 1), Symmetry(1, 1, -1),
            Symmetry(1, 1, 1, Fraction(1, 2), Fraction(0)), Symmetry(1, 1, 1, Fraction(0), Fraction(1, 2)))


def _sym_array(H):
    H = [] if H is None else ([H] if isinstance(H, Symmetry) else list(H))
    arr = (_CSym * max(1, len(H)))()
    for n, g in enumerate(H):
        arr[n] = g._c()
    return arr, len(H)


_symmetric = _sig("ca_symmetric", c_i64p, C.POINTER(_CSym), C.c_int32, C.POINTER(C.c_int32))


def basisIndices(J: int, K: int, L: int, H=None) -> np.ndarray:
    """m x 4 int64 matrix of (i,j,k,l) -- BasisFunctions.jl:489-574."""
    arr, nH = _sym_array(H)
    m = C.c_int64(0)
    check(_indices(J, K, L, arr, nH, None, C.byref(m)))
    out = np.zeros((m.value, 4), dtype=np.int64, order="F")
    check(_indices(J, K, L, arr, nH, out.ctypes.data_as(c_i64p), C.byref(m)))
    return out


def symmetric(ijkl, sigma) -> bool:
    """symmetric(ijkl, sigma) for one Symmetry or a list of generators -- Symmetries.jl:25-60: the sign computation
    on the row itself (``ca_symmetric``), for any (i,j,k,l) with i in 1:6."""
    row = np.ascontiguousarray([int(v) for v in ijkl], dtype=np.int64)
    arr, nH = _sym_array(sigma)
    out = C.c_int32(0)
    check(_symmetric(row.ctypes.data_as(c_i64p), arr, nH, C.byref(out)))
    return bool(out.value)


def basisComponents(alpha, gamma, ijkl, normalize=False):
    """Tabular basisSet -- BasisFunctions.jl:630-702: dict of m x 3 arrays coef, jx, kz, l, d."""
    ijkl = np.asfortranarray(ijkl, dtype=np.int64)
    m = ijkl.shape[0]
    coef = np.zeros((m, 3), order="F")
    ints = [np.zeros((m, 3), dtype=np.int32, order="F") for _ in range(4)]
    check(_components(float(alpha), float(gamma), ijkl.ctypes.data_as(c_i64p), m, int(bool(normalize)),
                      coef.ctypes.data_as(c_f64p), *[a.ctypes.data_as(c_i32p) for a in ints]))
    return {"coef": coef, "jx": ints[0], "kz": ints[1], "l": ints[2], "d": ints[3]}


_shear = _sig("ca_basis_shear", C.c_double, C.c_double, c_i64p, C.c_int64, C.c_int32, c_f64p)


def shearVector(alpha, gamma, ijkl, normalize=False):
    """Psi_shear of shear_function (ODEModels.jl:65-78): shear(x) = 1 + dot(x, Psi_shear)."""
    ijkl = np.asfortranarray(ijkl, dtype=np.int64)
    out = np.zeros(ijkl.shape[0])
    check(_shear(float(alpha), float(gamma), ijkl.ctypes.data_as(c_i64p), ijkl.shape[0], int(bool(normalize)),
                 out.ctypes.data_as(c_f64p)))
    return out
