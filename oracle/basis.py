"""Index sets, symmetry filter and basis construction (test oracle only).

Follows /root/reference/src/Symmetries.jl:3-60 and
/root/reference/src/BasisFunctions.jl:471-476 (fourierIndices), :489-574 (basisIndices),
:580-591 (legendrePolynomials), :630-702 (basisSet), :847-907 (parity tables).
Integer work is bit-exact by construction (plain Python ints; Julia's truncating ``%``
is reproduced with ``math.fmod`` semantics where operands may be negative).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from fractions import Fraction
from typing import List, Sequence

import numpy as np

from .algebra import (BasisComponent, FourierMode, Poly, bf_innerproduct, bf_scale)


# ------------------------------------------------------------------------ symmetries
@dataclass(frozen=True)
class Symmetry:
    """Symmetries.jl:3-14.  ax, az are exact rationals (half-box shifts)."""
    sx: int = 1
    sy: int = 1
    sz: int = 1
    ax: Fraction = Fraction(0)
    az: Fraction = Fraction(0)
    s: int = 1

    def __mul__(self, o: "Symmetry") -> "Symmetry":
        # Symmetries.jl:16-17 ; Julia % on non-negative rationals == Python %
        return Symmetry(self.sx * o.sx, self.sy * o.sy, self.sz * o.sz,
                        (Fraction(self.ax) + Fraction(o.ax)) % 1,
                        (Fraction(self.az) + Fraction(o.az)) % 1, self.s * o.s)


def halfbox_symmetries():
    """Symmetries.jl:68-75."""
    return (Symmetry(-1, 1, 1), Symmetry(1, -1, 1), Symmetry(1, 1, -1),
            Symmetry(1, 1, 1, Fraction(1, 2), Fraction(0)),
            Symmetry(1, 1, 1, Fraction(0), Fraction(1, 2)))


def _jrem(a: int, b: int) -> int:
    """Julia's ``%`` (truncated remainder, sign of the dividend)."""
    return int(math.fmod(a, b))


def xreflection(ijkl) -> int:
    i, j, k, l = ijkl
    if i == 1:
        return -1
    if i == 2:
        return 1
    if 3 <= i <= 6:
        return 1 if j <= 0 else -1
    raise ValueError(f"invalid index i = {i} not in 1:6")


def yreflection(ijkl) -> int:
    return 1 if _jrem(ijkl[3], 2) == 0 else -1


def zreflection(ijkl) -> int:
    i, j, k, l = ijkl
    if i == 3:
        return -1
    if i == 4:
        return 1
    if i in (1, 2, 5, 6):
        return 1 if k <= 0 else -1
    raise ValueError(f"invalid index i = {i} not in 1:6")


def xtranslationLx2(ijkl) -> int:
    return 1 if _jrem(ijkl[1], 2) == 0 else -1


def ztranslationLz2(ijkl) -> int:
    return 1 if _jrem(ijkl[2], 2) == 0 else -1


def symmetric(ijkl, sigma) -> bool:
    """Symmetries.jl:25-60 (single symmetry or a list of generators)."""
    if isinstance(sigma, (list, tuple)):
        return all(symmetric(ijkl, s) for s in sigma)
    r = sigma.s
    r *= xreflection(ijkl) if sigma.sx == -1 else 1
    r *= yreflection(ijkl) if sigma.sy == -1 else 1
    r *= zreflection(ijkl) if sigma.sz == -1 else 1
    if sigma.ax == Fraction(1, 2):
        r *= xtranslationLx2(ijkl)
    elif sigma.ax != 0:
        raise ValueError("symmetries only implemented for half-box shifts")
    if sigma.az == Fraction(1, 2):
        r *= ztranslationLz2(ijkl)
    elif sigma.az != 0:
        raise ValueError("symmetries only implemented for half-box shifts")
    return r == 1


# ---------------------------------------------------------------------------- indices
def fourierIndices(J: int) -> List[int]:
    """Order [0, -1, 1, -2, 2, ...]  --  BasisFunctions.jl:471-476."""
    out = [0]
    for n in range(1, J + 1):
        out += [-n, n]
    return out


def basisIndices(J: int, K: int, L: int, H: Sequence[Symmetry] | None = None) -> np.ndarray:
    """N x 4 int64 matrix of (i,j,k,l)  --  BasisFunctions.jl:489-574."""
    rows = []
    for j in fourierIndices(J):
        for k in fourierIndices(K):
            if j == 0:
                rows += [(1, 0, k, l) for l in range(0, L + 1)]
            if j == 0 and k != 0:
                rows += [(2, 0, k, l) for l in range(1, L + 1)]
            if k == 0:
                rows += [(3, j, 0, l) for l in range(0, L + 1)]
            if k == 0 and j != 0:
                rows += [(4, j, 0, l) for l in range(1, L + 1)]
            if j != 0 and k != 0:
                rows += [(5, j, k, l) for l in range(0, L + 1)]
                rows += [(6, j, k, l) for l in range(1, L + 1)]
    assert len(rows) == (2 * J + 1) * (2 * K + 1) * (2 * L + 1) + 1
    if H is not None:
        rows = [r for r in rows if symmetric(r, list(H))]
    return np.array(rows, dtype=np.int64).reshape(-1, 4)


# ------------------------------------------------------------------------ polynomials
def legendrePolynomials(T, L: int) -> List[Poly]:
    """P_0..P_L in monomial form  --  BasisFunctions.jl:580-591 (incl. the L==1 quirk)."""
    if L < 0:
        raise ValueError("invalid L")
    P = [Poly([T(0)]) for _ in range(L + 1)]
    P[0] = Poly([T(1)])
    if L > 1:
        P[1] = Poly([T(0), T(1)])
    for n in range(2, L + 1):
        # (1//n) * ((2n-1) * P1 * P[n-1] - (n-1) * P[n-2])
        inner = (P[1] * (2 * n - 1)) * P[n - 1] - P[n - 2] * (n - 1)
        # Julia promotes the Rational 1//n to T first: coefficient = T(1//n) * a
        recip = Fraction(1, n) if T is Fraction else 1.0 / n
        P[n] = Poly([recip * a for a in inner.c])
    return P


def wallPolynomials(T, L: int):
    """S_0..S_L and S'_0..S'_L  --  BasisFunctions.jl:636-642.

    S_0 = y - y^3/3 ; S_l = (1-y^2)^2 P_{l-1}, l>=1.
    """
    smasher = Poly([T(1), T(0), T(-1)])
    smasher2 = smasher * smasher
    P = legendrePolynomials(T, L)
    third = Fraction(1, 3) if T is Fraction else (1.0 / 3.0)
    S = [Poly([T(0), T(1), T(0), -third])] + [smasher2 * P[l] for l in range(0, L)]
    Sp = [s.derivative() for s in S]
    return S, Sp


def parity(k: int) -> int:
    return 1 if _jrem(k, 2) == 0 else -1


def basisSet(alpha, gamma, ijkl: np.ndarray, normalize: bool = False):
    """List of 3-tuples of BasisComponent  --  BasisFunctions.jl:630-702."""
    T = Fraction if isinstance(alpha, Fraction) else float
    alpha, gamma = T(alpha), T(gamma)
    L = int(ijkl[:, 3].max())
    S, Sp = wallPolynomials(T, L)
    one, zero = T(1), T(0)
    Ex = lambda j: FourierMode(one, int(j), alpha)
    Ez = lambda k: FourierMode(one, int(k), gamma)
    zerocomp = BasisComponent(zero, FourierMode(zero, 0, alpha), FourierMode(zero, 0, gamma),
                              Poly([zero]), 0)
    Psi = []
    for n in range(ijkl.shape[0]):
        i, j, k, l = (int(v) for v in ijkl[n])
        if i == 1:
            f = (BasisComponent(one, Ex(0), Ez(k), Sp[l], parity(l)), zerocomp, zerocomp)
        elif i == 2:
            f = (zerocomp,
                 BasisComponent(gamma * k, Ex(0), Ez(k), S[l], parity(l + 1)),
                 BasisComponent(one, Ex(0), Ez(-k), Sp[l], parity(l)))
        elif i == 3:
            f = (zerocomp, zerocomp, BasisComponent(one, Ex(j), Ez(0), Sp[l], parity(l)))
        elif i == 4:
            f = (BasisComponent(one, Ex(-j), Ez(0), Sp[l], parity(l)),
                 BasisComponent(alpha * j, Ex(j), Ez(0), S[l], parity(l + 1)),
                 zerocomp)
        elif i == 5:
            f = (BasisComponent(gamma * k, Ex(-j), Ez(k), Sp[l], parity(l)),
                 zerocomp,
                 BasisComponent(-alpha * j, Ex(j), Ez(-k), Sp[l], parity(l)))
        elif i == 6:
            f = (BasisComponent(gamma * k, Ex(-j), Ez(k), Sp[l], parity(l)),
                 BasisComponent(2 * alpha * gamma * j * k, Ex(j), Ez(k), S[l], parity(l + 1)),
                 BasisComponent(alpha * j, Ex(j), Ez(-k), Sp[l], parity(l)))
        else:
            raise ValueError(f"bad mode family {i}")
        if normalize:
            if T is not float:
                raise TypeError("normalize needs sqrt; float only (as in the reference)")
            f = bf_scale(1 / math.sqrt(bf_innerproduct(f, f)), f)
        Psi.append(f)
    return Psi
