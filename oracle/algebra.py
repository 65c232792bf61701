"""Type-generic restatement of the reference's symbolic algebra (test oracle only).

Follows /root/reference/src/BasisFunctions.jl.  Works for ``float`` (IEEE double, the
reference's Float64 path, same operation order as far as the Julia source fixes it)
and for ``fractions.Fraction`` (the reference's Rational{BigInt} path, exact).

Third-party arithmetic restated here (Polynomials.jl, compat 4.1, version unpinned by
the reference): polynomial product = coefficient convolution, ``integrate`` = term-wise
antiderivative with zero constant, ``derivative`` = term-wise, evaluation = Horner.

Conventions (BasisFunctions.jl:30-42, 86-98, 284): a Fourier mode is
c*cos(a|j|x) for j<0, c for j==0, c*sin(a j x) for j>0; Fourier inner products are
box *averages* (1 or 1/2); the wall-normal inner product is (1/2) * int_{-1}^{1}.
Polynomial parity tag: +1 even, -1 odd, 0 neither.
"""
from __future__ import annotations

from dataclasses import dataclass
from fractions import Fraction
from typing import List, Sequence, Tuple


# --------------------------------------------------------------------------- numbers
def half_of(c):
    """c/2 in c's own number type (BasisFunctions.jl:142, ``T(c) / 2``)."""
    return c / 2


def is_zero(c) -> bool:
    return c == 0


# ----------------------------------------------------------------------- polynomials
class Poly:
    """Monomial-coefficient polynomial in y; coeffs[n] multiplies y**n."""

    __slots__ = ("c",)

    def __init__(self, coeffs: Sequence):
        c = list(coeffs)
        while len(c) > 1 and c[-1] == 0:  # Polynomials.jl chops trailing zeros
            c.pop()
        if not c:
            c = [0]
        self.c = c

    def __mul__(self, other):
        if isinstance(other, Poly):
            p, q = self.c, other.c
            zero = p[0] * 0
            out = [zero] * (len(p) + len(q) - 1)
            for i, pi in enumerate(p):
                for j, qj in enumerate(q):
                    out[i + j] = out[i + j] + pi * qj
            return Poly(out)
        return Poly([a * other for a in self.c])

    __rmul__ = __mul__

    def __sub__(self, other: "Poly"):
        n = max(len(self.c), len(other.c))
        zero = self.c[0] * 0
        a = self.c + [zero] * (n - len(self.c))
        b = other.c + [zero] * (n - len(other.c))
        return Poly([x - y for x, y in zip(a, b)])

    def __add__(self, other: "Poly"):
        n = max(len(self.c), len(other.c))
        zero = self.c[0] * 0
        a = self.c + [zero] * (n - len(self.c))
        b = other.c + [zero] * (n - len(other.c))
        return Poly([x + y for x, y in zip(a, b)])

    def derivative(self, n: int = 1) -> "Poly":
        c = self.c
        for _ in range(n):
            c = [k * c[k] for k in range(1, len(c))] or [c[0] * 0]
        return Poly(c)

    def integrate(self) -> "Poly":
        zero = self.c[0] * 0
        return Poly([zero] + [a / (k + 1) for k, a in enumerate(self.c)])

    def __call__(self, y):
        acc = self.c[-1]
        for a in reversed(self.c[:-1]):
            acc = acc * y + a
        return acc

    def degree(self) -> int:
        return len(self.c) - 1

    def __repr__(self):
        return f"Poly({self.c})"


def polyparity(p: Poly) -> int:
    """BasisFunctions.jl:17-25 (code convention: +1 even, -1 odd, 0 neither)."""
    if sum(abs(a) for a in p.c[0::2]) == 0:
        return -1
    if sum(abs(a) for a in p.c[1::2]) == 0:
        return 1
    return 0


def poly_innerproduct(f: Poly, g: Poly, fparity: int = 0, gparity: int = 0):
    """(1/2) int_{-1}^{1} f g dy  --  BasisFunctions.jl:278-285."""
    if fparity * gparity == -1:
        return f.c[0] * 0
    F = (f * g).integrate()
    return (F(1) - F(-1)) / 2


# --------------------------------------------------------------------- Fourier modes
@dataclass(frozen=True)
class FourierMode:
    """BasisFunctions.jl:39-51."""
    coeff: object
    waveindex: int
    wavenumber: object

    def __call__(self, x):
        import math
        if self.waveindex < 0:
            return self.coeff * math.cos(self.wavenumber * self.waveindex * x)
        if self.waveindex > 0:
            return self.coeff * math.sin(self.wavenumber * self.waveindex * x)
        return self.coeff


def fm_compatible(a: FourierMode, b: FourierMode) -> bool:
    return a.wavenumber == b.wavenumber


def fm_isorthogonal(a: FourierMode, b: FourierMode) -> bool:
    return a.waveindex != b.waveindex


def fm_innerproduct(a: FourierMode, b: FourierMode):
    """BasisFunctions.jl:87-99."""
    if not fm_compatible(a, b):
        raise ValueError(f"incompatible ej={a}, ek={b}")
    if fm_isorthogonal(a, b) or a.coeff == 0 or b.coeff == 0:
        return a.coeff * 0
    if a.waveindex == 0:
        return a.coeff * b.coeff
    one = a.coeff * 0 + 1
    return one / 2 * a.coeff * b.coeff


def fm_derivative(e: FourierMode) -> FourierMode:
    """d/dx E_j = c (a j) E_{-j}  --  BasisFunctions.jl:105-107."""
    return FourierMode(e.coeff * e.wavenumber * e.waveindex, -e.waveindex, e.wavenumber)


def _sign(n: int) -> int:
    return (n > 0) - (n < 0)


def fm_product(ej: FourierMode, ek: FourierMode) -> List[FourierMode]:
    """Trig product-to-sum  --  BasisFunctions.jl:133-212 (1 or 2 terms, same order)."""
    if not fm_compatible(ej, ek):
        raise ValueError("incompatible ej, ek")
    j, k, a = ej.waveindex, ek.waveindex, ej.wavenumber
    c = ej.coeff * ek.coeff
    c2 = half_of(c)
    jksum = abs(j) + abs(k)
    jkdiff = abs(abs(j) - abs(k))
    jkdiffsign = _sign(abs(j) - abs(k))
    if j == 0:
        return [FourierMode(c, k, a)]
    if k == 0:
        return [FourierMode(c, j, a)]
    if j < 0 and k < 0:
        return [FourierMode(c2, -jkdiff, a), FourierMode(c2, -jksum, a)]
    if j > 0 and k > 0:
        return [FourierMode(c2, -jkdiff, a), FourierMode(-c2, -jksum, a)]
    if j + k == 0:
        return [FourierMode(c2, abs(2 * j), a)]
    if j < 0 and k > 0:
        return [FourierMode(c2, jksum, a), FourierMode(-jkdiffsign * c2, jkdiff, a)]
    if j > 0 and k < 0:
        return [FourierMode(c2, jksum, a), FourierMode(jkdiffsign * c2, jkdiff, a)]
    raise AssertionError("unreachable")


# ------------------------------------------------------------------ basis components
@dataclass(frozen=True)
class BasisComponent:
    """c * E_j(x) * E_k(z) * p(y) with a parity tag  --  BasisFunctions.jl:222-252."""
    coeff: object
    ejx: FourierMode
    ekz: FourierMode
    p: Poly
    pparity: int

    def __call__(self, x, y, z):
        return self.coeff * self.ejx(x) * self.ekz(z) * self.p(y)


def bc_compatible(f: BasisComponent, g: BasisComponent) -> bool:
    return fm_compatible(f.ejx, g.ejx) and fm_compatible(f.ekz, g.ekz)


def bc_isorthogonal(f: BasisComponent, g: BasisComponent) -> bool:
    """BasisFunctions.jl:268."""
    return (fm_isorthogonal(f.ejx, g.ejx) or fm_isorthogonal(f.ekz, g.ekz)
            or f.pparity * g.pparity == -1)


def bc_innerproduct(f: BasisComponent, g: BasisComponent):
    """BasisFunctions.jl:290-298."""
    if not bc_compatible(f, g):
        raise ValueError("incompatible f,g")
    if bc_isorthogonal(f, g) or f.coeff == 0 or g.coeff == 0 or f.pparity * g.pparity == -1:
        return f.coeff * 0
    return (f.coeff * g.coeff * fm_innerproduct(f.ejx, g.ejx)
            * fm_innerproduct(f.ekz, g.ekz) * poly_innerproduct(f.p, g.p))


def bc_xderivative(f: BasisComponent) -> BasisComponent:
    return BasisComponent(f.coeff, fm_derivative(f.ejx), f.ekz, f.p, f.pparity)


def bc_yderivative(f: BasisComponent) -> BasisComponent:
    return BasisComponent(f.coeff, f.ejx, f.ekz, f.p.derivative(), -f.pparity)


def bc_zderivative(f: BasisComponent) -> BasisComponent:
    return BasisComponent(f.coeff, f.ejx, fm_derivative(f.ekz), f.p, f.pparity)


def bc_scale(c, f: BasisComponent) -> BasisComponent:
    return BasisComponent(c * f.coeff, f.ejx, f.ekz, f.p, f.pparity)


def _square(x):
    return x * x


def bc_laplacian(f: BasisComponent) -> BasisComponent:
    """BasisFunctions.jl:316-320."""
    pyy = f.p.derivative(2)
    k2 = (_square(f.ejx.wavenumber * f.ejx.waveindex)
          + _square(f.ekz.wavenumber * f.ekz.waveindex))
    return BasisComponent(f.coeff, f.ejx, f.ekz, pyy - k2 * f.p, f.pparity)


def bc_product(f: BasisComponent, g: BasisComponent) -> List[BasisComponent]:
    """Component product, 1-4 terms  --  BasisFunctions.jl:328-356."""
    fgcoeff = f.coeff * g.coeff
    fgejx = fm_product(f.ejx, g.ejx)
    fgekz = fm_product(f.ekz, g.ekz)
    fgpoly = f.p * g.p
    fgpparity = f.pparity * g.pparity
    return [BasisComponent(fgcoeff, ex, ez, fgpoly, fgpparity) for ex in fgejx for ez in fgekz]


# -------------------------------------------------------------------- basis functions
BasisFunction = Tuple[BasisComponent, BasisComponent, BasisComponent]


def bf_innerproduct(f: BasisFunction, g: BasisFunction):
    """BasisFunctions.jl:399-405."""
    return bc_innerproduct(f[0], g[0]) + bc_innerproduct(f[1], g[1]) + bc_innerproduct(f[2], g[2])


def bf_norm(f: BasisFunction):
    import math
    return math.sqrt(bf_innerproduct(f, f))


def bf_scale(c, f: BasisFunction) -> BasisFunction:
    return (bc_scale(c, f[0]), bc_scale(c, f[1]), bc_scale(c, f[2]))


def bf_xderivative(f: BasisFunction) -> BasisFunction:
    return (bc_xderivative(f[0]), bc_xderivative(f[1]), bc_xderivative(f[2]))


def bf_laplacian(f: BasisFunction) -> BasisFunction:
    return (bc_laplacian(f[0]), bc_laplacian(f[1]), bc_laplacian(f[2]))


def bf_polymul(q: Poly, f: BasisFunction) -> BasisFunction:
    """BasisFunctions.jl:420-426."""
    qp = polyparity(q)
    return tuple(BasisComponent(u.coeff, u.ejx, u.ekz, q * u.p, qp * u.pparity) for u in f)


def bf_vex(f: BasisFunction) -> BasisFunction:
    """[v, 0, 0]  --  BasisFunctions.jl:431-436."""
    zero = f[1].coeff * 0
    u = BasisComponent(f[1].coeff, f[1].ejx, f[1].ekz, f[1].p, f[1].pparity)
    v = BasisComponent(zero, f[1].ejx, f[1].ekz, Poly([zero]), 0)
    w = BasisComponent(zero, f[2].ejx, f[2].ekz, Poly([zero]), 0)
    return (u, v, w)


def bf_dotgrad(f: BasisFunction, g: BasisFunction) -> List[List[BasisComponent]]:
    """(f . grad) g as three lists of product terms  --  BasisFunctions.jl:444-451."""
    out = []
    for c in range(3):
        out.append(bc_product(f[0], bc_xderivative(g[c]))
                   + bc_product(f[1], bc_yderivative(g[c]))
                   + bc_product(f[2], bc_zderivative(g[c])))
    return out


def bf_innerproduct_terms(f: BasisFunction, g: List[List[BasisComponent]]):
    """BasisFunctions.jl:453-463."""
    s = f[0].coeff * 0
    for i in range(3):
        for term in g[i]:
            s = s + bc_innerproduct(f[i], term)
    return s


__all__ = [n for n in dir() if not n.startswith("_")]
