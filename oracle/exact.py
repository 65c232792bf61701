"""Fast exact-rational assembly by separable tabulation (test oracle only).

Same mathematics as ``oracle/model.py`` (i.e. /root/reference/src/ODEModels.jl:31-49 over
the algebra of /root/reference/src/BasisFunctions.jl), evaluated in exact rational
arithmetic (the reference is type-generic and was written for Rational{BigInt},
BasisFunctions.jl:10-12) but organised so that any basis size is reachable:

* every component is  coeff * E_jx(x) * E_kz(z) * S_l^(d)(y)  (BasisFunctions.jl:655-696), so a
  triple inner product factors into (coeff product) * X3 * Z3 * Y3 where X3, Z3 are the
  closed-form Fourier triple averages (obtained here from the reference's own product and
  inner-product rules, BasisFunctions.jl:87-99,133-212) and
  Y3[p,q,r] = (1/2) int_{-1}^{1} S_p S_q S_r dy over the polynomial family
  {S_l, S_l', S_l''} (BasisFunctions.jl:278-285);
* modes that differ only in the wall-normal index l share their Fourier factors, so N is
  filled block-by-block over triples of (family, j, k) groups.

``tests/test_oracle_exact.py`` checks this module against the slow generic restatement
(``oracle/model.py`` with Fraction inputs) for *equality* at 17d/27d/44d.
The reference's absolute 1e-15 threshold (ODEModels.jl:46) is applied to the exact value;
no exactly-non-zero entry of any tested configuration comes near it.
"""
from __future__ import annotations

from fractions import Fraction
from functools import lru_cache
from math import lcm
from typing import Dict, List, Tuple

import numpy as np

from .algebra import FourierMode, fm_derivative, fm_innerproduct, fm_product
from .basis import wallPolynomials


# ----------------------------------------------------------------- component tables
def component_table(alpha, gamma, ijkl: np.ndarray):
    """Per mode: three (coeff, jx, kz, d) entries, d = 0 for S_l, 1 for S_l'; coeff 0 = absent.

    Restates BasisFunctions.jl:655-696 without building polynomials.
    """
    a, g = Fraction(alpha), Fraction(gamma)
    Z = (Fraction(0), 0, 0, 0)
    out = []
    for i, j, k, l in ijkl.tolist():
        if i == 1:
            out.append(((Fraction(1), 0, k, 1), Z, Z))
        elif i == 2:
            out.append((Z, (g * k, 0, k, 0), (Fraction(1), 0, -k, 1)))
        elif i == 3:
            out.append((Z, Z, (Fraction(1), j, 0, 1)))
        elif i == 4:
            out.append(((Fraction(1), -j, 0, 1), (a * j, j, 0, 0), Z))
        elif i == 5:
            out.append(((g * k, -j, k, 1), Z, (-a * j, j, -k, 1)))
        elif i == 6:
            out.append(((g * k, -j, k, 1), (2 * a * g * j * k, j, k, 0), (a * j, j, -k, 1)))
        else:
            raise ValueError(i)
    return out


# --------------------------------------------------------------------- Fourier factors
@lru_cache(maxsize=None)
def _fourier_pair(wn: Fraction, ja: int, jb: int, db: int) -> Fraction:
    """< E_ja , d^db E_jb >  (box average)."""
    one = Fraction(1)
    eb = FourierMode(one, jb, wn)
    for _ in range(db):
        eb = fm_derivative(eb)
    return fm_innerproduct(FourierMode(one, ja, wn), eb)


@lru_cache(maxsize=None)
def _fourier_triple(wn: Fraction, ja: int, jb: int, jc: int, dc: int) -> Fraction:
    """< E_ja , E_jb * d^dc E_jc >  (box average)."""
    one = Fraction(1)
    ec = FourierMode(one, jc, wn)
    for _ in range(dc):
        ec = fm_derivative(ec)
    ea = FourierMode(one, ja, wn)
    return sum((fm_innerproduct(ea, t) for t in fm_product(FourierMode(one, jb, wn), ec)),
               Fraction(0))


# ------------------------------------------------------------------- wall-normal tables
class WallTables:
    """Exact Y2 / Y3 tables over pid = 3*l + d  (d = derivative order 0,1,2)."""

    def __init__(self, L: int):
        self.L = L
        S, Sp = wallPolynomials(Fraction, L)
        polys = []
        for l in range(L + 1):
            polys += [S[l], Sp[l], Sp[l].derivative()]
        # common denominator -> integer coefficient vectors (python ints in object arrays)
        den = 1
        for p in polys:
            for c in p.c:
                den = lcm(den, Fraction(c).denominator)
        self.den = den
        self.deg = max(p.degree() for p in polys)
        n = self.deg + 1
        self.coef = np.zeros((len(polys), n), dtype=object)
        for r, p in enumerate(polys):
            for k, c in enumerate(p.c):
                self.coef[r, k] = int(Fraction(c) * den)
        # moments (1/2) int y^n = 1/(n+1) (n even), scaled to integers
        nm = 3 * self.deg + 2
        self.mden = 1
        for k in range(0, nm, 2):
            self.mden = lcm(self.mden, k + 1)
        self.mom = np.array([self.mden // (k + 1) if k % 2 == 0 else 0 for k in range(nm)],
                            dtype=object)
        self._y3: Dict[Tuple[int, int], np.ndarray] = {}
        self._pair: Dict[Tuple[int, int], np.ndarray] = {}
        # M[r, a] = sum_b coef[r,b] * mom[a+b]  (so that int p*q*r = dot(conv(p,q), M[r]))
        na = 2 * self.deg + 1
        self.M = np.zeros((len(polys), na), dtype=object)
        for r in range(len(polys)):
            for a_ in range(na):
                self.M[r, a_] = sum(int(self.coef[r, b]) * int(self.mom[a_ + b]) for b in range(n))

    @staticmethod
    def pid(l: int, d: int) -> int:
        return 3 * l + d

    def _conv(self, p: int, q: int) -> np.ndarray:
        key = (p, q) if p <= q else (q, p)
        v = self._pair.get(key)
        if v is None:
            n = self.deg + 1
            v = np.zeros(2 * self.deg + 1, dtype=object)
            cp, cq = self.coef[key[0]], self.coef[key[1]]
            for a_ in range(n):
                if cp[a_] != 0:
                    v[a_:a_ + n] += cp[a_] * cq
            self._pair[key] = v
        return v

    def Y3(self, p: int, q: int, r: int) -> Fraction:
        """(1/2) int S_p S_q S_r dy, exact."""
        num = int(np.dot(self._conv(p, q), self.M[r]))
        return Fraction(num, self.den ** 3 * self.mden)

    def Y3_block(self, ps, qs, rs) -> np.ndarray:
        """Object array [len(ps), len(qs), len(rs)] of exact integer numerators
        (common denominator den^3 * mden)."""
        out = np.zeros((len(ps), len(qs), len(rs)), dtype=object)
        Mr = self.M[list(rs)]  # [nr, na]
        for a_, p in enumerate(ps):
            for b_, q in enumerate(qs):
                out[a_, b_, :] = Mr.dot(self._conv(p, q))
        return out

    def Y2(self, p: int, q: int, ypow: int = 0) -> Fraction:
        """(1/2) int y^ypow S_p S_q dy, exact."""
        c = self._conv(p, q)
        num = sum(int(c[a_]) * int(self.mom[a_ + ypow]) for a_ in range(len(c)))
        return Fraction(num, self.den ** 2 * self.mden)


# ---------------------------------------------------------------------------- assembly
def _groups(ijkl: np.ndarray):
    """Modes grouped by (family, j, k) in first-appearance order -> list of (key, [n...])."""
    order: List[Tuple[int, int, int]] = []
    members: Dict[Tuple[int, int, int], List[int]] = {}
    for n, (i, j, k, l) in enumerate(ijkl.tolist()):
        key = (i, j, k)
        if key not in members:
            members[key] = []
            order.append(key)
        members[key].append(n)
    return [(key, members[key]) for key in order]


def assemble_linear_exact(alpha, gamma, ijkl: np.ndarray, tables: WallTables | None = None):
    """Exact B, A1, A2 (object arrays of Fraction)  --  ODEModels.jl:31-36."""
    a, g = Fraction(alpha), Fraction(gamma)
    m = ijkl.shape[0]
    T = tables or WallTables(int(ijkl[:, 3].max()))
    comp = component_table(a, g, ijkl)
    ls = ijkl[:, 3].tolist()
    B = np.zeros((m, m), dtype=object); B[...] = Fraction(0)
    A1 = B.copy(); A2 = B.copy()
    for i in range(m):
        for j in range(m):
            b = a1 = a2 = Fraction(0)
            for c in range(3):
                ci, xi, zi, di = comp[i][c]
                cj, xj, zj, dj = comp[j][c]
                if ci != 0 and cj != 0 and xi == xj and zi == zj:
                    F = _fourier_pair(a, xi, xj, 0) * _fourier_pair(g, zi, zj, 0)
                    pi_, pj_ = T.pid(ls[i], di), T.pid(ls[j], dj)
                    y2 = T.Y2(pi_, pj_)
                    b += ci * cj * F * y2
                    # laplacian: p'' - ((a jx)^2 + (g kz)^2) p     (BasisFunctions.jl:316-320)
                    if dj + 2 <= 2:
                        ypp = T.Y2(pi_, T.pid(ls[j], dj + 2))
                    else:
                        ypp = _Y2_third(T, pi_, ls[j])
                    a2 += ci * cj * F * (ypp - ((a * xj) ** 2 + (g * zj) ** 2) * y2)
                # -<Psi_i, y d/dx Psi_j>
                if ci != 0 and cj != 0 and zi == zj:
                    Fx = _fourier_pair(a, xi, xj, 1)
                    if Fx != 0:
                        a1 -= (ci * cj * Fx * _fourier_pair(g, zi, zj, 0)
                               * T.Y2(T.pid(ls[i], di), T.pid(ls[j], dj), ypow=1))
            # -<Psi_i.u, Psi_j.v>   (vex, BasisFunctions.jl:431-436)
            ci, xi, zi, di = comp[i][0]
            cj, xj, zj, dj = comp[j][1]
            if ci != 0 and cj != 0 and xi == xj and zi == zj:
                a1 -= (ci * cj * _fourier_pair(a, xi, xj, 0) * _fourier_pair(g, zi, zj, 0)
                       * T.Y2(T.pid(ls[i], di), T.pid(ls[j], dj)))
            B[i, j], A1[i, j], A2[i, j] = b, a1, a2
    return B, A1, A2


def _Y2_third(T: WallTables, pi_: int, l: int) -> Fraction:
    """(1/2) int S_pi * S_l''' dy  -- needed when the laplacian hits an S' component."""
    # derivative of the d=2 polynomial, formed on the fly from integer coefficients
    c2 = T.coef[T.pid(l, 2)]
    c3 = [k * int(c2[k]) for k in range(1, len(c2))] + [0]
    cp = T.coef[pi_]
    n = T.deg + 1
    num = 0
    for a_ in range(n):
        if cp[a_] == 0:
            continue
        for b_ in range(n):
            if c3[b_] != 0:
                num += int(cp[a_]) * c3[b_] * int(T.mom[a_ + b_])
    return Fraction(num, T.den ** 2 * T.mden)


def assemble_N_exact(alpha, gamma, ijkl: np.ndarray, tables: WallTables | None = None,
                     rows=None, threshold=1e-15):
    """Exact N as COO in lexicographic (i,j,k) order: (ijk int64 [nnz,3] 1-based, [Fraction]).

    ``rows`` optionally restricts the output modes i (0-based iterable) -- used to sample
    large bases.  ODEModels.jl:39-51.
    """
    a, g = Fraction(alpha), Fraction(gamma)
    T = tables or WallTables(int(ijkl[:, 3].max()))
    comp = component_table(a, g, ijkl)
    ls = ijkl[:, 3].tolist()
    groups = _groups(ijkl)
    rowset = None if rows is None else set(int(r) for r in rows)
    den = Fraction(1, T.den ** 3 * T.mden)
    entries: Dict[Tuple[int, int, int], Fraction] = {}
    for _, Gi in groups:
        Gi_sel = Gi if rowset is None else [n for n in Gi if n in rowset]
        if not Gi_sel:
            continue
        ri = comp[Gi_sel[0]]
        for _, Gj in groups:
            rj = comp[Gj[0]]
            for _, Gk in groups:
                rk = comp[Gk[0]]
                block = None
                for c in range(3):
                    ci, xi, zi, di = ri[c]
                    ck, xk, zk, dk = rk[c]
                    if ci == 0 or ck == 0:
                        continue
                    for ax in range(3):
                        cj, xj, zj, dj = rj[ax]
                        if cj == 0:
                            continue
                        X = _fourier_triple(a, xi, xj, xk, 1 if ax == 0 else 0)
                        if X == 0:
                            continue
                        Zf = _fourier_triple(g, zi, zj, zk, 1 if ax == 2 else 0)
                        if Zf == 0:
                            continue
                        Kf = ci * cj * ck * X * Zf
                        ps = [T.pid(ls[n], di) for n in Gi_sel]
                        qs = [T.pid(ls[n], dj) for n in Gj]
                        rs = [T.pid(ls[n], dk + (1 if ax == 1 else 0)) for n in Gk]
                        yb = T.Y3_block(ps, qs, rs)
                        # keep numerators integral: Kf = kn/kd
                        contrib = yb * Kf.numerator
                        if block is None:
                            block = {}
                        block[Kf.denominator] = block.get(Kf.denominator, 0) + contrib
                if block is None:
                    continue
                total = None
                for kd, arr in block.items():
                    fr = np.frompyfunc(lambda v, kd=kd: -Fraction(v, kd) * den, 1, 1)(arr)
                    total = fr if total is None else total + fr
                for a_, ni in enumerate(Gi_sel):
                    for b_, nj in enumerate(Gj):
                        for c_, nk in enumerate(Gk):
                            v = total[a_, b_, c_]
                            if v != 0 and abs(v) > threshold:
                                entries[(ni + 1, nj + 1, nk + 1)] = v
    keys = sorted(entries)
    ijk = np.array(keys, dtype=np.int64).reshape(-1, 3)
    return ijk, [entries[k] for k in keys]
