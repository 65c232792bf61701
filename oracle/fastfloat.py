"""Float64 evaluation of the exact-rational tables, vectorised (test oracle only).

Same separable formula as ``oracle/exact.py`` (reference: ODEModels.jl:31-49 over
BasisFunctions.jl:87-99,133-212,278-285,328-356,444-463): every factor (component coefficients,
Fourier triple averages, wall-normal integrals Y3) is computed EXACTLY as a rational and rounded
once to Float64; only the <=9-term sum per entry is done in floating point.  Entry error is
therefore ~1e-16 * sum|terms|, independent of L (unlike the reference's monomial Float64 path, which
loses a digit per unit of L -- SURVEY.md exec. summary).  Used to produce N at m = 307 / 999 in
seconds for the parity tests; ``tests/test_oracle_exact.py`` checks it against the all-rational tier.

Zero test: an entry is kept iff |sum| > 1e-11 * sum|terms| (cancellation-aware) and |sum| > 1e-15
(the reference's absolute cut, ODEModels.jl:46).  On every configuration checked this reproduces the
exact-rational sparsity pattern bit for bit.
"""
from __future__ import annotations

from fractions import Fraction

import numpy as np

from .exact import WallTables, _fourier_pair, _fourier_triple, component_table

REL_ZERO = 1e-11
ABS_ZERO = 1e-15


class FloatTables:
    def __init__(self, alpha, gamma, ijkl: np.ndarray):
        a, g = Fraction(alpha), Fraction(gamma)
        self.m = m = ijkl.shape[0]
        self.J = J = int(np.abs(ijkl[:, 1]).max())
        self.K = K = int(np.abs(ijkl[:, 2]).max())
        self.L = L = int(ijkl[:, 3].max())
        comp = component_table(a, g, ijkl)
        self.coef = np.array([[float(c[0]) for c in row] for row in comp])            # [m,3]
        self.jx = np.array([[c[1] for c in row] for row in comp], dtype=np.int64)      # [m,3]
        self.kz = np.array([[c[2] for c in row] for row in comp], dtype=np.int64)
        d = np.array([[c[3] for c in row] for row in comp], dtype=np.int64)
        self.pid = 3 * ijkl[:, 3][:, None] + d                                           # [m,3]
        # Fourier triples  <E_a, E_b * d^dc E_c>  indexed [a+J, b+J, c+J, dc]
        self.X3 = self._triple_table(a, J)
        self.Z3 = self._triple_table(g, K)
        self.X2 = self._pair_table(a, J)
        self.Z2 = self._pair_table(g, K)
        T = WallTables(L)
        self.T = T
        npid = 3 * (L + 1)
        ids = list(range(npid))
        num = T.Y3_block(ids, ids, ids)
        den = T.den ** 3 * T.mden
        self.Y3 = np.array([float(Fraction(int(v), den)) for v in num.ravel()]).reshape(npid, npid, npid)

    @staticmethod
    def _triple_table(wn, J):
        n = 2 * J + 1
        out = np.zeros((n, n, n, 2))
        for ia in range(-J, J + 1):
            for ib in range(-J, J + 1):
                for ic in range(-J, J + 1):
                    for dc in (0, 1):
                        out[ia + J, ib + J, ic + J, dc] = float(_fourier_triple(wn, ia, ib, ic, dc))
        return out

    @staticmethod
    def _pair_table(wn, J):
        n = 2 * J + 1
        out = np.zeros((n, n, 2))
        for ia in range(-J, J + 1):
            for ib in range(-J, J + 1):
                for db in (0, 1):
                    out[ia + J, ib + J, db] = float(_fourier_pair(wn, ia, ib, db))
        return out


def assemble_N_float(alpha, gamma, ijkl: np.ndarray, rows=None, tables: FloatTables | None = None):
    """COO (ijk int64 [nnz,3] 1-based lexicographic, val float64 [nnz])."""
    t = tables or FloatTables(alpha, gamma, ijkl)
    m, J, K = t.m, t.J, t.K
    rows = range(m) if rows is None else rows
    out_ijk, out_val = [], []
    for i in rows:
        acc = np.zeros((m, m))
        mag = np.zeros((m, m))
        for c in range(3):
            ci = t.coef[i, c]
            if ci == 0:
                continue
            ck = t.coef[:, c]                                   # [m] over k
            for ax in range(3):
                cj = t.coef[:, ax]                              # [m] over j
                X = t.X3[t.jx[i, c] + J, t.jx[:, ax][:, None] + J, t.jx[:, c][None, :] + J, 1 if ax == 0 else 0]
                Z = t.Z3[t.kz[i, c] + K, t.kz[:, ax][:, None] + K, t.kz[:, c][None, :] + K, 1 if ax == 2 else 0]
                Y = t.Y3[t.pid[i, c], t.pid[:, ax][:, None], t.pid[:, c][None, :] + (1 if ax == 1 else 0)]
                term = (ci * cj[:, None] * ck[None, :]) * X * Z * Y
                acc -= term
                mag += np.abs(term)
        keep = (np.abs(acc) > REL_ZERO * mag) & (np.abs(acc) > ABS_ZERO)
        jj, kk = np.nonzero(keep)
        out_ijk.append(np.stack([np.full(jj.shape, i + 1), jj + 1, kk + 1], axis=1))
        out_val.append(acc[jj, kk])
    return np.concatenate(out_ijk).astype(np.int64), np.concatenate(out_val)


def assemble_linear_float(alpha, gamma, ijkl: np.ndarray, tables: FloatTables | None = None):
    """B, A1, A2 as float64, vectorised over (i,j), from exact-rational wall-normal tables
    (ODEModels.jl:31-36; same formulas as oracle/exact.py assemble_linear_exact)."""
    from .exact import _Y2_third
    t = tables or FloatTables(alpha, gamma, ijkl)
    a, g = float(Fraction(alpha)), float(Fraction(gamma))
    m, J, K, T = t.m, t.J, t.K, t.T
    npid = 3 * (t.L + 1)
    Y2 = np.array([[float(T.Y2(p, q)) for q in range(npid)] for p in range(npid)])
    Y2y = np.array([[float(T.Y2(p, q, ypow=1)) for q in range(npid)] for p in range(npid)])
    # <S_p, (S_q)''> for q = (l, d): d+2 <= 2 is a table entry, d = 1 needs the third derivative
    Y2pp = np.zeros((npid, npid))
    for p in range(npid):
        for l in range(t.L + 1):
            Y2pp[p, 3 * l + 0] = float(T.Y2(p, 3 * l + 2))
            Y2pp[p, 3 * l + 1] = float(_Y2_third(T, p, l))
    B = np.zeros((m, m)); A1 = np.zeros((m, m)); A2 = np.zeros((m, m))
    for c in range(3):
        ci, cj = t.coef[:, c][:, None], t.coef[:, c][None, :]
        xi, xj = t.jx[:, c][:, None] + J, t.jx[:, c][None, :] + J
        zi, zj = t.kz[:, c][:, None] + K, t.kz[:, c][None, :] + K
        pi, pj = t.pid[:, c][:, None], t.pid[:, c][None, :]
        F = ci * cj * t.X2[xi, xj, 0] * t.Z2[zi, zj, 0]
        y2 = Y2[pi, pj]
        k2 = (a * t.jx[:, c][None, :]) ** 2 + (g * t.kz[:, c][None, :]) ** 2
        B += F * y2
        A2 += F * (Y2pp[pi, pj] - k2 * y2)
        A1 -= ci * cj * t.X2[xi, xj, 1] * t.Z2[zi, zj, 0] * Y2y[pi, pj]
    ci, cj = t.coef[:, 0][:, None], t.coef[:, 1][None, :]
    A1 -= (ci * cj * t.X2[t.jx[:, 0][:, None] + J, t.jx[:, 1][None, :] + J, 0]
           * t.Z2[t.kz[:, 0][:, None] + K, t.kz[:, 1][None, :] + K, 0] * Y2[t.pid[:, 0][:, None], t.pid[:, 1][None, :]])
    return B, A1, A2
