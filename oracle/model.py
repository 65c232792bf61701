"""Galerkin assembly of the ODE model and its closures (test oracle only).

Follows /root/reference/src/ODEModels.jl:22-61:
  B_ij  = <Psi_i, Psi_j>                                   (:32)
  A1_ij = -<Psi_i, y d/dx Psi_j> - <Psi_i, [v_j,0,0]>      (:33-36)
  A2_ij = <Psi_i, laplacian Psi_j>                         (:35)
  N_ijk = -<Psi_i, (Psi_j . grad) Psi_k>, kept iff |val| > 1e-15   (:39-49)
  f(x,R)  = B \\ (A1 x + (1/R) A2 x + N(x))                 (:57)
  Df(x,R) = B \\ (A1 + (1/R) A2 + derivative(N,x))          (:58)
"""
from __future__ import annotations

from fractions import Fraction

import numpy as np

from .algebra import (Poly, bf_dotgrad, bf_innerproduct, bf_innerproduct_terms, bf_laplacian,
                      bf_polymul, bf_vex, bf_xderivative)
from .basis import basisIndices, basisSet
from .bilinear import SparseBilinear

N_THRESHOLD = 1e-15  # ODEModels.jl:46 (absolute)


def assemble_linear(Psi):
    """B, A1, A2 as m x m nested lists in the number type of Psi."""
    m = len(Psi)
    T = type(Psi[0][0].coeff)
    y = Poly([T(0), T(1)])
    B = [[bf_innerproduct(Psi[i], Psi[j]) for j in range(m)] for i in range(m)]
    yPx = [bf_polymul(y, bf_xderivative(Psi[j])) for j in range(m)]
    vex = [bf_vex(Psi[j]) for j in range(m)]
    lap = [bf_laplacian(Psi[j]) for j in range(m)]
    A1a = [[-bf_innerproduct(Psi[i], yPx[j]) for j in range(m)] for i in range(m)]
    A1b = [[-bf_innerproduct(Psi[i], vex[j]) for j in range(m)] for i in range(m)]
    A2 = [[bf_innerproduct(Psi[i], lap[j]) for j in range(m)] for i in range(m)]
    A1 = [[A1a[i][j] + A1b[i][j] for j in range(m)] for i in range(m)]
    return B, A1, A2


def assemble_N_dense(Psi, threshold=N_THRESHOLD):
    """Dense m x m x m object/float array with the reference's absolute threshold."""
    m = len(Psi)
    T = type(Psi[0][0].coeff)
    dt = np.float64 if T is float else object
    Nd = np.zeros((m, m, m), dtype=dt)
    if dt is object:
        Nd[...] = Fraction(0)
    for j in range(m):
        for k in range(m):
            g = bf_dotgrad(Psi[j], Psi[k])
            for i in range(m):
                val = -bf_innerproduct_terms(Psi[i], g)
                if abs(val) > threshold:
                    Nd[i, j, k] = val
    return Nd


class ODEModel:
    """Restatement of the reference struct (fields alpha, gamma, H, ijkl, Psi, B, A1, A2, N)."""

    def __init__(self, alpha, gamma, J, K, L, H, normalize=False, ijkl=None):
        self.alpha, self.gamma, self.H = alpha, gamma, list(H) if H is not None else None
        self.ijkl = basisIndices(J, K, L, self.H) if ijkl is None else np.asarray(ijkl, np.int64)
        self.Psi = basisSet(alpha, gamma, self.ijkl, normalize=normalize)
        B, A1, A2 = assemble_linear(self.Psi)
        self.exact = isinstance(self.Psi[0][0].coeff, Fraction)
        if self.exact:
            self.B, self.A1, self.A2 = B, A1, A2
        else:
            self.B, self.A1, self.A2 = (np.array(M, dtype=np.float64) for M in (B, A1, A2))
        self.N = SparseBilinear.from_dense(assemble_N_dense(self.Psi))

    def __len__(self):
        return len(self.Psi)

    # closures (float models only)
    def f(self, x, R):
        x = np.asarray(x, dtype=np.float64)
        rhs = self.A1 @ x + (1.0 / R) * (self.A2 @ x) + self.N(x)
        return np.linalg.solve(self.B, rhs)

    def Df(self, x, R):
        x = np.asarray(x, dtype=np.float64)
        return np.linalg.solve(self.B, self.A1 + (1.0 / R) * self.A2 + self.N.derivative(x))


def cnab2(x0, model: ODEModel, R, dt, Nsteps, nsave):
    """Crank-Nicolson / Adams-Bashforth-2 time stepping  --  ODEModels.jl:147-181, quirks included:
    t = (0:Nsave-1)*dt (not nsave*dt), and the row counter starts at 1, so X[0] = x0 is overwritten by the
    first saved step (:164-177)."""
    A = model.A1 + (1.0 / R) * model.A2
    LHS = model.B - (dt / 2) * A
    RHS = model.B + (dt / 2) * A
    xn = np.asarray(x0, dtype=np.float64).copy()
    Nxn1 = model.N(xn)
    Nxn = model.N(xn)
    Nsave = Nsteps // nsave
    X = np.zeros((Nsave, len(xn)))
    t = np.arange(Nsave) * dt
    if Nsave > 0:
        X[0] = xn
    k = 0
    for n in range(1, Nsteps + 1):
        Nxn1 = Nxn
        Nxn = model.N(xn)
        xn = np.linalg.solve(LHS, RHS @ xn + (3 * dt / 2) * Nxn - (dt / 2) * Nxn1)
        if n % nsave == 0:
            X[k] = xn
            k += 1
    return t, X


def shear_vector(Psi):
    """Psi_shear of shear_function  --  ODEModels.jl:65-78 (shear(x) = 1 + dot(x, Psi_shear))."""
    from .algebra import bc_yderivative
    out = np.zeros(len(Psi))
    for i, f in enumerate(Psi):
        u = f[0]
        if u.ejx.waveindex == 0 and u.ekz.waveindex == 0:
            d = bc_yderivative(u)
            out[i] = u.coeff * (d.p(1.0) + d.p(-1.0)) / 2
    return out


def dissipation_matrix(Psi):
    """D_ij = <curl Psi_i, curl Psi_j>  --  ODEModels.jl:87-137 (dissipation(x) = 1 + x' D x)."""
    from .algebra import bc_innerproduct as ip, bc_xderivative as dx, bc_yderivative as dy, bc_zderivative as dz
    m = len(Psi)
    D = np.zeros((m, m))
    for j in range(m):
        g = Psi[j]
        for i in range(j + 1):
            f = Psi[i]
            vx = (ip(dy(f[2]), dy(g[2])) - ip(dy(f[2]), dz(g[1])) - ip(dz(f[1]), dy(g[2])) + ip(dz(f[1]), dz(g[1])))
            vy = (ip(dz(f[0]), dz(g[0])) - ip(dz(f[0]), dx(g[2])) - ip(dx(f[2]), dz(g[0])) + ip(dx(f[2]), dx(g[2])))
            vz = (ip(dx(f[1]), dx(g[1])) - ip(dx(f[1]), dy(g[0])) - ip(dy(f[0]), dx(g[1])) + ip(dy(f[0]), dy(g[0])))
            D[i, j] = D[j, i] = vx + vy + vz
    return D
