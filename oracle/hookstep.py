"""Newton-hookstep solver (test oracle only).

Follows /root/reference/src/Hookstep.jl: SearchParams :6-14, hookstep :21-74,
Df_finitediff :77-93, hookstepsolve :95,:109-352.  The decision tree, constants
(mu0 = 1e-6, dtol = 1e-4, factors 3/2, 1/2, 2/3, 4/3, 0.01/0.10/0.8/1.2) and exit flags are
the reference's: ``success`` is True only on the ``norm(f) <= ftol`` exit.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class SearchParams:
    ftol: float = 1e-08
    xtol: float = 1e-08
    delta: float = 0.02
    Nnewton: int = 20
    Nhook: int = 4
    Nmusearch: int = 6
    verbosity: int = 0


def hookstep(fx, Dfx, delta, dx_newt, dtol=1e-04, Nmusearch=10):
    """Hookstep.jl:21-74."""
    norm_dx = np.linalg.norm(dx_newt)
    if norm_dx <= delta:
        return dx_newt
    mu = 1e-06
    dx = dx_newt.copy()
    H = Dfx.T @ Dfx
    I = np.eye(H.shape[0])
    for _ in range(Nmusearch):
        phi = norm_dx - delta
        phiprime = -(1.0 / norm_dx) * np.dot(dx, np.linalg.solve(H + mu * I, dx))
        mu = mu - (norm_dx / delta) * (phi / phiprime)
        dx = -np.linalg.solve(H + mu * I, Dfx.T @ fx)
        norm_dx = np.linalg.norm(dx)
        if abs(norm_dx - delta) / delta <= dtol:
            return dx
    return dx


def Df_finitediff(f, x, eps=1e-06):
    """Hookstep.jl:77-93."""
    fx = np.asarray(f(x), dtype=np.float64)
    M, N = len(fx), len(x)
    Df = np.zeros((M, N))
    for j in range(N):
        d = np.zeros(N)
        d[j] = eps
        Df[:, j] = (np.asarray(f(x + d), dtype=np.float64) - fx) / eps
    return Df


def hookstepsolve(f, Df, xguess, params: SearchParams | None = None, log=None):
    """Hookstep.jl:109-352.  Returns (x, success).  ``log`` (a list) receives per-Newton-step
    tuples (norm_f, delta, n_hooksteps) for comparison with the device solver's log."""
    if params is None:
        params = SearchParams()
    if Df is None:
        Df = lambda x: Df_finitediff(f, x)
    norm2 = lambda v: float(np.dot(v, v))
    x = np.asarray(xguess, dtype=np.float64).copy()
    ftol, xtol, Nhook, Nmusearch = params.ftol, params.xtol, params.Nhook, params.Nmusearch
    delta = params.delta
    for n in range(params.Nnewton):
        fx = np.asarray(f(x), dtype=np.float64)
        rx = 0.5 * norm2(fx)
        if np.sqrt(2 * rx) <= ftol:
            if log is not None:
                log.append((np.sqrt(2 * rx), delta, 0))
            return x, True
        Dfx = np.asarray(Df(x), dtype=np.float64)
        dx = -np.linalg.solve(Dfx, fx)
        norm_dx = np.linalg.norm(dx)
        if np.dot(fx, Dfx @ dx) >= 0:
            return x, False
        if np.linalg.norm(dx) < xtol:
            return x + dx, False
        dx_newt, norm_dx_newt = dx, norm_dx
        x_hook = x + dx
        nh = 0
        for h in range(Nhook):
            nh += 1
            if norm_dx_newt > delta:
                dx = hookstep(fx, Dfx, delta, dx_newt, Nmusearch=Nmusearch)
            Dfdx = Dfx @ dx
            if np.linalg.norm(dx) < xtol:
                return x + dx, False
            within = norm_dx_newt <= delta
            x_hook = x + dx
            r_hook = 0.5 * norm2(np.asarray(f(x + dx), dtype=np.float64))
            r_linear = rx + np.dot(fx, Dfdx)
            r_quadratic = 0.5 * norm2(fx + Dfdx)
            dr_hook = rx - r_hook
            dr_linear = rx - r_linear
            dr_quadratic = rx - r_quadratic
            if dr_hook > dr_linear:
                if within:
                    break
                delta = 3 * delta / 2
                continue
            elif dr_hook < 0.01 * dr_quadratic:
                delta = delta / 2
                continue
            elif dr_hook < 0.10 * dr_quadratic:
                delta = delta / 2
                break
            elif 0.8 * dr_quadratic < dr_hook < 1.2 * dr_quadratic:
                mm = -dr_linear / delta
                c = (r_hook - rx - mm * delta) / delta ** 2
                dnew = -mm / (2 * c)
                if within:
                    pass
                elif dnew < 2 * delta / 3:
                    delta = 2 * delta / 3
                elif dnew > 4 * delta / 3:
                    delta = 4 * delta / 3
                else:
                    delta = dnew
                break
            else:
                break
        if log is not None:
            log.append((np.sqrt(2 * rx), delta, nh))
        x = x_hook
    return x, False
