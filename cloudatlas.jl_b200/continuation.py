"""Continuation of equilibria in the Reynolds number (BASELINE.json configs[1]).

The reference does this with the third-party BifurcationKit.jl, calling ``model.f(x, p[1])`` thousands of times
(notebooks/continue-eqb.ipynb:654-668, PALC, both sides).  Here:
* ``natural_continuation`` -- a sweep over given Re values, each search started from the previous solution, every
  search one device-resident hookstep kernel (``ca_hookstep_solve_model``);
* ``pseudo_arclength_continuation`` -- a predictor / Newton-corrector scheme on the extended system
  [f(x,R); t.(z - z_pred)] = 0 that follows the branch around folds; f, Df come from the device,
  df/dR = -(L2 x)/R^2 is obtained from two evaluations of f (f is affine in 1/R).
Both record the wall shear rate (the notebooks' ordinate) when the model has observables.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import CloudAtlasError, _sig, c_f64p, check, vp
from .hookstep import SearchParams, hookstepsolve_model


def natural_continuation(model, x0, R_values, params: SearchParams | None = None):
    """Returns (X [n x m], ok [n], shear [n] or None).  Stops at the first failed search."""
    params = params or SearchParams()
    x = np.asarray(x0, dtype=np.float64)
    X, ok = [], []
    for R in R_values:
        xn, good = hookstepsolve_model(model, float(R), x, params)
        X.append(xn)
        ok.append(good)
        if not good:
            break
        x = xn
    X = np.array(X)
    shear = model.shear(X) if hasattr(model, "shear") and len(X) else None
    return X, np.array(ok), shear


def _dfdR(model, x, R):
    # f(x,R) = c(x) + (1/R) L2 x   =>   df/dR = -(L2 x) / R^2,   L2 x = (f(x,R) - f(x,2R)) / (1/R - 1/(2R))
    L2x = (model.f(x, R) - model.f(x, 2 * R)) / (0.5 / R)
    return -L2x / R ** 2


def pseudo_arclength_continuation(model, x0, R0, ds=1.0, nsteps=50, R_min=-np.inf, R_max=np.inf, direction=+1,
                                  tol=1e-10, max_newton=12, ds_min=1e-4, ds_max=20.0, scale_R=None):
    """Follows the branch through (x0, R0).  ``direction`` +1 / -1: initial sign of dR/ds.  The arclength metric
    weights R by 1/scale_R (default |R0|) so that x and R are commensurate.  Returns a dict with R [n], X [n x m],
    residual [n], shear [n] or None."""
    x = np.asarray(x0, dtype=np.float64).copy()
    R = float(R0)
    sR = float(scale_R or abs(R0))
    m = len(x)

    def tangent(x, R, prev=None):
        J = np.asarray(model.Df(x, R))
        fR = _dfdR(model, x, R) * sR                     # derivative w.r.t. r = R / sR
        dx = np.linalg.solve(J, -fR)                      # dx/dr along the branch
        t = np.append(dx, 1.0)
        t /= np.linalg.norm(t)
        if prev is not None and t @ prev < 0:
            t = -t
        return t

    t = tangent(x, R)
    if t[-1] * direction < 0:
        t = -t
    out = {"R": [R], "X": [x.copy()], "residual": [float(np.linalg.norm(model.f(x, R)))]}
    while len(out["R"]) <= nsteps:
        zp = np.append(x, R / sR) + ds * t                # predictor
        z = zp.copy()
        good = False
        for _ in range(max_newton):
            xc, Rc = z[:m], z[m] * sR
            F = np.append(np.asarray(model.f(xc, Rc)), t @ (z - zp))
            if np.linalg.norm(F) < tol:
                good = True
                break
            J = np.zeros((m + 1, m + 1))
            J[:m, :m] = np.asarray(model.Df(xc, Rc))
            J[:m, m] = _dfdR(model, xc, Rc) * sR
            J[m, :] = t
            z = z - np.linalg.solve(J, F)
        else:
            xc, Rc = z[:m], z[m] * sR
            good = np.linalg.norm(np.append(np.asarray(model.f(xc, Rc)), t @ (z - zp))) < tol
        if not good:
            ds *= 0.5
            if abs(ds) < ds_min:
                break
            continue
        x, R = z[:m].copy(), float(z[m] * sR)
        if not (R_min <= R <= R_max):
            break
        t = tangent(x, R, prev=t)
        out["R"].append(R)
        out["X"].append(x.copy())
        out["residual"].append(float(np.linalg.norm(model.f(x, R))))
        ds = min(ds * 1.3, ds_max)
    out["R"] = np.array(out["R"])
    out["X"] = np.array(out["X"])
    out["residual"] = np.array(out["residual"])
    out["shear"] = model.shear(out["X"]) if hasattr(model, "shear") else None
    return out


class _KrylovParams(C.Structure):
    _fields_ = [("ftol", C.c_double), ("gmres_rtol", C.c_double), ("max_newton", C.c_int32), ("restart", C.c_int32),
                ("max_restarts", C.c_int32), ("verbosity", C.c_int32), ("precondition", C.c_int32)]


class _KrylovInfo(C.Structure):
    _fields_ = [("converged", C.c_int32), ("newton_steps", C.c_int32), ("jvps", C.c_int32), ("f_evals", C.c_int32),
                ("world", C.c_int32), ("final_residual", C.c_double), ("device_seconds", C.c_double),
                ("residual", C.c_double * 64)]


_newton_krylov = _sig("ca_newton_krylov", C.POINTER(vp), C.c_int32, vp, C.c_double, c_f64p, C.POINTER(_KrylovParams), c_f64p,
                      C.POINTER(C.c_int32), C.POINTER(_KrylovInfo))


def newton_krylov(model, R, x0, ftol=1e-10, max_newton=30, gmres_rtol=1e-8, restart=60, max_restarts=5, verbose=False,
                  comm=None, precondition=True):
    """Matrix-free Newton-GMRES for f(x,R) = 0 inside the library (``ca_newton_krylov``, csrc/newton_krylov.cu;
    BASELINE.json configs[4]: Jacobian actions for Newton-Krylov at m ~ 3000, where the dense m x m Jacobian and its LU
    are what dominate the reference's hookstep).  Every Krylov vector costs ONE Jacobian action
    Df(x,R) v = (L1 + L2/R) v + Q(x,v) + Q(v,x)  by the HBM-bound streaming kernel; Arnoldi vectors, orthogonalisation
    and line search stay on the device.  ``model``: one model, or a list of replicas of one model on the local devices
    of ``comm`` (a ``LibraryComm``): the Jacobian action is then sharded by row slab over the communicator's ranks with
    one all-gather of m doubles per Krylov step.  ``precondition``: GMRES on M^-1 Df with M = L1 + L2/R factored once
    per solve on the device (needed beyond ~100 modes).  Returns (x, converged, info)."""
    models = list(model) if isinstance(model, (list, tuple)) else [model]
    m = models[0].m
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    if x0.shape != (m,):
        raise CloudAtlasError(1, f"x0 has shape {x0.shape}, expected ({m},)")
    handles = (vp * len(models))(*[mm._h for mm in models])
    p = _KrylovParams(ftol, gmres_rtol, max_newton, restart, max_restarts, int(bool(verbose)), int(bool(precondition)))
    info = _KrylovInfo()
    out = np.empty(m)
    ok = C.c_int32(0)
    check(_newton_krylov(handles, len(models), None if comm is None else comm._h, float(R), x0.ctypes.data_as(c_f64p),
                         C.byref(p), out.ctypes.data_as(c_f64p), C.byref(ok), C.byref(info)))
    n = min(info.newton_steps + 1, 64)
    return out, bool(ok.value), {"newton_steps": info.newton_steps, "jvps": info.jvps, "f_evals": info.f_evals,
                                 "world": info.world, "final_residual": info.final_residual,
                                 "device_seconds": info.device_seconds, "residuals": list(info.residual[:n])}
