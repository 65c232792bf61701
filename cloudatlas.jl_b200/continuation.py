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

import numpy as np

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


def newton_krylov(model, R, x0, ftol=1e-10, max_newton=30, gmres_rtol=1e-8, restart=60, max_restarts=5, verbose=False):
    """Matrix-free Newton-GMRES for f(x,R) = 0 (BASELINE.json configs[4]: Jacobian actions for Newton-Krylov at
    m ~ 3000, where the dense m x m Jacobian and its LU are what dominate the reference's hookstep).  Every Krylov
    vector costs ONE matrix-free Jacobian action  Df(x,R) v = (L1 + L2/R) v + Q(x,v) + Q(v,x)  on the device
    (``ca_model_jvp_batched_dev``: the HBM-bound streaming kernel, 0.38 ms at m = 2962); the Arnoldi process runs on
    device tensors.  Backtracking line search on |f|.  Returns (x, converged, info)."""
    import torch
    dev = torch.device("cuda")
    x = torch.as_tensor(np.asarray(x0, dtype=np.float64), device=dev).reshape(1, -1)
    m = x.shape[1]

    def f(xx):
        return model.f(xx, R).reshape(-1)

    def jv(xx, v):
        return model.Df_action(xx, v.reshape(1, -1), R).reshape(-1)

    info = {"newton_steps": 0, "jvps": 0, "residuals": []}
    fx = f(x)
    for _ in range(max_newton):
        nf = float(torch.linalg.norm(fx))
        info["residuals"].append(nf)
        if nf <= ftol:
            return x.reshape(-1).cpu().numpy(), True, info
        # ---- restarted GMRES for Df dx = -fx
        dx = torch.zeros(m, dtype=torch.float64, device=dev)
        b = -fx
        for _r in range(max_restarts):
            r0 = b - (jv(x, dx) if _r > 0 else 0)
            beta = float(torch.linalg.norm(r0))
            if beta <= gmres_rtol * nf:
                break
            V = torch.zeros((restart + 1, m), dtype=torch.float64, device=dev)
            Hm = np.zeros((restart + 1, restart))
            V[0] = r0 / beta
            k_used = 0
            for k in range(restart):
                w = jv(x, V[k])
                info["jvps"] += 1
                h = V[:k + 1] @ w                       # classical Gram-Schmidt, twice (stable enough, two GEMVs)
                w = w - h @ V[:k + 1]
                h2 = V[:k + 1] @ w
                w = w - h2 @ V[:k + 1]
                Hm[:k + 1, k] = (h + h2).cpu().numpy()
                hn = float(torch.linalg.norm(w))
                Hm[k + 1, k] = hn
                k_used = k + 1
                e1 = np.zeros(k + 2)
                e1[0] = beta
                y, res, *_ = np.linalg.lstsq(Hm[:k + 2, :k + 1], e1, rcond=None)
                rnorm = np.linalg.norm(Hm[:k + 2, :k + 1] @ y - e1)
                if hn == 0.0 or rnorm <= gmres_rtol * nf:
                    break
                V[k + 1] = w / hn
            dx = dx + torch.as_tensor(y, device=dev) @ V[:k_used]
            if rnorm <= gmres_rtol * nf:
                break
        # ---- backtracking
        lam = 1.0
        while True:
            xt = x + lam * dx.reshape(1, -1)
            ft = f(xt)
            if float(torch.linalg.norm(ft)) < (1 - 1e-4 * lam) * nf or lam < 1e-4:
                break
            lam *= 0.5
        x, fx = xt, ft
        info["newton_steps"] += 1
        if verbose:
            print(f"newton {info['newton_steps']}: |f| = {float(torch.linalg.norm(fx)):.3e}, lambda = {lam}, jvps = {info['jvps']}")
    nf = float(torch.linalg.norm(fx))
    info["residuals"].append(nf)
    return x.reshape(-1).cpu().numpy(), nf <= ftol, info
