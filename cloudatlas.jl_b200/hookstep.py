"""SearchParams and hookstepsolve -- mirror of src/Hookstep.jl over the C ABI.

Two forms, as SURVEY 8(b) lays out:
* ``hookstepsolve(f, Df, xguess, params)`` / ``hookstepsolve(f, xguess, params)`` -- the reference's
  generic closure form (Hookstep.jl:95,109).  The loop runs on the host because f and Df are arbitrary
  user callables; with ``model.f`` / ``model.Df`` closures every evaluation is a CUDA call.
* ``hookstepsolve(model, R, xguess, params)`` -- the device-resident search: one persistent kernel per
  solve (``ca_hookstep_solve_model``), also for batches of (R, guess) pairs.
Both return ``(x, success)``; ``success`` is True only on the ``norm(f) <= ftol`` exit (Hookstep.jl:146-149).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._lib import CloudAtlasError, _sig, c_f64p, c_i32p, check, vp
from .model import ModelOperators

CA_LOG_STEPS = 64
DEVICE_RESIDENT_MAX_M = 95   # 3 m^2 + 9 m doubles of shared memory (include/cloudatlas_b200.h)


@dataclass
class SearchParams:
    """Hookstep.jl:6-14 (the reference's delta field is spelled with a Greek letter)."""
    ftol: float = 1e-08
    xtol: float = 1e-08
    delta: float = 0.02
    Nnewton: int = 20
    Nhook: int = 4
    Nmusearch: int = 6
    verbosity: int = 0


class _CParams(C.Structure):
    _fields_ = [("ftol", C.c_double), ("xtol", C.c_double), ("delta", C.c_double),
                ("Nnewton", C.c_int32), ("Nhook", C.c_int32), ("Nmusearch", C.c_int32), ("verbosity", C.c_int32)]


class _CLog(C.Structure):
    _fields_ = [("exit_reason", C.c_int32), ("newton_steps", C.c_int32), ("f_evals", C.c_int32),
                ("df_evals", C.c_int32), ("final_residual", C.c_double), ("final_delta", C.c_double),
                ("device_seconds", C.c_double), ("residual", C.c_double * CA_LOG_STEPS),
                ("delta", C.c_double * CA_LOG_STEPS), ("hooksteps", C.c_int32 * CA_LOG_STEPS)]


EXIT_REASONS = {1: "converged", 2: "residual_increasing", 3: "xtol_newton", 4: "xtol_hookstep",
                5: "max_newton", 6: "singular"}

_solve = _sig("ca_hookstep_solve_model", vp, c_f64p, c_f64p, C.c_int64, C.POINTER(_CParams), c_f64p, c_i32p,
              C.POINTER(_CLog))


def _log_dict(lg: _CLog):
    n = min(lg.newton_steps + 1, CA_LOG_STEPS)
    return {"exit_reason": EXIT_REASONS.get(lg.exit_reason, str(lg.exit_reason)), "newton_steps": lg.newton_steps,
            "f_evals": lg.f_evals, "df_evals": lg.df_evals, "final_residual": lg.final_residual,
            "final_delta": lg.final_delta, "device_seconds": lg.device_seconds,
            "residual": list(lg.residual[:n]), "delta": list(lg.delta[:n]), "hooksteps": list(lg.hooksteps[:n])}


def hookstepsolve_model(model: ModelOperators, R, xguess, params: SearchParams | None = None, return_log=False):
    """Device-resident search(es).  xguess 1-D (one search, R scalar) or nsolve x m with R scalar or nsolve."""
    params = params or SearchParams()
    X = np.asarray(xguess, dtype=np.float64)
    single = X.ndim == 1
    # any m goes through the library: up to DEVICE_RESIDENT_MAX_M modes one persistent CTA per search; beyond that the
    # global-memory solver (csrc/hookstep_large.cu: blocked LU, Df'Df and all vectors on the device, searches one after
    # the other)
    X = np.asfortranarray(X.reshape(1, -1) if single else X)
    if X.shape[1] != model.m:
        raise CloudAtlasError(1, f"xguess has {X.shape[1]} columns, expected {model.m}")
    n = X.shape[0]
    Rv = np.ascontiguousarray(np.broadcast_to(np.asarray(R, dtype=np.float64), (n,)))
    out = np.empty((n, model.m), order="F")
    succ = np.zeros(n, dtype=np.int32)
    logs = (_CLog * n)()
    cp = _CParams(params.ftol, params.xtol, params.delta, params.Nnewton, params.Nhook, params.Nmusearch,
                  params.verbosity)
    check(_solve(model._h, Rv.ctypes.data_as(c_f64p), X.ctypes.data_as(c_f64p), n, C.byref(cp),
                 out.ctypes.data_as(c_f64p), succ.ctypes.data_as(c_i32p), logs))
    if single:
        res = (out[0].copy(), bool(succ[0]))
        return res + (_log_dict(logs[0]),) if return_log else res
    res = (out, succ.astype(bool))
    return res + ([_log_dict(l) for l in logs],) if return_log else res


def Df_finitediff(f, x, eps=1e-06):
    """Hookstep.jl:77-93."""
    fx = np.asarray(f(x), dtype=np.float64)
    Df = np.zeros((len(fx), len(x)))
    for j in range(len(x)):
        d = np.zeros(len(x))
        d[j] = eps
        Df[:, j] = (np.asarray(f(x + d), dtype=np.float64) - fx) / eps
    return Df


def _hookstep(fx, Dfx, delta, dx_newt, dtol=1e-04, Nmusearch=10):
    """Hookstep.jl:21-74: the step of radius delta minimising 1/2 |fx + Df dx|^2 (Newton iteration on mu)."""
    norm_dx = np.linalg.norm(dx_newt)
    if norm_dx <= delta:
        return dx_newt
    mu, dx = 1e-06, dx_newt.copy()
    H = Dfx.T @ Dfx
    eye = np.eye(H.shape[0])
    for _ in range(Nmusearch):
        phi = norm_dx - delta
        phiprime = -(1.0 / norm_dx) * np.dot(dx, np.linalg.solve(H + mu * eye, dx))
        mu = mu - (norm_dx / delta) * (phi / phiprime)
        dx = -np.linalg.solve(H + mu * eye, Dfx.T @ fx)
        norm_dx = np.linalg.norm(dx)
        if abs(norm_dx - delta) / delta <= dtol:
            break
    return dx


def _hookstepsolve_closures(f, Df, xguess, params: SearchParams):
    """Hookstep.jl:109-352, host loop over user callables."""
    as_np = lambda v: np.asarray(v.cpu() if hasattr(v, "cpu") else v, dtype=np.float64)
    x = np.asarray(xguess, dtype=np.float64).copy()
    delta = params.delta
    for _ in range(params.Nnewton):
        fx = as_np(f(x))
        rx = 0.5 * float(fx @ fx)
        if np.sqrt(2 * rx) <= params.ftol:
            return x, True
        Dfx = as_np(Df(x))
        dx = -np.linalg.solve(Dfx, fx)
        norm_newt = np.linalg.norm(dx)
        if float(fx @ (Dfx @ dx)) >= 0:
            return x, False
        if norm_newt < params.xtol:
            return x + dx, False
        dx_newt, x_hook = dx, x + dx
        for _h in range(params.Nhook):
            if norm_newt > delta:
                dx = _hookstep(fx, Dfx, delta, dx_newt, Nmusearch=params.Nmusearch)
            Dfdx = Dfx @ dx
            if np.linalg.norm(dx) < params.xtol:
                return x + dx, False
            within = norm_newt <= delta
            x_hook = x + dx
            fh = as_np(f(x_hook))
            r_hook = 0.5 * float(fh @ fh)
            dr_hook = rx - r_hook
            dr_linear = -float(fx @ Dfdx)
            dr_quadratic = rx - 0.5 * float((fx + Dfdx) @ (fx + Dfdx))
            if dr_hook > dr_linear:
                if within:
                    break
                delta = 3 * delta / 2
            elif dr_hook < 0.01 * dr_quadratic:
                delta = delta / 2
            elif dr_hook < 0.10 * dr_quadratic:
                delta = delta / 2
                break
            elif 0.8 * dr_quadratic < dr_hook < 1.2 * dr_quadratic:
                slope = -dr_linear / delta
                curv = (r_hook - rx - slope * delta) / delta ** 2
                dnew = -slope / (2 * curv)
                if not within:
                    delta = min(max(dnew, 2 * delta / 3), 4 * delta / 3)
                break
            else:
                break
        x = x_hook
    return x, False


def hookstepsolve(*args, **kw):
    """hookstepsolve(f, Df, xguess[, params]) | hookstepsolve(f, xguess[, params]) |
    hookstepsolve(model, R, xguess[, params])  -> (x, success)."""
    if isinstance(args[0], ModelOperators):
        return hookstepsolve_model(*args, **kw)
    args = list(args)
    params = args.pop() if isinstance(args[-1], SearchParams) else kw.get("params") or SearchParams()
    if len(args) == 2:  # Hookstep.jl:95
        f, xguess = args
        return _hookstepsolve_closures(f, lambda x: Df_finitediff(f, x), xguess, params)
    f, Df, xguess = args
    return _hookstepsolve_closures(f, Df, xguess, params)
