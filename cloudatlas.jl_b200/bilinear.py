"""SparseBilinear -- mirror of src/SparseBilinear.jl over the C ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import CloudAtlasError, _sig, c_f64p, c_i64p, check, lib, vp

_create = _sig("ca_bilinear_create", c_i64p, c_f64p, C.c_int64, C.c_int64, C.POINTER(vp))
_destroy = _sig("ca_bilinear_destroy", vp)
_info = _sig("ca_bilinear_info", vp, c_i64p, c_i64p, c_i64p, c_i64p, c_i64p)
_eval = _sig("ca_bilinear_eval", vp, c_f64p, c_f64p, c_f64p)
_eval_b = _sig("ca_bilinear_eval_batched", vp, c_f64p, C.c_int64, c_f64p)
_eval_b_dev = _sig("ca_bilinear_eval_batched_dev", vp, vp, C.c_int64, C.c_int64, vp, C.c_int64, vp)
_eval2_b_dev = _sig("ca_bilinear_eval2_batched_dev", vp, vp, vp, C.c_int64, C.c_int64, vp, C.c_int64, vp)
_jvp_b_dev = _sig("ca_bilinear_jvp_batched_dev", vp, vp, vp, C.c_int64, C.c_int64, vp, C.c_int64, vp)
_jac = _sig("ca_bilinear_jacobian", vp, c_f64p, c_f64p)
_jac_dev = _sig("ca_bilinear_jacobian_dev", vp, vp, vp, vp)
_jac_b = _sig("ca_bilinear_jacobian_batched", vp, c_f64p, C.c_int64, c_f64p)
_jac_b_dev = _sig("ca_bilinear_jacobian_batched_dev", vp, vp, C.c_int64, C.c_int64, vp, C.c_int64, vp)


def _f64(a):
    return a.ctypes.data_as(c_f64p)


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _stream():
    import torch
    return vp(torch.cuda.current_stream().cuda_stream)


def _check_device_tensor(X):
    """The library dereferences raw pointers: only float64 tensors on the current CUDA device may reach it."""
    import torch
    if not X.is_cuda:
        raise CloudAtlasError(1, "expected a CUDA tensor (host data goes through the numpy entry points)")
    if X.dtype != torch.float64:
        raise CloudAtlasError(1, f"expected a float64 tensor, got {X.dtype}")
    if X.device.index != torch.cuda.current_device():
        raise CloudAtlasError(1, f"tensor is on cuda:{X.device.index}, the current device is "
                                 f"cuda:{torch.cuda.current_device()} (handles live on the device that was current "
                                 "when they were created)")


def _colmajor_ensemble(X, m=None):
    """torch CUDA ensemble -> (tensor, nstates, ld).  Accepts an (nstates, m) float64 tensor on the current device
    whose memory is column-major (stride (1, ld)), which is what Julia hands over; ``m``: required column count."""
    _check_device_tensor(X)
    if m is not None and (X.dim() != 2 or X.shape[1] != m):
        raise CloudAtlasError(1, f"ensemble has shape {tuple(X.shape)}, expected nstates x {m}")
    if X.dim() == 2 and X.shape[0] == 1:  # a single state: any stride along the (length-1) state axis
        return X, 1, max(X.stride(1), 1)
    if X.dim() != 2 or X.stride(0) != 1:
        raise CloudAtlasError(1, "ensemble must be nstates x m with unit stride along states "
                                 "(column-major, as a Julia Matrix); use ensemble_empty() or X.t().contiguous().t()")
    return X, X.shape[0], max(X.stride(1), X.shape[0])


def jacobian_ensemble(X, m, call):
    """Shared device path of the batched dense Jacobians: (nstates, m, m) CUDA tensor with strides (1, n, n*m)."""
    import torch
    X, n, ld = _colmajor_ensemble(X)
    if X.shape[1] != m:
        raise CloudAtlasError(1, f"ensemble has {X.shape[1]} columns, expected {m}")
    D = torch.empty((m, m, n), dtype=torch.float64, device=X.device).permute(2, 1, 0)
    check(call(vp(X.data_ptr()), n, ld, vp(D.data_ptr())))
    return D


def ensemble_empty(nstates, m, device="cuda"):
    """An nstates x m float64 CUDA tensor in column-major (Julia) layout."""
    import torch
    return torch.empty((m, nstates), dtype=torch.float64, device=device).t()


class SparseBilinear:
    """SparseBilinear{Float64}: ``ijk`` (nnz x 3, 1-based), ``val`` (nnz), ``m``.

    Reference: SparseBilinear.jl:3-13 (struct + checks), :24-49 (dense constructor),
    :56-70 (evaluation), :75-91 (derivative).
    """

    def __init__(self, ijk, val, m):
        ijk = np.asarray(ijk)
        if ijk.ndim != 2 or ijk.shape[1] != 3:
            raise CloudAtlasError(1, "ijk must have three columns")           # SparseBilinear.jl:8
        val = np.ascontiguousarray(val, dtype=np.float64)
        if ijk.shape[0] != val.shape[0]:
            raise CloudAtlasError(1, "ijk and val must have same number of rows")  # :9
        self.ijk = np.asfortranarray(ijk, dtype=np.int64)
        self.val = val
        self.m = int(m)
        h = vp()
        check(_create(self.ijk.ctypes.data_as(c_i64p), _f64(self.val), self.ijk.shape[0], self.m, C.byref(h)))
        self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _destroy is not None:  # module globals may already be gone at interpreter exit
            _destroy(h)
            self._h = None

    @classmethod
    def from_dense(cls, N):
        """SparseBilinear(N::Array{T,3}) -- lexicographic (i,j,k), keeps ``!= 0`` (SparseBilinear.jl:24-49)."""
        N = np.asarray(N, dtype=np.float64)
        idx = np.nonzero(N != 0)
        return cls(np.stack(idx, axis=1).astype(np.int64) + 1, N[idx], N.shape[0])

    @property
    def nnz(self):
        return self.ijk.shape[0]

    def info(self):
        v = [C.c_int64() for _ in range(5)]
        check(_info(self._h, *[C.byref(a) for a in v]))
        return dict(zip(("m", "nnz", "nnz_sym", "npairs", "padded_fma_per_state"), (a.value for a in v)))

    # ---- evaluation -------------------------------------------------------------------------
    def __call__(self, x, y=None):
        """N(x,y) / N(x).  1-D input: one state (the reference call).  2-D input (nstates x m):
        the ensemble entry point.  numpy in -> numpy out (host path); CUDA tensor in -> CUDA tensor
        out on the current stream (device path)."""
        if _is_torch(x):
            return self._call_dev(x, y)
        x = np.asarray(x, dtype=np.float64)
        if x.ndim == 1:
            if x.shape[0] != self.m:
                raise CloudAtlasError(1, f"x has length {x.shape[0]}, expected {self.m}")
            x = np.ascontiguousarray(x)
            out = np.empty(self.m)
            yy = None if y is None else np.ascontiguousarray(y, dtype=np.float64)
            if yy is not None and yy.shape != (self.m,):
                raise CloudAtlasError(1, f"y has shape {yy.shape}, expected ({self.m},)")
            check(_eval(self._h, _f64(x), None if yy is None else _f64(yy), _f64(out)))
            return out
        if y is not None:
            raise CloudAtlasError(5, "host ensembles support N(X) only; use CUDA tensors for N(X,Y)")
        if x.shape[1] != self.m:
            raise CloudAtlasError(1, f"ensemble has {x.shape[1]} columns, expected {self.m}")
        X = np.asfortranarray(x)
        OUT = np.empty(X.shape, order="F")
        check(_eval_b(self._h, _f64(X), X.shape[0], _f64(OUT)))
        return OUT

    def _call_dev(self, X, Y=None):
        single = X.dim() == 1
        if single:
            X = X.reshape(1, -1)
            Y = None if Y is None else Y.reshape(1, -1)
        X, n, ld = _colmajor_ensemble(X)
        if X.shape[1] != self.m:
            raise CloudAtlasError(1, f"ensemble has {X.shape[1]} columns, expected {self.m}")
        OUT = ensemble_empty(n, self.m, X.device)
        if Y is None:
            check(_eval_b_dev(self._h, vp(X.data_ptr()), n, ld, vp(OUT.data_ptr()), n, _stream()))
        else:
            Y, n2, ld2 = _colmajor_ensemble(Y, self.m)
            if n2 != n or ld2 != ld:
                raise CloudAtlasError(1, "X and Y must have the same shape and leading dimension")
            check(_eval2_b_dev(self._h, vp(X.data_ptr()), vp(Y.data_ptr()), n, ld, vp(OUT.data_ptr()), n, _stream()))
        return OUT.reshape(-1) if single else OUT

    def jvp(self, X, V):
        """DN(x_s) v_s = N(x_s, v_s) + N(v_s, x_s) for CUDA ensembles (matrix-free Jacobian action)."""
        X, n, ld = _colmajor_ensemble(X, self.m)
        V, n2, ld2 = _colmajor_ensemble(V, self.m)
        if n2 != n or ld2 != ld:
            raise CloudAtlasError(1, "X and V must have the same shape and leading dimension")
        OUT = ensemble_empty(n, self.m, X.device)
        check(_jvp_b_dev(self._h, vp(X.data_ptr()), vp(V.data_ptr()), n, ld, vp(OUT.data_ptr()), n, _stream()))
        return OUT

    def derivative(self, x):
        """derivative(N, x): dense m x m  (SparseBilinear.jl:75-91).  A 2-d ``x`` (nstates x m) is an ensemble: the
        result is nstates x m x m with the state index fastest in memory (a Julia ``Array{Float64,3}``)."""
        if _is_torch(x) and x.dim() == 2:
            return jacobian_ensemble(x, self.m, lambda X, n, ld, D: _jac_b_dev(self._h, X, n, ld, D, n, _stream()))
        if not _is_torch(x) and np.ndim(x) == 2:
            X = np.asfortranarray(np.asarray(x, dtype=np.float64))
            if X.shape[1] != self.m:
                raise CloudAtlasError(1, f"ensemble has {X.shape[1]} columns, expected {self.m}")
            DN = np.empty((X.shape[0], self.m, self.m), order="F")
            check(_jac_b(self._h, _f64(X), X.shape[0], _f64(DN)))
            return DN
        if _is_torch(x):
            import torch
            _check_device_tensor(x)
            if x.dim() != 1 or x.numel() != self.m:
                raise CloudAtlasError(1, f"x has {x.numel()} elements, expected {self.m}")
            DN = torch.empty((self.m, self.m), dtype=torch.float64, device=x.device).t()  # column-major
            check(_jac_dev(self._h, vp(x.contiguous().data_ptr()), vp(DN.data_ptr()), _stream()))
            return DN
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.shape != (self.m,):
            raise CloudAtlasError(1, f"x has shape {x.shape}, expected ({self.m},)")
        DN = np.empty((self.m, self.m), order="F")
        check(_jac(self._h, _f64(x), _f64(DN)))
        return DN


def derivative(N: SparseBilinear, x):
    return N.derivative(x)


def sparse(Ndense) -> SparseBilinear:
    """sparse(N::Array{T,3}) -- SparseBilinear.jl:51."""
    return SparseBilinear.from_dense(Ndense)
