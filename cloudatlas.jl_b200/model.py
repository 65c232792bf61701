"""ODEModel evaluation side -- mirror of src/ODEModels.jl:6-63 over the C ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import CloudAtlasError, _sig, c_f64p, c_i64p, check, vp
from .bilinear import (SparseBilinear, _check_device_tensor, _colmajor_ensemble, _f64, _is_torch, _stream,
                       ensemble_empty)

_create = _sig("ca_model_create", c_f64p, c_f64p, c_f64p, C.c_int64, c_i64p, c_f64p, C.c_int64, C.POINTER(vp))
_create_dev = _sig("ca_model_create_dev", vp, vp, vp, C.c_int64, vp, vp, vp, vp, C.c_int64, vp, C.POINTER(vp))
_create_from_assembly = _sig("ca_model_create_from_assembly", vp, C.POINTER(vp))
_build_info = _sig("ca_model_build_info", vp, C.POINTER(C.c_int32), c_f64p, c_i64p, c_i64p, c_i64p)
_digest = _sig("ca_model_digest", vp, C.POINTER(C.c_uint64))
_destroy = _sig("ca_model_destroy", vp)
_info = _sig("ca_model_info", vp, c_i64p, c_i64p, c_i64p, c_i64p, c_i64p, c_i64p)
_rhs = _sig("ca_model_rhs", vp, c_f64p, C.c_double, c_f64p)
_rhs_b = _sig("ca_model_rhs_batched", vp, c_f64p, C.c_int64, C.c_double, c_f64p)
_rhs_b_dev = _sig("ca_model_rhs_batched_dev", vp, vp, C.c_int64, C.c_int64, C.c_double, vp, C.c_int64, vp)
_jac = _sig("ca_model_jac", vp, c_f64p, C.c_double, c_f64p)
_jac_dev = _sig("ca_model_jac_dev", vp, vp, C.c_double, vp, vp)
_jac_b = _sig("ca_model_jac_batched", vp, c_f64p, C.c_int64, C.c_double, c_f64p)
_jac_b_dev = _sig("ca_model_jac_batched_dev", vp, vp, C.c_int64, C.c_int64, C.c_double, vp, C.c_int64, vp)
_cnab2 = _sig("ca_model_cnab2_batched_dev", vp, vp, C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_int64,
              C.c_int64, vp, vp, vp)
_set_obs = _sig("ca_model_set_observables", vp, c_f64p, c_f64p)
_obs_dev = _sig("ca_model_observables_batched_dev", vp, vp, C.c_int64, C.c_int64, vp, vp)
_jvp_b_host = _sig("ca_model_jvp_batched", vp, c_f64p, c_f64p, C.c_int64, C.c_double, c_f64p)
_cnab2_host = _sig("ca_model_cnab2_batched", vp, c_f64p, C.c_int64, C.c_double, C.c_double, C.c_int64, C.c_int64,
                   c_f64p, c_f64p)
_obs_host = _sig("ca_model_observables_batched", vp, c_f64p, C.c_int64, c_f64p)
_jvp_b_dev = _sig("ca_model_jvp_batched_dev", vp, vp, vp, C.c_int64, C.c_int64, C.c_double, vp, C.c_int64, vp)


class ModelOperators:
    """Device side of an ODEModel built from given operators (B, A1, A2 dense; N COO).

    ``f(x, R)`` and ``Df(x, R)`` have the reference closures' meaning (ODEModels.jl:57-58); ``f`` also
    accepts an ensemble (nstates x m), host (numpy) or device (CUDA tensor, column-major).
    """

    def __init__(self, B, A1, A2, ijk, val):
        self._dev = None
        self._B = np.asfortranarray(B, dtype=np.float64)
        self._A1 = np.asfortranarray(A1, dtype=np.float64)
        self._A2 = np.asfortranarray(A2, dtype=np.float64)
        m = self._B.shape[0]
        if self._B.shape != (m, m) or self._A1.shape != (m, m) or self._A2.shape != (m, m):
            raise CloudAtlasError(1, "B, A1, A2 must be square and of equal size")
        ijk = np.asarray(ijk)
        if ijk.ndim != 2 or ijk.shape[1] != 3:
            raise CloudAtlasError(1, "ijk must have three columns")
        self._ijk = np.asfortranarray(ijk, dtype=np.int64)
        self._val = np.ascontiguousarray(val, dtype=np.float64)
        if self._ijk.shape[0] != self._val.shape[0]:
            raise CloudAtlasError(1, "ijk and val must have same number of rows")
        self.m = m
        h = vp()
        check(_create(_f64(self._B), _f64(self._A1), _f64(self._A2), m, self._ijk.ctypes.data_as(c_i64p),
                      _f64(self._val), self._ijk.shape[0], C.byref(h)))
        self._h = h

    @classmethod
    def from_device(cls, B, A1, A2, idx, val):
        """Model from operators that are already on the GPU (``ca_model_create_dev``): B, A1, A2 column-major m x m
        CUDA tensors (stride (1, m)), ``idx`` [3, nnz] int64 (rows i, j, k; 1-based, lexicographic), ``val`` [nnz].
        Nothing passes through the host; numpy copies (``.B``, ``.ijk``, ...) are fetched only if asked for."""
        import torch
        self = cls.__new__(cls)
        m = B.shape[0]
        for M in (B, A1, A2):
            if not (M.is_cuda and M.dtype == torch.float64 and tuple(M.shape) == (m, m) and (m == 1 or M.stride() == (1, m))):
                raise CloudAtlasError(1, "B, A1, A2 must be column-major m x m float64 CUDA tensors")
        if not (idx.is_cuda and idx.dtype == torch.int64 and idx.dim() == 2 and idx.shape[0] == 3 and idx.is_contiguous()):
            raise CloudAtlasError(1, "idx must be a contiguous [3, nnz] int64 CUDA tensor")
        if not (val.is_cuda and val.dtype == torch.float64 and val.dim() == 1 and val.is_contiguous()
                and val.shape[0] == idx.shape[1]):
            raise CloudAtlasError(1, "ijk and val must have same number of rows")
        self.m = m
        self._dev = (B, A1, A2, idx, val)
        self._B = self._A1 = self._A2 = self._ijk = self._val = None
        nnz = int(val.shape[0])
        h = vp()
        check(_create_dev(vp(B.data_ptr()), vp(A1.data_ptr()), vp(A2.data_ptr()), m, vp(idx[0].data_ptr()),
                          vp(idx[1].data_ptr()), vp(idx[2].data_ptr()), vp(val.data_ptr()), nnz, _stream(), C.byref(h)))
        torch.cuda.current_stream().synchronize()
        self._h = h
        return self

    @classmethod
    def from_assembly(cls, slab):
        """Model straight from a full assembly handle (``ca_model_create_from_assembly``): assembled operator ->
        folded, packed model on the assembly's device; the observables come along.  Host copies of the operators are
        fetched from the assembly only if asked for."""
        self = cls.__new__(cls)
        self.m = slab.m
        self._dev = None
        self._slab = slab
        self._B = self._A1 = self._A2 = self._ijk = self._val = None
        h = vp()
        check(_create_from_assembly(slab._h, C.byref(h)))
        self._h = h
        return self

    def release_device_inputs(self):
        """Drops the CUDA tensors ``from_device`` was given (the library keeps its own compact copies)."""
        if self._dev is not None:
            self.B, self.A1, self.A2, self.ijk, self.val  # noqa: B018  (materialise the host copies first)
            self._dev = None

    def _host(self, name, q):
        v = getattr(self, name)
        if v is None and getattr(self, "_slab", None) is not None:
            if q < 3:
                self._B, self._A1, self._A2 = self._slab.linear_host()
            else:
                self._ijk, self._val = self._slab.coo_host()
            return getattr(self, name)
        if v is None and self._dev is not None:
            t = self._dev[q]
            if q < 3:
                v = np.asfortranarray(t.cpu().numpy())
            elif q == 3:
                v = np.asfortranarray(t.t().cpu().numpy())
            else:
                v = t.cpu().numpy()
            setattr(self, name, v)
        return v

    B = property(lambda self: self._host("_B", 0))
    A1 = property(lambda self: self._host("_A1", 1))
    A2 = property(lambda self: self._host("_A2", 2))
    ijk = property(lambda self: self._host("_ijk", 3))
    val = property(lambda self: self._host("_val", 4))

    def build_info(self):
        """How the model was built (``ca_model_build_info``): device or host path and seconds per phase."""
        on = C.c_int32()
        sec = (C.c_double * 9)()
        ncl, nt, ntm = C.c_int64(), C.c_int64(), C.c_int64()
        check(_build_info(self._h, C.byref(on), sec, C.byref(ncl), C.byref(nt), C.byref(ntm)))
        names = ("total", "blocks", "fold", "symmetrise", "cluster", "tile_unions", "layout_host", "place", "stream")
        return {"on_device": bool(on.value), "seconds": dict(zip(names, (float(v) for v in sec))),
                "nclusters": ncl.value, "ntiles": nt.value, "nterms": ntm.value}

    def digest(self):
        """Twelve FNV-1a digests of the build products (``ca_model_digest``)."""
        d = (C.c_uint64 * 12)()
        check(_digest(self._h, d))
        return [int(v) for v in d]

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _destroy is not None:  # module globals may already be gone at interpreter exit
            _destroy(h)
            self._h = None

    def __len__(self):
        return self.m

    def info(self):
        v = [C.c_int64() for _ in range(6)]
        check(_info(self._h, *[C.byref(a) for a in v]))
        return dict(zip(("m", "nnz_q_sym", "nnz_lin", "npairs", "padded_fma_per_state", "nblocks_B"),
                        (a.value for a in v)))

    def kernel_path(self, nstates):
        """(path, generated_source, nvrtc_log): which kernel serves an ensemble of this size."""
        fn = _sig("ca_model_kernel_path", vp, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_char_p), C.POINTER(C.c_char_p))
        path, src, log = C.c_int32(), C.c_char_p(), C.c_char_p()
        check(fn(self._h, int(nstates), C.byref(path), C.byref(src), C.byref(log)))
        names = {0: "dmma_row_tiles", 1: "csr_stream", 2: "nvrtc_straight_line", 3: "dmma_row_tiles_ksplit",
                 4: "dmma_row_tiles_windowed"}
        return names[path.value], (src.value or b"").decode(), (log.value or b"").decode()

    def f(self, x, R, out=None):
        if _is_torch(x):
            single = x.dim() == 1
            X = x.reshape(1, -1) if single else x
            X, n, ld = _colmajor_ensemble(X, self.m)
            OUT = ensemble_empty(n, self.m, X.device) if out is None else out
            _, n_o, ldo = _colmajor_ensemble(OUT, self.m)
            if n_o != n or OUT.shape[1] != self.m:
                raise CloudAtlasError(1, "out must be nstates x m, column-major")
            check(_rhs_b_dev(self._h, vp(X.data_ptr()), n, ld, float(R), vp(OUT.data_ptr()), ldo, _stream()))
            return OUT.reshape(-1) if single else OUT
        x = np.asarray(x, dtype=np.float64)
        if x.ndim == 1:
            if x.shape[0] != self.m:
                raise CloudAtlasError(1, f"x has length {x.shape[0]}, expected {self.m}")
            out = np.empty(self.m)
            check(_rhs(self._h, _f64(np.ascontiguousarray(x)), float(R), _f64(out)))
            return out
        if x.shape[1] != self.m:
            raise CloudAtlasError(1, f"ensemble has {x.shape[1]} columns, expected {self.m}")
        X = np.asfortranarray(x)
        OUT = np.empty(X.shape, order="F")
        check(_rhs_b(self._h, _f64(X), X.shape[0], float(R), _f64(OUT)))
        return OUT

    def f_into(self, X, R, OUT):
        """Host ensemble into a caller-owned (e.g. page-locked) column-major buffer."""
        for A, name in ((X, "X"), (OUT, "OUT")):
            if not (isinstance(A, np.ndarray) and A.dtype == np.float64 and A.ndim == 2 and A.shape[1] == self.m
                    and (A.flags.f_contiguous or A.shape[0] == 1)):
                raise CloudAtlasError(1, f"{name} must be a column-major float64 array with {self.m} columns")
        if OUT.shape != X.shape:
            raise CloudAtlasError(1, "OUT must have the shape of X")
        check(_rhs_b(self._h, _f64(X), X.shape[0], float(R), _f64(OUT)))
        return OUT

    def Df(self, x, R):
        """Df(x,R) (ODEModels.jl:58); a 2-d ``x`` (nstates x m) gives nstates x m x m, state index fastest."""
        if _is_torch(x) and x.dim() == 2:
            from .bilinear import jacobian_ensemble
            return jacobian_ensemble(x, self.m, lambda X, n, ld, D: _jac_b_dev(self._h, X, n, ld, float(R), D, n, _stream()))
        if not _is_torch(x) and np.ndim(x) == 2:
            X = np.asfortranarray(np.asarray(x, dtype=np.float64))
            if X.shape[1] != self.m:
                raise CloudAtlasError(1, f"ensemble has {X.shape[1]} columns, expected {self.m}")
            D = np.empty((X.shape[0], self.m, self.m), order="F")
            check(_jac_b(self._h, _f64(X), X.shape[0], float(R), _f64(D)))
            return D
        if _is_torch(x):
            import torch
            _check_device_tensor(x)
            if x.dim() != 1 or x.numel() != self.m:
                raise CloudAtlasError(1, f"x has {x.numel()} elements, expected {self.m}")
            D = torch.empty((self.m, self.m), dtype=torch.float64, device=x.device).t()
            check(_jac_dev(self._h, vp(x.contiguous().data_ptr()), float(R), vp(D.data_ptr()), _stream()))
            return D
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.shape != (self.m,):
            raise CloudAtlasError(1, f"x has shape {x.shape}, expected ({self.m},)")
        D = np.empty((self.m, self.m), order="F")
        check(_jac(self._h, _f64(x), float(R), _f64(D)))
        return D

    def Df_action(self, X, V, R):
        """Df(x_s, R) v_s, matrix-free: CUDA ensembles (device path) or numpy ensembles (host-pointer path)."""
        if not _is_torch(X):
            Xh = np.asfortranarray(np.atleast_2d(np.asarray(X, dtype=np.float64)))
            Vh = np.asfortranarray(np.atleast_2d(np.asarray(V, dtype=np.float64)))
            if Xh.shape[1] != self.m or Vh.shape != Xh.shape:
                raise CloudAtlasError(1, f"X and V must both be nstates x {self.m}")
            OUT = np.empty(Xh.shape, order="F")
            check(_jvp_b_host(self._h, _f64(Xh), _f64(Vh), Xh.shape[0], float(R), _f64(OUT)))
            return OUT[0] if np.ndim(X) == 1 else OUT
        X, n, ld = _colmajor_ensemble(X, self.m)
        V, n2, ld2 = _colmajor_ensemble(V, self.m)
        if (n2, ld2) != (n, ld):
            raise CloudAtlasError(1, "X and V must have the same shape and leading dimension")
        OUT = ensemble_empty(n, self.m, X.device)
        check(_jvp_b_dev(self._h, vp(X.data_ptr()), vp(V.data_ptr()), n, ld, float(R), vp(OUT.data_ptr()), n, _stream()))
        return OUT


def cnab2(x0, odemodel, R, dt, Nsteps, nsave, verbosity=0):
    """cnab2(x0, odemodel, R, dt, Nsteps, nsave) -> (t, X)  --  ODEModels.jl:147-181.

    x0 1-D: one trajectory, X is Nsave x m like the reference's (including its quirks: t = (0:Nsave-1)*dt, and
    row 0 is the first SAVED step because the reference overwrites its x0 row).  x0 2-D (nstates x m, numpy or
    column-major CUDA tensor): an ensemble of trajectories, X is Nsave x nstates x m."""
    import torch
    Nsave = int(Nsteps) // int(nsave)
    single = (x0.dim() if _is_torch(x0) else np.ndim(x0)) == 1
    if _is_torch(x0):
        X0 = x0.reshape(1, -1) if single else x0
    else:
        X0 = torch.from_numpy(np.ascontiguousarray(np.atleast_2d(np.asarray(x0, dtype=np.float64)).T)).cuda().t()
    X0, n, ld = _colmajor_ensemble(X0, odemodel.m)
    save = torch.empty((max(Nsave, 1), odemodel.m, n), dtype=torch.float64, device=X0.device)
    check(_cnab2(odemodel._h, vp(X0.data_ptr()), n, ld, float(R), float(dt), int(Nsteps), int(nsave),
                 vp(save.data_ptr()), None, _stream()))
    X = save[:Nsave].permute(0, 2, 1)                      # Nsave x nstates x m (views of column-major snapshots)
    t = np.arange(Nsave) * dt
    if not _is_torch(x0):
        X = X.cpu().numpy()
    return (t, X[:, 0, :]) if single else (t, X)


class ODEModel(ModelOperators):
    """ODEModel(alpha, gamma, J, K, L, H; normalize=false) -- ODEModels.jl:6-61.

    Fields as in the reference struct: ``alpha, gamma, H, ijkl, Psi, B, A1, A2, N, f, Df`` (``Psi`` is the
    tabular basis, ``Bfact`` is folded into the device operators).  The constructor is the assembly
    driver: index set -> basis tables -> Galerkin assembly on the GPU(s) -> device packing."""

    def __init__(self, alpha, gamma, J, K, L, H, normalize=False, ijkl=None, group=None, verbose=False,
                 dense_assembly=False, grid=(0, 0, 0), mode_scale=None):
        from .assembly import assemble_full
        from .basis import basisComponents, basisIndices
        self.alpha, self.gamma = float(alpha), float(gamma)
        self.H = None if H is None else list(H)
        self.ijkl = basisIndices(J, K, L, self.H) if ijkl is None else np.asfortranarray(ijkl, dtype=np.int64)
        m = self.ijkl.shape[0]
        if verbose:  # the reference prints these (ODEModels.jl:25-26)
            print(f"J,K,L,m == {J},{K},{L},{m}")
            print(f"(2J+1)(2K+1)(2L+1) + 1 == {(2 * J + 1) * (2 * K + 1) * (2 * L + 1) + 1}")
        self.Psi = basisComponents(self.alpha, self.gamma, self.ijkl, normalize)
        slab, self.assembly_stats = assemble_full(self.alpha, self.gamma, self.ijkl, normalize, group,
                                                  dense=dense_assembly, grid=grid, mode_scale=mode_scale)
        self.mode_scale = None if mode_scale is None else np.asarray(mode_scale, dtype=np.float64)
        built = ModelOperators.from_assembly(slab)   # installs the observables as well
        self.__dict__.update(built.__dict__)
        built._h = None   # the handle now belongs to this object
        self._N = None
        from .basis import shearVector
        self.Psishear = shearVector(self.alpha, self.gamma, self.ijkl, normalize)   # ODEModels.jl:65-78
        if self.mode_scale is not None:
            self.Psishear = self.Psishear * self.mode_scale
        self.D = slab.dissipation_host()                                             # ODEModels.jl:129-137

    def observables(self, x):
        """(shear, dissipation) of shear_function / dissipation_function (ODEModels.jl:65-137) for one state
        (returns two floats) or an ensemble nstates x m (returns an nstates x 2 array: numpy in -> numpy out,
        column-major CUDA tensor in -> CUDA tensor out)."""
        import torch
        if _is_torch(x):
            X = x.reshape(1, -1) if x.dim() == 1 else x
        else:
            X = torch.from_numpy(np.ascontiguousarray(np.atleast_2d(np.asarray(x, dtype=np.float64)).T)).cuda().t()
        X, n, ld = _colmajor_ensemble(X, self.m)
        OUT = ensemble_empty(n, 2, X.device)
        check(_obs_dev(self._h, vp(X.data_ptr()), n, ld, vp(OUT.data_ptr()), _stream()))
        if _is_torch(x):
            return OUT[0] if x.dim() == 1 else OUT
        out = OUT.cpu().numpy()
        return (float(out[0, 0]), float(out[0, 1])) if np.ndim(x) == 1 else out

    def shear(self, x):
        """shear(x) = 1 + dot(x, Psi_shear): the wall shear rate I of the notebooks' bifurcation diagrams."""
        o = self.observables(x)
        return o[0] if isinstance(o, tuple) else o[..., 0]

    def dissipation(self, x):
        """dissipation(x) = 1 + x' D x."""
        o = self.observables(x)
        return o[1] if isinstance(o, tuple) else o[..., 1]

    @property
    def N(self) -> SparseBilinear:
        """The unfolded quadratic operator as a SparseBilinear (packed on first use)."""
        if self._N is None:
            self._N = SparseBilinear(self.ijk, self.val, self.m)
        return self._N
