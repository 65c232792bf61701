"""Galerkin assembly -- the constructor half of src/ODEModels.jl:22-51 over the C ABI, sharded by output
mode i across ranks with one all-gather (torch.distributed: NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import CloudAtlasError, _sig, c_f64p, c_i64p, check, vp

_assemble = _sig("ca_assemble", C.c_double, C.c_double, c_i64p, C.c_int64, C.c_int32, C.c_int64, C.c_int64,
                 C.POINTER(vp))
_assemble_dense = _sig("ca_assemble_dense", C.c_double, C.c_double, c_i64p, C.c_int64, C.c_int32, C.c_int64, C.c_int64,
                       C.c_int32, C.c_int32, C.c_int32, C.POINTER(vp))
_assemble_basis = _sig("ca_assemble_basis", C.c_double, C.c_double, c_i64p, C.c_int64, C.c_int32, c_f64p, c_f64p, C.c_int32,
                       C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.POINTER(vp))
_dense_info = _sig("ca_assembly_dense_info", vp, C.POINTER(C.c_int32), c_f64p, c_f64p)
_destroy = _sig("ca_assembly_destroy", vp)
_info = _sig("ca_assembly_info", vp, c_i64p, c_i64p, c_i64p, c_i64p, c_f64p)
_get_linear = _sig("ca_assembly_get_linear", vp, c_f64p, c_f64p, c_f64p)
_get_coo = _sig("ca_assembly_get_coo", vp, c_i64p, c_f64p)
_get_linear_dev = _sig("ca_assembly_get_linear_dev", vp, vp, vp, vp, vp)
_get_diss = _sig("ca_assembly_get_dissipation", vp, c_f64p)
_get_diss_dev = _sig("ca_assembly_get_dissipation_dev", vp, vp, vp)
_get_coo_dev = _sig("ca_assembly_get_coo_dev", vp, vp, vp, vp, vp, vp)


_comm_init_all = _sig("ca_comm_init_all", C.c_int32, C.POINTER(C.c_int32), C.POINTER(vp))
_comm_unique_id = _sig("ca_comm_unique_id", C.c_char_p)
_comm_init_rank = _sig("ca_comm_init_rank", C.c_int32, C.c_int32, C.c_char_p, C.POINTER(vp))
_comm_info = _sig("ca_comm_info", vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32))
_comm_destroy = _sig("ca_comm_destroy", vp)
_assemble_sharded = _sig("ca_assemble_sharded", vp, C.c_double, C.c_double, c_i64p, C.c_int64, C.c_int32, c_f64p, C.c_int32,
                         C.c_int32, C.c_int32, C.c_int32, C.POINTER(vp))
_gather_info = _sig("ca_assembly_gather_info", vp, C.POINTER(C.c_int32), c_f64p, c_i64p)


class LibraryComm:
    """The library's own NCCL communicator (``ca_comm_*``, csrc/comm.cu): the sharded assembly and its all-gather run
    inside the shared library, so a host language needs no NCCL binding."""

    def __init__(self, handle):
        self._h = handle
        v = [C.c_int32() for _ in range(4)]
        check(_comm_info(handle, *[C.byref(a) for a in v]))
        self.nranks, self.nlocal, self.rank, self.nccl_version = (a.value for a in v)

    @classmethod
    def init_all(cls, ngpus, devices=None):
        """One process driving ``ngpus`` devices (``ncclCommInitAll``)."""
        h = vp()
        arr = None if devices is None else (C.c_int32 * ngpus)(*devices)
        check(_comm_init_all(int(ngpus), arr, C.byref(h)))
        return cls(h)

    @classmethod
    def from_torch_group(cls, group=None):
        """One process per GPU: rank 0's NCCL unique id travels through torch.distributed, every rank joins."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        buf = C.create_string_buffer(128)
        if rank == 0:
            check(_comm_unique_id(buf))
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        t = torch.tensor(list(buf.raw), dtype=torch.uint8, device=dev)
        dist.broadcast(t, dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        raw = bytes(t.cpu().tolist())
        h = vp()
        check(_comm_init_rank(world, rank, raw, C.byref(h)))
        return cls(h)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _comm_destroy is not None:
            _comm_destroy(h)
            self._h = None


_group_comms = {}


def assemble_sharded(comm, alpha, gamma, ijkl, normalize=False, mode_scale=None, dense=False, grid=(0, 0, 0)):
    """``ca_assemble_sharded``: list of FULL assemblies, one per local device of ``comm`` (``AssemblySlab`` objects
    with ``row_begin == 0`` and ``row_end == m``), identical on every device."""
    ijkl = np.asfortranarray(ijkl, dtype=np.int64)
    m = ijkl.shape[0]
    sc = None if mode_scale is None else np.ascontiguousarray(mode_scale, dtype=np.float64)
    out = (vp * comm.nlocal)()
    check(_assemble_sharded(comm._h, float(alpha), float(gamma), ijkl.ctypes.data_as(c_i64p), m, int(bool(normalize)),
                            None if sc is None else sc.ctypes.data_as(c_f64p), int(bool(dense)), int(grid[0]), int(grid[1]),
                            int(grid[2]), out))
    return [AssemblySlab._from_handle(vp(out[d]), m) for d in range(comm.nlocal)]


def row_partition(m: int, world: int, align: int = 1):
    """Contiguous ranges of output modes per rank: rank r owns [bounds[r], bounds[r+1]).
    ``align``: range boundaries fall on multiples of it (the dense path computes whole 128-row output tiles, so its
    slabs should not cut a tile in two); trailing ranks may be empty."""
    units = -(-m // align)
    base, extra = divmod(units, world)
    bounds = [0]
    for r in range(world):
        bounds.append(min(m, bounds[-1] + (base + (1 if r < extra else 0)) * align))
    bounds[-1] = m
    return bounds


class AssemblySlab:
    """Rows [row_begin, row_end) of B, A1, A2 and the N entries with i in that range, on the device."""

    def __init__(self, alpha, gamma, ijkl, normalize=False, row_begin=0, row_end=None, dense=False, grid=(0, 0, 0),
                 mode_scale=None, mix=None):
        """dense=True: N by brute-force quadrature on the tensor grid ``grid`` = (Nx, Ny, Nz), 0 = smallest exact
        (the FP64-roofline path, ``ca_assemble_dense``); default: the separable path.
        ``mode_scale`` (m values): the basis s_n Psi_n (``c * Psi``, BasisFunctions.jl:417).  ``mix`` (m x m matrix C,
        dense path only): the basis Phi_n = sum_p Psi_p C[p, n] (``ca_assemble_basis``)."""
        ijkl = np.asfortranarray(ijkl, dtype=np.int64)
        if ijkl.ndim != 2 or ijkl.shape[1] != 4:
            raise CloudAtlasError(1, "ijkl must have four columns")
        self.m = ijkl.shape[0]
        row_end = self.m if row_end is None else row_end
        h = vp()
        if mode_scale is not None or mix is not None:
            sc = None if mode_scale is None else np.ascontiguousarray(mode_scale, dtype=np.float64)
            mx = None if mix is None else np.asfortranarray(mix, dtype=np.float64)
            if sc is not None and sc.shape != (self.m,):
                raise CloudAtlasError(1, f"mode_scale must have {self.m} entries")
            if mx is not None and mx.shape != (self.m, self.m):
                raise CloudAtlasError(1, f"mix must be {self.m} x {self.m}")
            check(_assemble_basis(float(alpha), float(gamma), ijkl.ctypes.data_as(c_i64p), self.m, int(bool(normalize)),
                                  None if sc is None else sc.ctypes.data_as(c_f64p),
                                  None if mx is None else mx.ctypes.data_as(c_f64p), int(bool(dense)),
                                  int(grid[0]), int(grid[1]), int(grid[2]), int(row_begin), int(row_end), C.byref(h)))
        elif dense:
            check(_assemble_dense(float(alpha), float(gamma), ijkl.ctypes.data_as(c_i64p), self.m, int(bool(normalize)),
                                  int(row_begin), int(row_end), int(grid[0]), int(grid[1]), int(grid[2]), C.byref(h)))
        else:
            check(_assemble(float(alpha), float(gamma), ijkl.ctypes.data_as(c_i64p), self.m, int(bool(normalize)),
                            int(row_begin), int(row_end), C.byref(h)))
        self._adopt(h)

    def _adopt(self, h):
        self._h = h
        v = [C.c_int64() for _ in range(4)]
        sec = C.c_double()
        check(_info(h, *[C.byref(a) for a in v], C.byref(sec)))
        _, self.row_begin, self.row_end, self.nnz = (a.value for a in v)
        self.device_seconds = sec.value

    @classmethod
    def _from_handle(cls, h, m):
        self = cls.__new__(cls)
        self.m = m
        self._adopt(h)
        return self

    def gather_info(self):
        """{world, gather_seconds, gather_bytes} of a sharded assembly (``ca_assembly_gather_info``)."""
        w, sec, nb = C.c_int32(), C.c_double(), C.c_int64()
        check(_gather_info(self._h, C.byref(w), C.byref(sec), C.byref(nb)))
        return {"world": w.value, "gather_seconds": sec.value, "gather_bytes": nb.value}

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _destroy is not None:  # module globals may already be gone at interpreter exit
            _destroy(h)
            self._h = None

    def dense_info(self):
        """{grid, contract_seconds, contract_flops, tflops} of the dense quadrature contraction."""
        g = (C.c_int32 * 3)()
        sec, fl = C.c_double(), C.c_double()
        check(_dense_info(self._h, g, C.byref(sec), C.byref(fl)))
        return {"grid": list(g), "contract_seconds": sec.value, "contract_flops": fl.value,
                "tflops": fl.value / sec.value / 1e12 if sec.value > 0 else None}

    def linear_host(self):
        B, A1, A2 = (np.zeros((self.m, self.m), order="F") for _ in range(3))
        check(_get_linear(self._h, *(a.ctypes.data_as(c_f64p) for a in (B, A1, A2))))
        return B, A1, A2

    def dissipation_host(self):
        D = np.zeros((self.m, self.m), order="F")
        check(_get_diss(self._h, D.ctypes.data_as(c_f64p)))
        return D

    def coo_host(self):
        ijk = np.empty((self.nnz, 3), dtype=np.int64, order="F")
        val = np.empty(self.nnz)
        check(_get_coo(self._h, ijk.ctypes.data_as(c_i64p), val.ctypes.data_as(c_f64p)))
        return ijk, val

    def to_torch(self):
        """Slab as CUDA tensors: linear [4, rows, m] float64 (row-major; B, A1, A2, D), idx [3, nnz] int64, val [nnz]."""
        import torch
        rows = self.row_end - self.row_begin
        lin = torch.empty((4, rows, self.m), dtype=torch.float64, device="cuda")
        idx = torch.empty((3, self.nnz), dtype=torch.int64, device="cuda")
        val = torch.empty((self.nnz,), dtype=torch.float64, device="cuda")
        st = vp(torch.cuda.current_stream().cuda_stream)
        check(_get_linear_dev(self._h, vp(lin[0].data_ptr()), vp(lin[1].data_ptr()), vp(lin[2].data_ptr()), st))
        if rows:
            check(_get_diss_dev(self._h, vp(lin[3].data_ptr()), st))
        check(_get_coo_dev(self._h, vp(idx[0].data_ptr()), vp(idx[1].data_ptr()), vp(idx[2].data_ptr()),
                           vp(val.data_ptr()), st))
        return lin, idx, val


def all_gather_slabs(lin, idx, val, m, group=None):
    """The one collective of the assembly path: every rank ends up with the full operators.

    lin: [nmat, rows_r, m], idx: [3, nnz_r], val: [nnz_r] torch tensors of this rank (any device the
    process group supports).  Slab sizes differ per rank, so sizes are exchanged first; then every rank's slab
    goes straight into its place in the full arrays -- one broadcast per contiguous piece (a matrix's rows, an index
    column, the values), all issued asynchronously on the collective stream: no padding to the largest slab, no
    staging buffers, no concatenation pass over the 10 GB of a 3000-mode operator.  Rank order is the
    reference's lexicographic (i,j,k) order because ranks own contiguous, increasing ranges of i."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    me = dist.get_rank(group)
    dev = val.device
    nmat = lin.shape[0]
    sizes = torch.tensor([lin.shape[1], val.shape[0]], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = torch.stack(all_sizes).cpu()
    rows = [int(v) for v in all_sizes[:, 0]]
    nnzs = [int(v) for v in all_sizes[:, 1]]
    roff = [0] + list(np.cumsum(rows))
    noff = [0] + list(np.cumsum(nnzs))
    lin_all = torch.empty((nmat, int(roff[-1]), m), dtype=lin.dtype, device=dev)
    idx_all = torch.empty((3, int(noff[-1])), dtype=idx.dtype, device=dev)
    val_all = torch.empty((int(noff[-1]),), dtype=val.dtype, device=dev)
    lin_all[:, roff[me]:roff[me + 1]] = lin
    idx_all[:, noff[me]:noff[me + 1]] = idx
    val_all[noff[me]:noff[me + 1]] = val
    pending = []
    for r in range(world):
        src = dist.get_global_rank(group, r) if group is not None else r
        pieces = [lin_all[k, roff[r]:roff[r + 1]] for k in range(nmat)] if rows[r] else []
        if nnzs[r]:
            pieces += [idx_all[c, noff[r]:noff[r + 1]] for c in range(3)] + [val_all[noff[r]:noff[r + 1]]]
        for piece in pieces:
            pending.append(dist.broadcast(piece, src, group=group, async_op=True))
    for w in pending:
        w.wait()
    return lin_all, idx_all, val_all


def library_comm_for(group=None):
    """The library communicator (``LibraryComm``) that mirrors a torch.distributed group; created on first use."""
    key = id(group)
    if key not in _group_comms:
        _group_comms[key] = LibraryComm.from_torch_group(group)
    return _group_comms[key]


def assemble_full(alpha, gamma, ijkl, normalize=False, group=None, distributed=None, dense=False, grid=(0, 0, 0),
                  mode_scale=None):
    """(assembly, stats): the full operator as an ``AssemblySlab`` on this process's device.  Single process: one
    ``ca_assemble`` call.  Under torch.distributed with the NCCL backend: ``ca_assemble_sharded`` through the library's
    own communicator (rows sharded by output mode, one grouped all-gather over NVLink, csrc/comm.cu)."""
    m = np.asarray(ijkl).shape[0]
    if distributed is None:
        import torch.distributed as dist
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if not distributed:
        slab = AssemblySlab(alpha, gamma, ijkl, normalize, dense=dense, grid=grid, mode_scale=mode_scale)
        return slab, {"world": 1, "gather_seconds": 0.0, "gather_bytes": 0, "device_seconds": slab.device_seconds,
                      "nnz": slab.nnz, "dense": slab.dense_info() if dense else None}
    slab = assemble_sharded(library_comm_for(group), alpha, gamma, ijkl, normalize, mode_scale, dense, grid)[0]
    st = slab.gather_info()
    st.update(device_seconds=slab.device_seconds, nnz=slab.nnz, dense=slab.dense_info() if dense else None)
    return slab, st


def assemble_device(alpha, gamma, ijkl, normalize=False, group=None, distributed=None, dense=False, grid=(0, 0, 0),
                    mode_scale=None):
    """The assembled operators as CUDA tensors, for ``ModelOperators.from_device``: (lin, idx, val, stats) with
    ``lin`` [4, m, m] (B, A1, A2, D; each COLUMN-major, i.e. ``lin[q].t()`` is the matrix), ``idx`` [3, nnz] int64
    (1-based i, j, k in the reference's lexicographic order) and ``val`` [nnz].  With torch.distributed initialised
    every rank assembles its range of output modes and the slabs are exchanged with one all-gather over NCCL; the
    gathered operator stays on the device -- the host sees none of it."""
    import time
    import torch
    m = np.asarray(ijkl).shape[0]
    if distributed is None:
        import torch.distributed as dist
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    stats = {"world": 1, "gather_seconds": 0.0, "gather_bytes": 0}
    if not distributed:
        slab = AssemblySlab(alpha, gamma, ijkl, normalize, dense=dense, grid=grid, mode_scale=mode_scale)
        lin, idx, val = slab.to_torch()
    else:
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        bounds = row_partition(m, world, 128 if dense else 1)
        slab = AssemblySlab(alpha, gamma, ijkl, normalize, bounds[rank], bounds[rank + 1], dense=dense, grid=grid,
                            mode_scale=mode_scale)
        lin, idx, val = slab.to_torch()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        lin, idx, val = all_gather_slabs(lin, idx, val, m, group)
        ev1.record()
        torch.cuda.synchronize()
        stats.update(world=world, gather_seconds=ev0.elapsed_time(ev1) * 1e-3, gather_wall_seconds=time.perf_counter() - t0,
                     gather_bytes=int(lin.numel() * 8 + idx.numel() * 8 + val.numel() * 8))
    stats.update(device_seconds=slab.device_seconds, nnz_local=slab.nnz, dense=slab.dense_info() if dense else None)
    lin = lin.transpose(1, 2).contiguous()   # row-major slabs -> column-major matrices (4 m^2 doubles, on the device)
    return lin, idx, val, stats


def assemble(alpha, gamma, ijkl, normalize=False, group=None, distributed=None, dense=False, grid=(0, 0, 0)):
    """Full B, A1, A2 (m x m, column-major numpy) and N's COO (ijk nnz x 3 1-based, val) on every rank.

    With torch.distributed initialised (or ``distributed=True``) each rank assembles its range of output
    modes on its own GPU and the slabs are exchanged with one all-gather; otherwise a single device does
    everything.  Returns (B, A1, A2, ijk, val, stats)."""
    import torch
    m = np.asarray(ijkl).shape[0]
    if distributed is None:
        import torch.distributed as dist
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if not distributed:
        slab = AssemblySlab(alpha, gamma, ijkl, normalize, dense=dense, grid=grid)
        B, A1, A2 = slab.linear_host()
        ijk, val = slab.coo_host()
        return B, A1, A2, ijk, val, {"device_seconds": slab.device_seconds, "world": 1, "nnz_local": slab.nnz,
                                     "D": slab.dissipation_host(), "dense": slab.dense_info() if dense else None}
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    bounds = row_partition(m, world, 128 if dense else 1)
    slab = AssemblySlab(alpha, gamma, ijkl, normalize, bounds[rank], bounds[rank + 1], dense=dense, grid=grid)
    lin, idx, val = slab.to_torch()
    lin, idx, val = all_gather_slabs(lin, idx, val, m, group)
    lin = lin.cpu().numpy()
    B, A1, A2, D = (np.asfortranarray(lin[q]) for q in range(4))
    ijk = np.asfortranarray(idx.t().cpu().numpy())
    return B, A1, A2, ijk, val.cpu().numpy(), {"device_seconds": slab.device_seconds, "world": world,
                                               "nnz_local": slab.nnz, "D": D,
                                               "dense": slab.dense_info() if dense else None}
