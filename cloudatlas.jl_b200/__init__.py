"""cloudatlas_b200 -- host-side mirror of CloudAtlas.jl's interface for the hot path.

Same names, argument meaning and error behaviour as the reference's Julia API
(src/CloudAtlas.jl:16-32 export list), implemented as thin calls into the C-ABI library
``lib/libcloudatlas_b200.so`` (hand-written sm_100a CUDA).  Julia users bind the same library with
``ccall`` (see INTEGRATION.md and ``julia/CloudAtlasB200.jl``); this Python mirror exists because
Julia is not available in the build image, and is what tests/ and bench.py drive.

The directory name contains a dot, so load it with ``__graft_entry__.load_package()``.
"""
from ._lib import CloudAtlasError, LIB_PATH, lib  # noqa: F401
from .bilinear import SparseBilinear, derivative, ensemble_empty, sparse  # noqa: F401
from .model import ModelOperators, ODEModel, cnab2  # noqa: F401
from .ascio import basis_index_dict, changebasis, ijkl2file, load, save  # noqa: F401
from .continuation import natural_continuation, newton_krylov, pseudo_arclength_continuation  # noqa: F401
from .hookstep import Df_finitediff, SearchParams, hookstepsolve, hookstepsolve_model  # noqa: F401
from .assembly import (AssemblySlab, LibraryComm, all_gather_slabs, assemble, assemble_device, assemble_full,  # noqa: F401
                       assemble_sharded, library_comm_for, row_partition)  # noqa: F401
from .basis import Symmetry, basisComponents, shearVector, basisIndices, halfbox_symmetries, symmetric  # noqa: F401


def launch_count() -> int:
    return int(lib.ca_launch_count())


def device_info():
    import ctypes as C
    from ._lib import check
    nd, sm, ma, mi = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    name = C.create_string_buffer(128)
    lib.ca_device_info.restype = C.c_int32
    check(lib.ca_device_info(C.byref(nd), C.byref(sm), C.byref(ma), C.byref(mi), name, 128))
    return {"ndevices": nd.value, "sms": sm.value, "cc": (ma.value, mi.value), "name": name.value.decode()}


def measure_fp64_peak():
    """(DMMA TFLOP/s, DFMA TFLOP/s) of the current device."""
    import ctypes as C
    from ._lib import _sig, c_f64p, check
    fn = _sig("ca_measure_fp64_peak", c_f64p, c_f64p)
    a, b = C.c_double(), C.c_double()
    check(fn(C.byref(a), C.byref(b)))
    return a.value, b.value
