"""Library-side multi-GPU (csrc/comm.cu): the sharded assembly with its single NCCL all-gather, in both communicator
modes -- one process driving all devices (ca_comm_init_all, the Julia case) and one process per GPU (torchrun; the
NCCL id travels through torch.distributed).  Needs >= 2 GPUs: skipped on a single-GPU box (the builder ran it with
`gpurun --gpus 2`, log under profiles/)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import __graft_entry__ as g

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(scope="module")
def ca():
    return g.load_package()


def nagata(ca):
    sx, sy, sz, tx, tz = ca.halfbox_symmetries()
    return [sx * sy * sz, sz * tx * tz]


def test_single_rank_communicator_is_a_plain_assembly(ca):
    """world = 1: no NCCL needed, the sharded entry point degenerates to ca_assemble."""
    ijkl = ca.basisIndices(2, 2, 3, nagata(ca))
    comm = ca.LibraryComm.init_all(1)
    assert (comm.nranks, comm.nlocal) == (1, 1)
    full = ca.assemble_sharded(comm, 1.0, 2.0, ijkl)[0]
    ref = ca.AssemblySlab(1.0, 2.0, ijkl)
    assert np.array_equal(full.coo_host()[0], ref.coo_host()[0]) and np.array_equal(full.coo_host()[1], ref.coo_host()[1])
    model = ca.ModelOperators.from_assembly(full)
    x = 0.1 * np.random.default_rng(0).standard_normal(44)
    assert np.array_equal(model.f(x, 200.0), ca.ModelOperators(*ref.linear_host(), *ref.coo_host()).f(x, 200.0))
    s, d = model_obs(ca, model, x)
    full_model = ca.ODEModel(1.0, 2.0, 2, 2, 3, nagata(ca))
    assert abs(s - full_model.shear(x)) < 1e-14 and abs(d - full_model.dissipation(x)) < 1e-13


def model_obs(ca, model, x):
    import ctypes as C
    from cloudatlas_b200._lib import _sig, c_f64p, check, vp
    fn = _sig("ca_model_observables_batched", vp, c_f64p, C.c_int64, c_f64p)
    out = np.zeros(2)
    X = np.ascontiguousarray(x.reshape(1, -1))
    check(fn(model._h, X.ctypes.data_as(c_f64p), 1, out.ctypes.data_as(c_f64p)))
    return out[0], out[1]


@pytest.mark.skipif("ngpus() < 2")
@pytest.mark.parametrize("gather", ["bcast", "allgather"])
@pytest.mark.parametrize("JKL,dense", [((2, 2, 3), False), ((3, 3, 12), False), ((2, 2, 3), True)], ids=["44d", "307", "44d-dense"])
def test_one_process_many_devices(ca, JKL, dense, gather, monkeypatch):
    """ca_comm_init_all: every device ends up with the full operator, bit-identical to the single-GPU assembly -- by the
    default gather (grouped in-place broadcasts) and by the padded ncclAllGather + compaction (CA_GATHER=allgather)."""
    import torch
    monkeypatch.setenv("CA_GATHER", gather)
    n = min(ngpus(), 4)
    ijkl = ca.basisIndices(*JKL, nagata(ca))
    ref = ca.AssemblySlab(1.0, 2.0, ijkl, dense=dense)
    rijk, rval = ref.coo_host()
    rlin = ref.linear_host()
    comm = ca.LibraryComm.init_all(n)
    fulls = ca.assemble_sharded(comm, 1.0, 2.0, ijkl, dense=dense)
    assert len(fulls) == n
    x = 0.1 * np.random.default_rng(1).standard_normal(ijkl.shape[0])
    want = None
    for d, full in enumerate(fulls):
        with torch.cuda.device(d):
            assert full.gather_info()["world"] == n
            ijk, val = full.coo_host()
            assert np.array_equal(ijk, rijk) and np.array_equal(val, rval)
            for P, Q in zip(full.linear_host(), rlin):
                assert np.array_equal(P, Q)
            model = ca.ModelOperators.from_assembly(full)
            got = model.f(x, 200.0)
            want = got if want is None else want
            assert np.array_equal(got, want)


@pytest.mark.skipif("ngpus() < 2")
def test_one_process_per_gpu_under_torchrun(ca):
    """torchrun x 2: rank 0's NCCL id through torch.distributed, ca_assemble_sharded, model from the gathered assembly."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tools", "sharded_case.py"), "3", "3", "12", "--check"]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "sharded ok" in out.stdout


def test_newton_krylov_single_gpu_in_the_library(ca):
    """ca_newton_krylov on one GPU (no communicator): converges to the hookstep equilibrium of a 111-mode model; the
    Jacobian actions are counted."""
    from conftest import read_asc
    H = nagata(ca)
    small = ca.ODEModel(1.0, 2.0, 2, 2, 3, H)
    x44, ok = ca.hookstepsolve(small, 200.0, read_asc("xeq1projection-Re200-2-2-3-44d.asc"))
    model = ca.ODEModel(1.0, 2.0, 3, 3, 4, H)
    xg = ca.changebasis(x44, small.ijkl, model.ijkl)
    xe, oke = ca.hookstepsolve(model, 200.0, xg, ca.SearchParams(Nnewton=30))
    assert oke
    # one continuation step in Re by both solvers (from a far guess the two may land on different branches)
    xh, okh = ca.hookstepsolve(model, 205.0, xe, ca.SearchParams(Nnewton=30))
    xk, conv, info = ca.newton_krylov(model, 205.0, xe, ftol=1e-11)
    assert okh and conv, info
    assert info["world"] == 1 and info["jvps"] > 0 and info["device_seconds"] > 0
    assert np.linalg.norm(model.f(xk, 205.0)) < 1e-10
    assert np.linalg.norm(xk - xh) / np.linalg.norm(xh) < 1e-7
    assert info["residuals"][0] > info["residuals"][-1]
    # without the preconditioner GMRES needs far more Jacobian actions on the viscous spectrum
    xu, convu, infou = ca.newton_krylov(model, 205.0, xe, ftol=1e-11, precondition=False)
    assert infou["jvps"] > info["jvps"]


@pytest.mark.skipif("ngpus() < 2")
def test_newton_krylov_row_slab_sharded(ca):
    """SURVEY 8e, last row: the Jacobian action sharded by row slab over the GPUs of one process, one all-gather of m
    doubles per Krylov step, replicated Arnoldi.  Same iterates as the single-GPU solve (the slabs partition the rows;
    every row is computed by the same code on the same numbers)."""
    import torch
    from conftest import read_asc
    H = nagata(ca)
    n = min(ngpus(), 4)
    small = ca.ODEModel(1.0, 2.0, 2, 2, 3, H)
    x44, ok = ca.hookstepsolve(small, 200.0, read_asc("xeq1projection-Re200-2-2-3-44d.asc"))
    ijkl = ca.basisIndices(3, 3, 12, H)
    comm = ca.LibraryComm.init_all(n)
    fulls = ca.assemble_sharded(comm, 1.0, 2.0, ijkl)
    models = []
    for d, full in enumerate(fulls):
        with torch.cuda.device(d):
            models.append(ca.ModelOperators.from_assembly(full))
    with torch.cuda.device(0):   # an equilibrium of the 307-mode model at Re = 200 (global-memory hookstep solver) ...
        xe, ok = ca.hookstepsolve(models[0], 200.0, ca.changebasis(x44, small.ijkl, ijkl), ca.SearchParams(Nnewton=30))
        assert ok
    # ... continued to Re = 210 by Newton-Krylov, sharded and on one GPU
    xs, convs, infos = ca.newton_krylov(models, 210.0, xe, ftol=1e-11, comm=comm)
    assert convs and infos["world"] == n, infos
    with torch.cuda.device(0):
        x1, conv1, info1 = ca.newton_krylov(models[0], 210.0, xe, ftol=1e-11)
        assert conv1 and np.linalg.norm(models[0].f(xs, 210.0)) < 1e-10
    assert np.array_equal(xs, x1)
    assert infos["jvps"] == info1["jvps"] and infos["newton_steps"] == info1["newton_steps"]
