"""Newton-hookstep through the CUDA path: the reference's own solver tests (test/runtests.jl:29-107) and
the notebooks' solves, device-resident and closure form, against the stored equilibria and the oracle."""
import numpy as np
import pytest

import __graft_entry__ as g
from conftest import read_asc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ca():
    return g.load_package()


@pytest.fixture(scope="module")
def H(ca):
    sx, sy, sz, tx, tz = ca.halfbox_symmetries()
    return [sx * sy * sz, sz * tx * tz]


@pytest.fixture(scope="module")
def model17(ca, H):
    return ca.ODEModel(1.0, 2.0, 1, 1, 3, H)


@pytest.fixture(scope="module")
def model17n(ca, H):
    return ca.ODEModel(1.0, 2.0, 1, 1, 3, H, normalize=True)


def test_hookstep_2d_closure_form(ca, known):
    # test/runtests.jl:29-38 (generic closures; also the finite-difference form, Hookstep.jl:95)
    k = known["hookstep2d"]
    f = lambda x: np.array([x[0] ** 2 + x[1] ** 2 - 1, x[0] - x[1]])
    Df = lambda x: np.array([[2 * x[0], 2 * x[1]], [1.0, -1.0]])
    x, ok = ca.hookstepsolve(f, Df, np.array(k["xguess"]))
    assert np.linalg.norm(x - np.array(k["xsolution"])) < k["tol"]
    x2, _ = ca.hookstepsolve(f, np.array(k["xguess"]))
    assert np.linalg.norm(x2 - np.array(k["xsolution"])) < 1e-6


def test_17d_device_resident(ca, known, model17):
    # test/runtests.jl:61-82
    k = known["m17_unnormalized"]
    x, ok, log = ca.hookstepsolve(model17, 200.0, np.array(k["xguess"]), return_log=True)
    xe = np.array(k["xeqb"])
    assert np.linalg.norm(x - xe) / np.linalg.norm(xe) < k["solve_tol"]
    assert log["f_evals"] >= log["newton_steps"] and log["device_seconds"] > 0


def test_17d_normalized_device_resident(ca, known, model17n):
    # test/runtests.jl:84-107 (delta = 0.01) and find-eqb.ipynb:306-351 (default params, success = true)
    k = known["m17_normalized"]
    xe = np.array(k["xeqb"])
    x, ok = ca.hookstepsolve(model17n, 200.0, np.array(k["xguess"]), ca.SearchParams(delta=k["delta"]))
    assert np.linalg.norm(x - xe) / np.linalg.norm(xe) < k["solve_tol"]
    x, ok = ca.hookstepsolve(model17n, 200.0, np.array(k["xguess"]))
    assert ok
    xs = np.array(k["xsoln_default_params"])
    assert np.linalg.norm(x - xs) / np.linalg.norm(xs) < 1e-9
    assert np.linalg.norm(model17n.f(x, 200.0)) / np.linalg.norm(x) < 1e-10


def test_device_search_follows_the_oracle_step_by_step(ca, known, model17n, nagata_H):
    """Same decision tree: per-Newton-step residuals, trust radii and hookstep counts agree with the
    oracle's restatement of Hookstep.jl run on the oracle's own model."""
    from oracle.hookstep import SearchParams as RP, hookstepsolve as ref_solve
    from oracle.model import ODEModel as Ref
    ref = Ref(1.0, 2.0, 1, 1, 3, nagata_H, normalize=True)
    xg = np.array(known["m17_normalized"]["xguess"])
    for delta in (0.02, 0.01, 0.1, 0.003):
        rlog = []
        xr, okr = ref_solve(lambda x: ref.f(x, 200.0), lambda x: ref.Df(x, 200.0), xg, RP(delta=delta), log=rlog)
        x, ok, log = ca.hookstepsolve(model17n, 200.0, xg, ca.SearchParams(delta=delta), return_log=True)
        assert ok == okr
        assert np.linalg.norm(x - xr) / np.linalg.norm(xr) < 1e-9
        assert len(rlog) == len(log["residual"])
        for (res, dl, nh), r2, d2, n2 in zip(rlog, log["residual"], log["delta"], log["hooksteps"]):
            assert abs(res - r2) <= 1e-6 * max(res, 1e-12) + 1e-13
            assert abs(dl - d2) <= 1e-6 * dl
            assert nh == n2


def test_closure_form_with_device_closures(ca, known, model17):
    k = known["m17_unnormalized"]
    x, ok = ca.hookstepsolve(lambda x: model17.f(x, 200.0), lambda x: model17.Df(x, 200.0), np.array(k["xguess"]))
    xe = np.array(k["xeqb"])
    assert np.linalg.norm(x - xe) / np.linalg.norm(xe) < k["solve_tol"]


def test_27d_and_44d_from_shipped_dns_projections(ca, known, H):
    # continue-eqb.ipynb:115-143 (27d) and the 44d guess files of config 2
    for key, jkl in (("m27", (1, 2, 3)), ("m44", (2, 2, 3))):
        model = ca.ODEModel(1.0, 2.0, *jkl, H)
        xg = read_asc(known[key]["guess_file"])
        x, ok, log = ca.hookstepsolve(model, 200.0, xg, return_log=True)
        assert np.linalg.norm(model.f(x, 200.0)) / np.linalg.norm(x) < 1e-8, log
    # nested bases: the 27d file is the head of the 44d file
    assert np.array_equal(read_asc("xeq1projection-Re200-2-2-3-44d.asc")[:27], read_asc(known["m27"]["guess_file"]))


def test_batched_searches_re_sweep(ca, known, model17):
    """Independent searches in one launch: a natural-parameter sweep in Re (config 2's continuation,
    SURVEY 8f rank 2) started from the Re = 200 equilibrium."""
    xe = np.array(known["m17_unnormalized"]["xeqb"])
    Rs = np.linspace(190.0, 210.0, 9)
    X, ok, logs = ca.hookstepsolve(model17, Rs, np.tile(xe, (9, 1)), ca.SearchParams(Nnewton=30), return_log=True)
    assert ok.all(), [l["exit_reason"] for l in logs]
    for R, x in zip(Rs, X):
        assert np.linalg.norm(model17.f(x, R)) < 1e-8
    assert np.linalg.norm(X[4] - xe) / np.linalg.norm(xe) < 1e-9
    single, _ = ca.hookstepsolve(model17, Rs[2], xe, ca.SearchParams(Nnewton=30))
    assert np.array_equal(single, X[2])           # bit-reproducible, independent of the batch


def test_larger_models_use_the_host_driven_search(ca, known, H):
    """m > 95: the C entry point refuses (shared memory), the mirror drives Hookstep.jl's loop from the host with every
    f and Df evaluated on the device.  The 44d equilibrium, prolonged to a 111-mode basis, converges there."""
    model = ca.ODEModel(1.0, 2.0, 3, 3, 4, H)
    assert model.m > 95
    from cloudatlas_b200.hookstep import _solve, _CParams
    import ctypes as C
    small = ca.ODEModel(1.0, 2.0, 2, 2, 3, H)
    x44, ok = ca.hookstepsolve(small, 200.0, read_asc(known["m44"]["guess_file"]))
    xg = ca.changebasis(x44, small.ijkl, model.ijkl)
    x, ok = ca.hookstepsolve(model, 200.0, xg, ca.SearchParams(Nnewton=30))
    assert np.linalg.norm(model.f(x, 200.0)) / np.linalg.norm(x) < 1e-8
    out, succ = np.zeros(model.m), np.zeros(1, dtype=np.int32)
    Rv, cp = np.array([200.0]), _CParams(1e-8, 1e-8, 0.02, 20, 4, 6, 0)
    st = _solve(model._h, Rv.ctypes.data_as(C.POINTER(C.c_double)), xg.ctypes.data_as(C.POINTER(C.c_double)), 1,
                C.byref(cp), out.ctypes.data_as(C.POINTER(C.c_double)), succ.ctypes.data_as(C.POINTER(C.c_int32)), None)
    assert st == 5   # CA_ERR_UNSUPPORTED


def test_re_continuation_of_the_27d_lower_branch(ca, known, H):
    """continue-eqb.ipynb:654-668 without BifurcationKit: natural-parameter sweep and pseudo-arclength continuation
    of the 27d equilibrium (config 2 of BASELINE.json runs this loop)."""
    model = ca.ODEModel(1.0, 2.0, 1, 2, 3, H)
    x0, ok = ca.hookstepsolve(model, 200.0, read_asc(known["m27"]["guess_file"]))
    Rs = np.linspace(200.0, 300.0, 11)
    X, good, shear = ca.natural_continuation(model, x0, Rs, ca.SearchParams(Nnewton=30))
    assert good.all()
    for R, x in zip(Rs, X):
        assert np.linalg.norm(model.f(x, R)) < 1e-8
    assert abs(shear[0] / 1.646052703437098 - 1) < 1e-8       # continue-eqb.ipynb:179
    # pseudo-arclength towards lower Re: the branch turns back at a saddle-node (the notebook's plot spans
    # 150 < Re < 300) and returns to higher Re on the other side with a different shear
    br = ca.pseudo_arclength_continuation(model, x0, 200.0, ds=0.02, nsteps=400, direction=-1, R_max=260.0,
                                          ds_max=0.2, tol=1e-10)
    assert br["residual"].max() < 1e-8
    Rmin = br["R"].min()
    assert 100.0 < Rmin < 200.0
    assert br["R"][-1] > Rmin + 20.0                            # went around the fold
    k = np.argmin(br["R"])
    assert 0 < k < len(br["R"]) - 1
    after = br["R"][k:] > 199.0
    if after.any():                                             # the other branch at Re ~ 200 is a different state
        j = k + np.argmax(after)
        assert abs(br["shear"][j] - br["shear"][0]) > 1e-3


def test_newton_krylov_with_matrix_free_jacobian_actions(ca, known, H):
    """Newton-GMRES on device Jacobian actions (config 4's evaluation path) reaches the same equilibria as the
    hookstep search with the dense Jacobian."""
    small = ca.ODEModel(1.0, 2.0, 2, 2, 3, H)
    x44, ok = ca.hookstepsolve(small, 200.0, read_asc(known["m44"]["guess_file"]))
    xk, conv, info = ca.newton_krylov(small, 200.0, read_asc(known["m44"]["guess_file"]), ftol=1e-11)
    assert conv and np.linalg.norm(xk - x44) / np.linalg.norm(x44) < 1e-8
    model = ca.ODEModel(1.0, 2.0, 3, 3, 4, H)                      # 111 modes
    xg = ca.changebasis(x44, small.ijkl, model.ijkl)
    xh, _ = ca.hookstepsolve(model, 200.0, xg, ca.SearchParams(Nnewton=30))
    xk, conv, info = ca.newton_krylov(model, 200.0, xg, ftol=1e-11)
    assert conv, info
    assert np.linalg.norm(model.f(xk, 200.0)) < 1e-10
    assert np.linalg.norm(xk - xh) / np.linalg.norm(xh) < 1e-7
    assert info["jvps"] > 0
