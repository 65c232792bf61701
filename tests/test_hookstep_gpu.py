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


def test_blocked_lu_of_the_large_solver(ca):
    """The device LU (panel of 16 columns, row interchanges, block-row triangular solve, DGEMM trailing update) and
    the blocked substitutions against LAPACK: same pivots, factors and solutions."""
    import ctypes as C
    from scipy.linalg import lu_factor
    from cloudatlas_b200._lib import _sig, c_f64p, c_i32p, check
    fn = _sig("ca_selftest_lu_solve", c_f64p, C.c_int64, c_f64p, c_i32p)
    rng = np.random.default_rng(7)
    for m in (1, 5, 16, 17, 33, 100, 333, 1000):
        A = np.asfortranarray(rng.standard_normal((m, m)))
        b = rng.standard_normal(m)
        lu, piv = lu_factor(A)
        want = np.linalg.solve(A, b)
        Ad, bd, sing = A.copy(order="F"), b.copy(), C.c_int32(0)
        check(fn(Ad.ctypes.data_as(c_f64p), m, bd.ctypes.data_as(c_f64p), C.byref(sing)))
        assert sing.value == 0
        assert np.max(np.abs(Ad - lu)) / np.max(np.abs(lu)) < 1e-11 * max(1, m / 10)
        assert np.linalg.norm(bd - want) / np.linalg.norm(want) < 1e-13 * np.linalg.cond(A)
    Z = np.zeros((4, 4), order="F")
    sing = C.c_int32(0)
    check(fn(Z.ctypes.data_as(c_f64p), 4, np.zeros(4).ctypes.data_as(c_f64p), C.byref(sing)))
    assert sing.value == 1


def test_large_solver_follows_the_oracle_step_by_step(ca, known, model17n, nagata_H, monkeypatch):
    """The global-memory solver (forced onto the 17-mode model) makes the same decisions as Hookstep.jl's restatement:
    per-step residuals, trust radii and hookstep counts; and returns what the one-CTA solver returns."""
    from oracle.hookstep import SearchParams as RP, hookstepsolve as ref_solve
    from oracle.model import ODEModel as Ref
    ref = Ref(1.0, 2.0, 1, 1, 3, nagata_H, normalize=True)
    xg = np.array(known["m17_normalized"]["xguess"])
    for delta in (0.02, 0.01, 0.1, 0.003):
        rlog = []
        xr, okr = ref_solve(lambda x: ref.f(x, 200.0), lambda x: ref.Df(x, 200.0), xg, RP(delta=delta), log=rlog)
        xs, oks = ca.hookstepsolve(model17n, 200.0, xg, ca.SearchParams(delta=delta))
        monkeypatch.setenv("CA_HOOK_LARGE", "1")
        x, ok, log = ca.hookstepsolve(model17n, 200.0, xg, ca.SearchParams(delta=delta), return_log=True)
        monkeypatch.delenv("CA_HOOK_LARGE")
        assert ok == okr == oks
        assert np.linalg.norm(x - xr) / np.linalg.norm(xr) < 1e-9
        assert np.linalg.norm(x - xs) / np.linalg.norm(xs) < 1e-9
        assert len(rlog) == len(log["residual"])
        for (res, dl, nh), r2, d2, n2 in zip(rlog, log["residual"], log["delta"], log["hooksteps"]):
            assert abs(res - r2) <= 1e-6 * max(res, 1e-12) + 1e-13
            assert abs(dl - d2) <= 1e-6 * dl
            assert nh == n2
        assert log["device_seconds"] > 0


def test_larger_models_search_on_the_device(ca, known, H):
    """m > 95: ca_hookstep_solve_model runs the global-memory solver (no CA_ERR_UNSUPPORTED, no host linear algebra).
    The 44d equilibrium prolonged to 111 modes (per-entry Jacobian kernel) and to 307 modes converges to what the
    closure form of hookstepsolve (Hookstep.jl:109 on device closures, LAPACK solves) finds."""
    small = ca.ODEModel(1.0, 2.0, 2, 2, 3, H)
    x44, ok = ca.hookstepsolve(small, 200.0, read_asc(known["m44"]["guess_file"]))
    assert ok
    for JKL in ((3, 3, 4), (3, 3, 12)):
        model = ca.ODEModel(1.0, 2.0, *JKL, H)
        assert model.m > 95
        xg = ca.changebasis(x44, small.ijkl, model.ijkl)
        x, ok, log = ca.hookstepsolve(model, 200.0, xg, ca.SearchParams(Nnewton=30), return_log=True)
        assert ok, log
        assert np.linalg.norm(model.f(x, 200.0)) <= 1e-8
        xc, okc = ca.hookstepsolve(lambda v: model.f(v, 200.0), lambda v: model.Df(v, 200.0), xg, ca.SearchParams(Nnewton=30))
        assert okc and np.linalg.norm(x - xc) / np.linalg.norm(xc) < 1e-7
        # a sweep of searches through the same entry point (one after the other at this size)
        Rs = np.array([195.0, 205.0])
        X, good = ca.hookstepsolve(model, Rs, np.tile(x, (2, 1)), ca.SearchParams(Nnewton=30))
        assert good.all()


def test_large_solver_with_jacobian_by_ensemble_action(ca, H):
    """m > 400: Df comes from the Jacobian action on the unit vectors through the windowed DMMA ensemble kernel.  A
    587-mode model: the search started near an equilibrium of the model (the prolonged 44d state, converged with the
    closure form) returns to it."""
    small = ca.ODEModel(1.0, 2.0, 2, 2, 3, H)
    x44, ok = ca.hookstepsolve(small, 200.0, read_asc("xeq1projection-Re200-2-2-3-44d.asc"))
    model = ca.ODEModel(1.0, 2.0, 4, 4, 14, H)
    assert model.m > 400
    xg = ca.changebasis(x44, small.ijkl, model.ijkl)
    x, ok, log = ca.hookstepsolve(model, 200.0, xg, ca.SearchParams(Nnewton=40), return_log=True)
    assert ok, log
    assert np.linalg.norm(model.f(x, 200.0)) <= 1e-8
    D = model.Df(x, 200.0)                                         # per-entry kernel; the solver used the ensemble action
    v = np.random.default_rng(0).standard_normal(model.m)
    jv = model.Df_action(x, v, 200.0)
    assert np.linalg.norm(D @ v - jv) / np.linalg.norm(jv) < 1e-12


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
    assert conv and np.linalg.norm(small.f(xk, 200.0)) < 1e-10
    model = ca.ODEModel(1.0, 2.0, 3, 3, 4, H)                      # 111 modes
    xg = ca.changebasis(x44, small.ijkl, model.ijkl)
    xe, _ = ca.hookstepsolve(model, 200.0, xg, ca.SearchParams(Nnewton=30))
    xh, _ = ca.hookstepsolve(model, 205.0, xe, ca.SearchParams(Nnewton=30))    # one continuation step by both solvers
    xk, conv, info = ca.newton_krylov(model, 205.0, xe, ftol=1e-11)
    assert conv, info
    assert np.linalg.norm(model.f(xk, 205.0)) < 1e-10
    assert np.linalg.norm(xk - xh) / np.linalg.norm(xh) < 1e-7
    assert info["jvps"] > 0


def test_307_mode_lower_branch_follows_the_dns_curve(ca, H):
    """continue-eqb.ipynb:654-668,757-761 with a model the reference cannot build (307 modes): the wall shear I of the
    lower-branch Nagata equilibrium against Re lies on the DNS curve the reference ships (notebooks/data/eq1ReD-DNS.asc,
    computed with channelflow on a 32 x 49 x 40 grid) to better than 1 % for 200 <= Re <= 300 -- at Re = 200 to 1e-4.
    Every equilibrium comes from the device-resident global-memory Newton-hookstep (natural continuation in Re)."""
    import os
    from conftest import GOLDEN
    d = np.array([[float(t) for t in l.split("#")[0].split()] for l in open(os.path.join(GOLDEN, "eq1ReD-DNS.asc"))
                  if l.strip() and not l.startswith("#")])
    fold = int(np.argmin(d[:, 0]))
    lower = d[: fold + 1][::-1]                                   # lower branch, ascending in Re
    small = ca.ODEModel(1.0, 2.0, 2, 2, 3, H)
    x44, ok = ca.hookstepsolve(small, 200.0, read_asc("xeq1projection-Re200-2-2-3-44d.asc"))
    model = ca.ODEModel(1.0, 2.0, 3, 3, 12, H)
    x = ca.changebasis(x44, small.ijkl, model.ijkl)
    for R in (200.0, 250.0, 300.0):
        x, ok = ca.hookstepsolve(model, R, x, ca.SearchParams(Nnewton=40))
        assert ok
        I_dns = float(np.interp(R, lower[:, 0], lower[:, 1]))
        assert abs(model.shear(x) / I_dns - 1) < 0.01, (R, model.shear(x), I_dns)
        if R == 200.0:
            assert abs(model.shear(x) / I_dns - 1) < 1e-3
        assert abs(model.shear(x) - model.dissipation(x)) < 2e-3   # energy balance of an equilibrium, D = I, up to the truncation
