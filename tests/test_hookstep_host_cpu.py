"""The closure form of hookstepsolve in the host mirror (Hookstep.jl:95,109-352) is host logic: it is checked here,
without a device, against the oracle's restatement on the reference's own solver tests."""
import numpy as np

import __graft_entry__ as g
from oracle.hookstep import SearchParams as RefParams, hookstepsolve as ref_solve
from oracle.model import ODEModel

ca = g.load_package()


def test_2d_problem(known):
    k = known["hookstep2d"]                                    # test/runtests.jl:29-38
    f = lambda x: np.array([x[0] ** 2 + x[1] ** 2 - 1, x[0] - x[1]])
    Df = lambda x: np.array([[2 * x[0], 2 * x[1]], [1.0, -1.0]])
    x, ok = ca.hookstepsolve(f, Df, np.array(k["xguess"]))
    xr, okr = ref_solve(f, Df, np.array(k["xguess"]))
    assert ok == okr and np.allclose(x, xr, rtol=0, atol=1e-14)
    assert np.linalg.norm(x - np.array(k["xsolution"])) < k["tol"]
    x2, _ = ca.hookstepsolve(f, np.array(k["xguess"]))          # finite-difference Jacobian form
    assert np.linalg.norm(x2 - np.array(k["xsolution"])) < 1e-6


def test_17d_closures_follow_the_oracle(known, nagata_H):
    ref = ODEModel(1.0, 2.0, 1, 1, 3, nagata_H, normalize=True)
    f = lambda x: ref.f(x, 200.0)
    Df = lambda x: ref.Df(x, 200.0)
    xg = np.array(known["m17_normalized"]["xguess"])
    for delta in (0.02, 0.01, 0.2, 0.002):
        x, ok = ca.hookstepsolve(f, Df, xg, ca.SearchParams(delta=delta))
        xr, okr = ref_solve(f, Df, xg, RefParams(delta=delta))
        assert ok == okr
        assert np.linalg.norm(x - xr) <= 1e-12 * np.linalg.norm(xr)
    xe = np.array(known["m17_normalized"]["xeqb"])
    x, ok = ca.hookstepsolve(f, Df, xg, ca.SearchParams(delta=0.01))
    assert np.linalg.norm(x - xe) / np.linalg.norm(xe) < known["m17_normalized"]["solve_tol"]
