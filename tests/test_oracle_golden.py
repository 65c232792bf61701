"""Pins the CPU oracle against every known-answer value the reference ships for the hot path
(SURVEY.md section 8c): test/runtests.jl:29-107 and the stored notebook outputs."""
import numpy as np
import pytest

from conftest import read_asc
from oracle.basis import basisIndices, basisSet
from oracle.hookstep import SearchParams, hookstepsolve
from oracle.model import ODEModel


@pytest.fixture(scope="module")
def model17(nagata_H):
    return ODEModel(1.0, 2.0, 1, 1, 3, nagata_H)


@pytest.fixture(scope="module")
def model17n(nagata_H):
    return ODEModel(1.0, 2.0, 1, 1, 3, nagata_H, normalize=True)


def test_hookstep_2d(known):
    # test/runtests.jl:29-38
    k = known["hookstep2d"]
    f = lambda x: np.array([x[0] ** 2 + x[1] ** 2 - 1, x[0] - x[1]])
    Df = lambda x: np.array([[2 * x[0], 2 * x[1]], [1.0, -1.0]])
    x, ok = hookstepsolve(f, Df, np.array(k["xguess"]))
    assert np.linalg.norm(x - np.array(k["xsolution"])) < k["tol"]


def test_mode_counts(known, nagata_H):
    for key in ("m17_unnormalized", "m27", "m44"):
        J, K, L = known[key]["JKL"]
        assert basisIndices(J, K, L, nagata_H).shape[0] == known[key]["m"]
        if "fullset" in known[key]:
            assert basisIndices(J, K, L).shape[0] == known[key]["fullset"]


def test_17d_known_equilibrium(known, model17):
    # test/runtests.jl:40-59
    k = known["m17_unnormalized"]
    x = np.array(k["xeqb"])
    assert np.linalg.norm(model17.f(x, 200.0)) / np.linalg.norm(x) < k["residual_tol"]


def test_17d_hookstep(known, model17):
    # test/runtests.jl:61-82
    k = known["m17_unnormalized"]
    x, ok = hookstepsolve(lambda x: model17.f(x, 200.0), lambda x: model17.Df(x, 200.0),
                          np.array(k["xguess"]))
    xe = np.array(k["xeqb"])
    assert np.linalg.norm(x - xe) / np.linalg.norm(xe) < k["solve_tol"]


def test_17d_normalized_hookstep(known, model17n):
    # test/runtests.jl:84-107
    k = known["m17_normalized"]
    x, ok = hookstepsolve(lambda x: model17n.f(x, 200.0), lambda x: model17n.Df(x, 200.0),
                          np.array(k["xguess"]), SearchParams(delta=k["delta"]))
    xe = np.array(k["xeqb"])
    assert np.linalg.norm(x - xe) / np.linalg.norm(xe) < k["solve_tol"]


def test_17d_normalized_notebook_values(known, model17n):
    # find-eqb.ipynb:135,163-167,203,306-351,507-516
    k = known["m17_normalized"]
    assert abs(np.linalg.cond(model17n.B) / k["cond_B"] - 1) < 1e-12
    assert model17n.ijkl[:5].tolist() == k["ijkl_first5"]
    for n in range(4):
        comp = model17n.Psi[n][0]
        assert abs(comp.coeff / k["psi_coeffs_first5"][n] - 1) < 1e-14
        assert [float(c) for c in comp.p.c] == [float(c) for c in k["psi_polys_first5"][n]]
    v, w = model17n.Psi[4][1], model17n.Psi[4][2]
    assert abs(v.coeff / k["psi_coeffs_first5"][4][0] - 1) < 1e-14
    assert abs(w.coeff / k["psi_coeffs_first5"][4][1] - 1) < 1e-14
    assert [float(c) for c in v.p.c] == k["psi_polys_first5"][4][0]
    assert [float(c) for c in w.p.c] == k["psi_polys_first5"][4][1]
    assert np.linalg.norm(model17n.f(np.zeros(17), 400.0)) == k["norm_f_zero_R400"]
    # default SearchParams (the notebook builds hookparams but never passes it)
    x, ok = hookstepsolve(lambda x: model17n.f(x, 200.0), lambda x: model17n.Df(x, 200.0),
                          np.array(k["xguess"]))
    assert ok
    xs = np.array(k["xsoln_default_params"])
    assert np.linalg.norm(x - xs) / np.linalg.norm(xs) < 1e-10
    assert abs(np.linalg.norm(x) / k["norm_xsoln"] - 1) < 1e-10
    lam = np.linalg.eigvals(model17n.Df(xs, 200.0))
    lam = lam[np.argsort(-lam.real, kind="stable")]
    for (re, im) in k["eigs_top"]:
        assert np.min(np.abs(lam - complex(re, im))) < 1e-9


def test_27d_residual_of_shipped_guess(known, nagata_H):
    # continue-eqb.ipynb:115-120
    k = known["m27"]
    model = ODEModel(1.0, 2.0, *k["JKL"], nagata_H)
    xg = read_asc(k["guess_file"])
    assert len(xg) == 27
    val = np.linalg.norm(model.f(xg, 200.0))
    assert abs(val / k["norm_f_guess_R200"] - 1) < 1e-12
    x, ok = hookstepsolve(lambda x: model.f(x, 200.0), lambda x: model.Df(x, 200.0), xg)
    assert np.linalg.norm(model.f(x, 200.0)) / np.linalg.norm(x) < 1e-8
