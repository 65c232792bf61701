"""cnab2 time stepping (ODEModels.jl:147-181, first "next" row of SURVEY 8f) on the device against the oracle."""
import numpy as np
import pytest

import __graft_entry__ as g

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ca():
    return g.load_package()


@pytest.fixture(scope="module")
def pair(ca, nagata_H):
    from oracle.model import ODEModel
    ref = ODEModel(1.0, 2.0, 1, 1, 3, nagata_H, normalize=True)
    return ca.ModelOperators(ref.B, ref.A1, ref.A2, ref.N.ijk, ref.N.val), ref


def test_single_trajectory_matches_oracle(ca, pair, known):
    from oracle.model import cnab2 as ref_cnab2
    dev, ref = pair
    x0 = np.array(known["m17_normalized"]["xguess"])
    for (dt, nsteps, nsave) in ((0.05, 40, 5), (0.1, 7, 2), (0.02, 3, 1)):
        t, X = ca.cnab2(x0, dev, 200.0, dt, nsteps, nsave)
        tr, Xr = ref_cnab2(x0, ref, 200.0, dt, nsteps, nsave)
        assert X.shape == Xr.shape and np.array_equal(t, tr)
        assert np.max(np.abs(X - Xr)) / np.max(np.abs(Xr)) < 1e-11


def test_equilibrium_is_a_fixed_point(ca, pair, known):
    dev, ref = pair
    xs = np.array(known["m17_normalized"]["xsoln_default_params"])
    t, X = ca.cnab2(xs, dev, 200.0, 0.05, 200, 50)
    assert np.max(np.abs(X - xs)) < 1e-9


def test_ensemble_of_trajectories(ca, pair, known):
    from oracle.model import cnab2 as ref_cnab2
    dev, ref = pair
    rng = np.random.default_rng(5)
    X0 = np.array(known["m17_normalized"]["xguess"]) + 0.01 * rng.standard_normal((200, 17))
    t, X = ca.cnab2(X0, dev, 200.0, 0.05, 20, 10)
    assert X.shape == (2, 200, 17)
    for s in (0, 17, 199):
        _, Xr = ref_cnab2(X0[s], ref, 200.0, 0.05, 20, 10)
        assert np.max(np.abs(X[:, s, :] - Xr)) / np.max(np.abs(Xr)) < 1e-11
    # few-state path (<= 16 states: streaming kernels) agrees with the ensemble path
    _, Xf = ca.cnab2(X0[:3], dev, 200.0, 0.05, 20, 10)
    assert np.max(np.abs(Xf - X[:, :3, :])) < 1e-13


def test_ensemble_with_windowed_operators(ca, nagata_H, known, monkeypatch):
    """Larger models pack N and the dense [M1 M2] operator (2m inputs) in windows; force that layout on the 44d model."""
    from oracle.model import ODEModel, cnab2 as ref_cnab2
    monkeypatch.setenv("CA_FORCE_WINDOW_MODES", "24")
    ref = ODEModel(1.0, 2.0, 2, 2, 3, nagata_H)
    dev = ca.ModelOperators(ref.B, ref.A1, ref.A2, ref.N.ijk, ref.N.val)
    rng = np.random.default_rng(9)
    X0 = 0.05 * rng.standard_normal((70, 44))
    t, X = ca.cnab2(X0, dev, 250.0, 0.02, 12, 4)
    assert X.shape == (3, 70, 44)
    for s in (0, 33, 69):
        _, Xr = ref_cnab2(X0[s], ref, 250.0, 0.02, 12, 4)
        assert np.max(np.abs(X[:, s, :] - Xr)) / np.max(np.abs(Xr)) < 1e-11


def test_dense_gemm_form_of_the_implicit_step(ca, nagata_H, monkeypatch):
    """Beyond 1024 modes the implicit operators M1 = LHS^-1 RHS, M2 = LHS^-1 (factored on the device per (R, dt)) are
    applied as two DGEMMs instead of a packed operator; forced here onto the 44-mode model, ensembles and few states."""
    from oracle.model import ODEModel, cnab2 as ref_cnab2
    monkeypatch.setenv("CA_CNAB2_DENSE", "1")
    ref = ODEModel(1.0, 2.0, 2, 2, 3, nagata_H)
    dev = ca.ModelOperators(ref.B, ref.A1, ref.A2, ref.N.ijk, ref.N.val)
    rng = np.random.default_rng(11)
    X0 = 0.05 * rng.standard_normal((70, 44))
    t, X = ca.cnab2(X0, dev, 250.0, 0.02, 12, 4)
    assert X.shape == (3, 70, 44)
    for s in (0, 33, 69):
        _, Xr = ref_cnab2(X0[s], ref, 250.0, 0.02, 12, 4)
        assert np.max(np.abs(X[:, s, :] - Xr)) / np.max(np.abs(Xr)) < 1e-11
    _, Xf = ca.cnab2(X0[:2], dev, 250.0, 0.02, 12, 4)
    assert np.max(np.abs(Xf - X[:, :2, :])) < 1e-13
