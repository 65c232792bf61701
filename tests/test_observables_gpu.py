"""shear_function / dissipation_function (ODEModels.jl:65-137) on the device against the oracle,
and the notebook's stored shear of the 27d equilibrium (continue-eqb.ipynb:179)."""
import numpy as np
import pytest

import __graft_entry__ as g
from conftest import read_asc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ca():
    return g.load_package()


def lib_H(ca):
    sx, sy, sz, tx, tz = ca.halfbox_symmetries()
    return [sx * sy * sz, sz * tx * tz]


@pytest.mark.parametrize("normalize", [False, True])
def test_shear_vector_and_dissipation_matrix(ca, nagata_H, normalize):
    from oracle.model import ODEModel, dissipation_matrix, shear_vector
    ref = ODEModel(1.0, 2.0, 2, 2, 3, nagata_H, normalize=normalize)
    dev = ca.ODEModel(1.0, 2.0, 2, 2, 3, lib_H(ca), normalize=normalize)
    s_ref, D_ref = shear_vector(ref.Psi), dissipation_matrix(ref.Psi)
    assert np.max(np.abs(dev.Psishear - s_ref)) <= 1e-12 * np.max(np.abs(s_ref))
    assert np.count_nonzero(s_ref) > 0
    assert np.max(np.abs(dev.D - D_ref)) <= 1e-12 * np.max(np.abs(D_ref))
    assert np.allclose(dev.D, dev.D.T, rtol=0, atol=1e-13 * np.max(np.abs(D_ref)))
    rng = np.random.default_rng(0)
    X = 0.1 * rng.standard_normal((300, 44))
    obs = dev.observables(X)
    want_s = 1 + X @ s_ref
    want_d = 1 + np.einsum("si,ij,sj->s", X, D_ref, X)
    assert np.max(np.abs(obs[:, 0] - want_s)) < 1e-13 * np.max(np.abs(want_s))
    assert np.max(np.abs(obs[:, 1] - want_d)) < 1e-12 * np.max(np.abs(want_d))
    s1, d1 = dev.observables(X[0])
    assert abs(s1 - want_s[0]) < 1e-13 and abs(d1 - want_d[0]) < 1e-12 * abs(want_d[0])


def test_notebook_shear_of_the_27d_equilibrium(ca, known):
    """continue-eqb.ipynb:142-179: hookstep from the shipped projection, then shear(xeq1, model) = 1.646052703437098."""
    model = ca.ODEModel(1.0, 2.0, 1, 2, 3, lib_H(ca))
    xg = read_asc(known["m27"]["guess_file"])
    x, ok = ca.hookstepsolve(model, 200.0, xg)
    assert abs(model.shear(x) / 1.646052703437098 - 1) < 1e-8
    assert model.dissipation(x) > 1.0
