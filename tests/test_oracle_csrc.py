"""The C restatement of the reference loops (the CPU baseline bench.py times) agrees with the Python oracle."""
import numpy as np

from oracle.cref import CRefModel
from oracle.model import ODEModel


def test_c_loops_match_python_oracle(nagata_H, known):
    ref = ODEModel(1.0, 2.0, 2, 2, 3, nagata_H)
    c = CRefModel(ref.B, ref.A1, ref.A2, ref.N.ijk, ref.N.val)
    rng = np.random.default_rng(0)
    X = 0.1 * rng.standard_normal((20, 44))
    for x in X[:5]:
        assert np.allclose(c.N(x), ref.N(x), rtol=0, atol=1e-16 * 44)
        assert np.allclose(c.derivative(x), ref.N.derivative(x), rtol=0, atol=1e-15)  # gcc contracts to FMA
        assert np.linalg.norm(c.f(x, 200.0) - ref.f(x, 200.0)) / np.linalg.norm(ref.f(x, 200.0)) < 1e-13
    a = c.f_ensemble(X, 200.0, threads=1)
    b = c.f_ensemble(X, 200.0, threads=4)
    assert np.array_equal(a, b)
    assert np.linalg.norm(a[3] - ref.f(X[3], 200.0)) / np.linalg.norm(a[3]) < 1e-13


def test_c_loops_known_equilibrium(nagata_H, known):
    ref = ODEModel(1.0, 2.0, 1, 1, 3, nagata_H)
    c = CRefModel(ref.B, ref.A1, ref.A2, ref.N.ijk, ref.N.val)
    x = np.array(known["m17_unnormalized"]["xeqb"])
    assert np.linalg.norm(c.f(x, 200.0)) / np.linalg.norm(x) < 1e-10
