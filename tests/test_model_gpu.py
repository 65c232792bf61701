"""Parity of the folded model RHS / Jacobian (ODEModels.jl:57-58) against the oracle's closures."""
import numpy as np
import pytest

import __graft_entry__ as g
from conftest import read_asc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ca():
    return g.load_package()


def relerr(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope="module", params=[((1, 1, 3), False), ((1, 1, 3), True), ((1, 2, 3), False), ((2, 2, 3), False)],
                ids=["17d", "17d-normalized", "27d", "44d"])
def pair(request, ca, nagata_H):
    from oracle.model import ODEModel
    JKL, normalize = request.param
    ref = ODEModel(1.0, 2.0, *JKL, nagata_H, normalize=normalize)
    dev = ca.ModelOperators(ref.B, ref.A1, ref.A2, ref.N.ijk, ref.N.val)
    return dev, ref


def test_rhs_single_and_batched(pair):
    dev, ref = pair
    rng = np.random.default_rng(0)
    X = 0.1 * rng.standard_normal((1000, dev.m))
    for R in (200.0, 400.0, 200.0):
        want = np.stack([ref.f(x, R) for x in X[:50]])
        # f = B^-1(...): forward error scales with cond(B); the bound is 1e-13 * max(1, cond/10)
        tol = 1e-13 * max(1.0, np.linalg.cond(ref.B) / 10)
        assert relerr(dev.f(X[0], R), want[0]) < tol
        got = dev.f(X, R)
        assert relerr(got[:50], want) < tol
        # backward parity, independent of conditioning: B f = A1 x + A2 x / R + N(x)
        rhs = np.stack([ref.A1 @ x + (ref.A2 @ x) / R + ref.N(x) for x in X[:50]])
        assert relerr(got[:50] @ ref.B.T, rhs) < 1e-13


def test_jacobian(pair):
    dev, ref = pair
    x = 0.1 * np.random.default_rng(1).standard_normal(dev.m)
    tol = 1e-13 * max(1.0, np.linalg.cond(ref.B) / 10)
    assert relerr(dev.Df(x, 200.0), ref.Df(x, 200.0)) < tol
    # finite-difference check (Hookstep.jl:77-93)
    from oracle.hookstep import Df_finitediff
    fd = Df_finitediff(lambda z: dev.f(z, 200.0), x)
    assert relerr(fd, dev.Df(x, 200.0)) < 1e-5


def test_jacobian_of_an_ensemble(pair):
    """Df(X, R) for an ensemble (nstates x m x m), host and device forms, against the reference closure per state."""
    import torch
    dev, ref = pair
    X = 0.1 * np.random.default_rng(9).standard_normal((45, dev.m))
    tol = 1e-13 * max(1.0, np.linalg.cond(ref.B) / 10)
    for R in (200.0, 320.0):
        want = np.stack([ref.Df(x, R) for x in X])
        got = dev.Df(X, R)
        assert got.shape == (45, dev.m, dev.m)
        assert relerr(got, want) < tol
        dX = torch.from_numpy(X.T.copy()).cuda().t()
        assert relerr(dev.Df(dX, R).cpu().numpy(), want) < tol
    assert relerr(dev.Df(X, 200.0)[3], dev.Df(X[3], 200.0)) < 1e-15


def test_jvp(pair, ca):
    import torch
    dev, ref = pair
    rng = np.random.default_rng(2)
    X = 0.1 * rng.standard_normal((40, dev.m))
    V = rng.standard_normal((40, dev.m))
    dX = torch.from_numpy(X.T.copy()).cuda().t()
    dV = torch.from_numpy(V.T.copy()).cuda().t()
    got = dev.Df_action(dX, dV, 250.0).cpu().numpy()
    want = np.stack([ref.Df(X[s], 250.0) @ V[s] for s in range(40)])
    assert relerr(got, want) < 1e-13 * max(1.0, np.linalg.cond(ref.B) / 10)


def test_known_equilibrium_17d(ca, known, nagata_H):
    """test/runtests.jl:40-59 through the CUDA path."""
    from oracle.model import ODEModel
    ref = ODEModel(1.0, 2.0, 1, 1, 3, nagata_H)
    dev = ca.ModelOperators(ref.B, ref.A1, ref.A2, ref.N.ijk, ref.N.val)
    x = np.array(known["m17_unnormalized"]["xeqb"])
    assert np.linalg.norm(dev.f(x, 200.0)) / np.linalg.norm(x) < 1e-10


def test_27d_residual_of_shipped_guess(ca, known, nagata_H):
    """continue-eqb.ipynb:115-120 through the CUDA path."""
    from oracle.model import ODEModel
    ref = ODEModel(1.0, 2.0, 1, 2, 3, nagata_H)
    dev = ca.ModelOperators(ref.B, ref.A1, ref.A2, ref.N.ijk, ref.N.val)
    xg = read_asc(known["m27"]["guess_file"])
    assert abs(np.linalg.norm(dev.f(xg, 200)) / known["m27"]["norm_f_guess_R200"] - 1) < 1e-12


def test_small_model_large_ensemble_uses_specialised_kernel(ca, nagata_H):
    """m <= 48 and >= 4096 states: the NVRTC straight-line kernel; same parity bar, and R re-targeting works."""
    from oracle.model import ODEModel
    for JKL in ((1, 1, 3), (2, 2, 3)):
        ref = ODEModel(1.0, 2.0, *JKL, nagata_H)
        dev = ca.ModelOperators(ref.B, ref.A1, ref.A2, ref.N.ijk, ref.N.val)
        rng = np.random.default_rng(7)
        X = 0.1 * rng.standard_normal((5000, dev.m))
        small = dev.f(X[:100], 200.0)                      # DMMA path
        for R in (200.0, 350.0):
            got = dev.f(X, R)
            path, src, log = dev.kernel_path(5000)
            assert path == "nvrtc_straight_line", log
            assert "ca_small" in src
            want = np.stack([ref.f(x, R) for x in X[:40]])
            assert relerr(got[:40], want) < 1e-13 * max(1.0, np.linalg.cond(ref.B) / 10)
        assert relerr(dev.f(X, 200.0)[:100], small) < 1e-14


def test_specialised_kernel_padded_leading_dimension_and_ragged_sizes(ca, nagata_H):
    """The generated kernel's per-thread cp.async ring with a padded leading dimension on both sides and ensemble
    sizes that are not multiples of the CTA width: same numbers as the contiguous call, nothing written past the end."""
    import torch
    from oracle.model import ODEModel
    ref = ODEModel(1.0, 2.0, 1, 2, 3, nagata_H)
    dev = ca.ModelOperators(ref.B, ref.A1, ref.A2, ref.N.ijk, ref.N.val)
    m = dev.m
    rng = np.random.default_rng(17)
    for n in (4096, 4097, 9999):
        X = 0.1 * rng.standard_normal((n, m))
        ld = n + 7
        xin = torch.zeros((m, ld), dtype=torch.float64, device="cuda")
        xin[:, :n] = torch.from_numpy(X.T)
        out = torch.full((m, ld), 123.0, dtype=torch.float64, device="cuda")
        got = dev.f(xin.t()[:n], 200.0, out=out.t()[:n])
        assert dev.kernel_path(n)[0] == "nvrtc_straight_line"
        want = dev.f(X, 200.0)
        assert np.array_equal(got.cpu().numpy(), want)
        assert torch.all(out[:, n:] == 123.0)
        sample = [0, n // 2, n - 1]
        assert relerr(want[sample], np.stack([ref.f(X[s], 200.0) for s in sample])) < 1e-13 * max(1.0, np.linalg.cond(ref.B) / 10)


def test_three_n_tile_windowed_model_matches_whole_tile_model(ca, monkeypatch):
    """The folded 307-mode model through the windowed kernel with tiles of three n-tiles (forced: the packer would pick
    them only for ~3000-mode models) against the same model through the whole-tile kernel: f, the Jacobian action and
    re-targeting R (the linear slots sit in three-n-tile fragments) agree to rounding."""
    import torch
    sx, sy, sz, tx, tz = ca.halfbox_symmetries()
    H = [sx * sy * sz, sz * tx * tz]
    base = ca.ODEModel(1.0, 2.0, 3, 3, 12, H)
    monkeypatch.setenv("CA_FORCE_WINDOW_MODES", "128")
    monkeypatch.setenv("CA_NT3", "1")
    wide = ca.ODEModel(1.0, 2.0, 3, 3, 12, H)
    assert wide.kernel_path(1000)[0] == "dmma_row_tiles_windowed"
    assert wide.info()["padded_fma_per_state"] != base.info()["padded_fma_per_state"]
    rng = np.random.default_rng(23)
    n = 1000 + 13
    X = torch.from_numpy((0.1 * rng.standard_normal((base.m, n)))).cuda().t()
    V = torch.from_numpy(rng.standard_normal((base.m, n))).cuda().t()
    for R in (200.0, 417.0):
        a, b = base.f(X, R).cpu().numpy(), wide.f(X, R).cpu().numpy()
        assert relerr(b, a) < 1e-13
        a, b = base.Df_action(X, V, R).cpu().numpy(), wide.Df_action(X, V, R).cpu().numpy()
        assert relerr(b, a) < 1e-13


def test_full_size_ensemble_properties(ca, nagata_H):
    """BASELINE.json configs[2] at full size (10^6 states, 307 modes): size-independent properties -- results do not
    depend on where a state sits in the ensemble (bit-exact under permutation), the host-pointer path equals the
    device path bit for bit, and a random sample agrees with the CPU port of the reference loops."""
    import torch
    from oracle.cref import CRefModel
    sx, sy, sz, tx, tz = ca.halfbox_symmetries()
    model = ca.ODEModel(1.0, 2.0, 3, 3, 12, [sx * sy * sz, sz * tx * tz])
    n, m = 1_000_000, model.m
    gen = torch.Generator(device="cuda").manual_seed(20251029)
    X = (0.1 * torch.randn((m, n), dtype=torch.float64, device="cuda", generator=gen)).t()
    F = model.f(X, 200.0)
    perm = torch.randperm(n, device="cuda", generator=gen)
    Xp = X.t()[:, perm].contiguous().t()
    Fp = model.f(Xp, 200.0)
    assert torch.equal(Fp, F.t()[:, perm].t())
    # host path on a slice (pageable memory): same bits
    sl = slice(123_456, 323_456)
    Fh = model.f(np.asfortranarray(X[sl].cpu().numpy()), 200.0)
    assert np.array_equal(Fh, F[sl].cpu().numpy())
    # reference loops on a sample
    idx = torch.randint(0, n, (96,), device="cuda", generator=gen)
    ref = CRefModel(model.B, model.A1, model.A2, model.ijk, model.val)
    want = ref.f_ensemble(np.asfortranarray(X[idx].cpu().numpy()), 200.0, threads=8)
    got = F[idx].cpu().numpy()
    tol = 1e-13 * max(1.0, np.linalg.cond(model.B) / 10)
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < tol
    # backward parity independent of cond(B): B f = A1 x + A2 x / R + N(x)
    xs = X[idx].cpu().numpy()
    rhs = np.stack([model.A1 @ x + (model.A2 @ x) / 200.0 + ref.N(x) for x in xs])
    assert np.linalg.norm(got @ model.B.T - rhs) / np.linalg.norm(rhs) < 1e-13


def test_host_pointer_forms_of_jvp_cnab2_observables(ca, nagata_H):
    """The host-array entry points a Julia binding uses (no device pointers) agree with the device-pointer forms."""
    import ctypes as C
    import torch
    from cloudatlas_b200.model import _cnab2_host, _obs_host
    from cloudatlas_b200._lib import c_f64p, check
    sx, sy, sz, tx, tz = ca.halfbox_symmetries()
    model = ca.ODEModel(1.0, 2.0, 2, 2, 3, [sx * sy * sz, sz * tx * tz])
    rng = np.random.default_rng(2)
    X = np.asfortranarray(0.1 * rng.standard_normal((50, 44)))
    V = np.asfortranarray(rng.standard_normal((50, 44)))
    dX, dV = (torch.from_numpy(A.T.copy()).cuda().t() for A in (X, V))
    assert np.array_equal(model.Df_action(X, V, 210.0), model.Df_action(dX, dV, 210.0).cpu().numpy())
    t, Xs = ca.cnab2(X, model, 210.0, 0.05, 6, 3)
    save = np.empty((2, 44, 50))
    final = np.empty((50, 44), order="F")
    check(_cnab2_host(model._h, X.ctypes.data_as(c_f64p), 50, 210.0, 0.05, 6, 3, save.ctypes.data_as(c_f64p),
                      final.ctypes.data_as(c_f64p)))
    assert np.array_equal(save.transpose(0, 2, 1), Xs) and np.array_equal(final, Xs[-1])
    obs = np.empty((50, 2), order="F")
    check(_obs_host(model._h, X.ctypes.data_as(c_f64p), 50, obs.ctypes.data_as(c_f64p)))
    assert np.array_equal(obs, model.observables(X))
