"""Parity of the CUDA SparseBilinear path (through the C ABI) against the CPU oracle.
Tolerances: 1e-13 relative (norm-wise) for RHS / Jacobian outputs, as north_star states."""
import numpy as np
import pytest

import __graft_entry__ as g

pytestmark = pytest.mark.gpu
TOL = 1e-13


@pytest.fixture(scope="module")
def ca():
    return g.load_package()


def relerr(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


def model_coo(JKL, nagata_H):
    from oracle.basis import basisIndices
    from oracle.fastfloat import assemble_N_float
    ijkl = basisIndices(*JKL, nagata_H)
    ijk, val = assemble_N_float(1, 2, ijkl)
    return ijk, val, ijkl.shape[0]


@pytest.fixture(scope="module", params=[(1, 1, 3), (1, 2, 3), (2, 2, 3)], ids=["17d", "27d", "44d"])
def pair(request, ca, nagata_H):
    from oracle.bilinear import SparseBilinear as Ref
    ijk, val, m = model_coo(request.param, nagata_H)
    return ca.SparseBilinear(ijk, val, m), Ref(ijk, val, m)


def test_single_state_eval(pair):
    N, ref = pair
    rng = np.random.default_rng(1)
    x, y = 0.1 * rng.standard_normal(N.m), 0.1 * rng.standard_normal(N.m)
    assert relerr(N(x), ref(x)) < TOL
    assert relerr(N(x, y), ref(x, y)) < TOL
    assert relerr(N(y, x), ref(y, x)) < TOL


@pytest.mark.parametrize("nstates", [1, 7, 8, 31, 32, 33, 100, 10000])
def test_batched_eval_host(pair, nstates):
    from oracle.bilinear import eval_batched
    N, ref = pair
    X = 0.1 * np.random.default_rng(nstates).standard_normal((nstates, N.m))
    got = N(X)
    assert got.shape == X.shape
    assert relerr(got, eval_batched(ref, X)) < TOL


def test_batched_empty(pair):
    N, _ = pair
    assert N(np.zeros((0, N.m))).shape == (0, N.m)


def test_batched_eval_device_and_two_vector(pair, ca):
    import torch
    from oracle.bilinear import eval_batched
    N, ref = pair
    rng = np.random.default_rng(5)
    X = 0.1 * rng.standard_normal((999, N.m))
    Y = 0.1 * rng.standard_normal((999, N.m))
    dX = torch.from_numpy(np.asfortranarray(X).T.copy()).cuda().t()   # column-major on the device
    dY = torch.from_numpy(np.asfortranarray(Y).T.copy()).cuda().t()
    assert relerr(N(dX).cpu().numpy(), eval_batched(ref, X)) < TOL
    want = np.stack([ref(X[s], Y[s]) for s in range(0, 999, 37)])
    got = N(dX, dY).cpu().numpy()[::37]
    assert relerr(got, want) < TOL
    # leading dimension larger than nstates: a view of the first 500 states
    assert relerr(N(dX[:500]).cpu().numpy(), eval_batched(ref, X[:500])) < TOL


def test_jacobian(pair):
    N, ref = pair
    x = 0.1 * np.random.default_rng(2).standard_normal(N.m)
    DN = N.derivative(x)
    assert DN.shape == (N.m, N.m)
    assert relerr(DN, ref.derivative(x)) < TOL
    # DN(x) x = 2 N(x)   (Euler, degree-2 homogeneity)
    assert relerr(DN @ x, 2 * ref(x)) < 1e-12


@pytest.mark.parametrize("nstates", [1, 31, 77])
def test_jacobian_of_an_ensemble(pair, nstates):
    """derivative(N, X) for an ensemble: nstates x m x m, one reference derivative per state (host and device forms,
    ragged ensemble sizes, padded leading dimension); bit-identical to the single-state kernel's entries."""
    import torch
    N, ref = pair
    m = N.m
    X = 0.1 * np.random.default_rng(40 + nstates).standard_normal((nstates, m))
    want = np.stack([ref.derivative(x) for x in X])
    got = N.derivative(X)
    assert got.shape == (nstates, m, m) and got.flags.f_contiguous
    assert relerr(got, want) < TOL
    ld = nstates + 5
    buf = torch.zeros((m, ld), dtype=torch.float64, device="cuda")
    buf[:, :nstates] = torch.from_numpy(X.T)
    D = N.derivative(buf.t()[:nstates])
    assert tuple(D.shape) == (nstates, m, m) and D.stride() == (1, nstates, nstates * m)
    assert relerr(D.cpu().numpy(), want) < TOL
    assert relerr(D[0].cpu().numpy(), N.derivative(X[0])) < 1e-15
    # structurally zero entries are written as exact zeros (no memset needed by the caller)
    assert np.all(D.cpu().numpy()[:, np.all(want == 0.0, axis=0)] == 0.0)


def test_jacobian_of_an_ensemble_m307(ca, nagata_H):
    from oracle.bilinear import SparseBilinear as Ref
    ijk, val, m = model_coo((3, 3, 12), nagata_H)
    N, ref = ca.SparseBilinear(ijk, val, m), Ref(ijk, val, m)
    X = 0.1 * np.random.default_rng(12).standard_normal((37, m))
    got = N.derivative(X)
    for s in (0, 17, 36):
        assert relerr(got[s], ref.derivative(X[s])) < TOL


def test_jvp_matches_jacobian(pair, ca):
    import torch
    N, ref = pair
    rng = np.random.default_rng(3)
    X = 0.1 * rng.standard_normal((64, N.m))
    V = rng.standard_normal((64, N.m))
    dX = torch.from_numpy(X.T.copy()).cuda().t()
    dV = torch.from_numpy(V.T.copy()).cuda().t()
    got = N.jvp(dX, dV).cpu().numpy()
    want = np.stack([ref.derivative(X[s]) @ V[s] for s in range(64)])
    assert relerr(got, want) < TOL


def test_random_operator_with_duplicates_and_empty_rows(ca):
    from oracle.bilinear import SparseBilinear as Ref, eval_batched
    rng = np.random.default_rng(11)
    m, nnz = 53, 4000
    ijk = rng.integers(1, m + 1, size=(nnz, 3))
    ijk[:, 0] = np.minimum(ijk[:, 0], 40)          # rows 41..53 have no terms -> zeros
    val = rng.standard_normal(nnz)
    N, ref = ca.SparseBilinear(ijk, val, m), Ref(ijk, val, m)
    X = rng.standard_normal((257, m))
    got = N(X)
    assert relerr(got, eval_batched(ref, X)) < TOL
    assert np.all(got[:, 40:] == 0.0)
    x = X[0]
    assert relerr(N.derivative(x), ref.derivative(x)) < TOL
    assert relerr(N(x, X[1]), ref(x, X[1])) < TOL


def test_energy_conservation_17d(ca, nagata_H):
    """x . N(x,x) = 0 for the Galerkin-projected advection term (invariant iv)."""
    ijk, val, m = model_coo((1, 1, 3), nagata_H)
    from oracle.fastfloat import assemble_linear_float
    from oracle.basis import basisIndices
    B, _, _ = assemble_linear_float(1, 2, basisIndices(1, 1, 3, nagata_H))
    N = ca.SparseBilinear(ijk, val, m)
    X = np.random.default_rng(4).standard_normal((100, m))
    E = np.einsum("sm,sm->s", X, N(X))
    assert np.max(np.abs(E)) < 1e-12 * np.max(np.abs(N(X))) * np.max(np.abs(X)) * m


def test_m307_ensemble(ca, nagata_H):
    """Config 3 of BASELINE.json at parity-test size: 307-mode basis, oracle on a sample of states."""
    from oracle.bilinear import SparseBilinear as Ref, eval_batched
    ijk, val, m = model_coo((3, 3, 12), nagata_H)
    assert m == 307 and len(val) == 1355658
    N, ref = ca.SparseBilinear(ijk, val, m), Ref(ijk, val, m)
    X = 0.1 * np.random.default_rng(6).standard_normal((4099, m))
    got = N(X)
    idx = np.arange(0, 4099, 41)
    assert relerr(got[idx], eval_batched(ref, X[idx])) < TOL
    # size-independent properties on the full ensemble: homogeneity N(2x) = 4 N(x), and row independence
    assert relerr(N(2 * X), 4 * got) < 1e-15
    assert np.array_equal(N(X[1000:1100]), got[1000:1100])
    x = X[0]
    assert relerr(N.derivative(x), ref.derivative(x)) < TOL


def test_windowed_kernel_full_size_windows_m307(ca, nagata_H, monkeypatch):
    """Windows of the production size (128 modes) on the 307-mode operator: long windows are cut at the k-step cap of
    the shared-memory pair staging, tiles span several windows, the ensemble is ragged (not a multiple of 32)."""
    import torch
    from oracle.bilinear import SparseBilinear as Ref, eval_batched
    ijk, val, m = model_coo((3, 3, 12), nagata_H)
    ref = Ref(ijk, val, m)
    whole = ca.SparseBilinear(ijk, val, m)
    rng = np.random.default_rng(8)
    X = 0.1 * rng.standard_normal((2087, m))
    V = rng.standard_normal((2087, m))
    want_whole = whole(X)                                   # whole-tile kernel, same operator
    monkeypatch.setenv("CA_FORCE_WINDOW_MODES", "128")
    N = ca.SparseBilinear(ijk, val, m)
    got = N(X)
    idx = np.arange(0, 2087, 97)
    assert relerr(got[idx], eval_batched(ref, X[idx])) < TOL
    assert relerr(got, want_whole) < TOL
    dX = torch.from_numpy(X.T.copy()).cuda().t()
    dV = torch.from_numpy(V.T.copy()).cuda().t()
    sub = [0, 31, 32, 1055, 2080, 2086]
    assert relerr(N.jvp(dX, dV).cpu().numpy()[sub], np.stack([ref.derivative(X[s]) @ V[s] for s in sub])) < TOL
    assert relerr(N(dX, dV).cpu().numpy()[sub], np.stack([ref(X[s], V[s]) for s in sub])) < TOL


def test_windowed_kernel_many_blocks_repeatable(ca, nagata_H, monkeypatch):
    """The role-split windowed kernel over several blocks of 32 states per team (the window ring and the mbarrier
    phases run on across blocks; 20,011 states > 2 teams x 148 CTAs x 32): identical bits on repeated launches (a race
    between producer and consumer warps would show as run-to-run differences), agreement with the whole-tile kernel,
    both leading-dimension parities (16-byte and 8-byte gathers) and all three evaluation modes."""
    import torch
    ijk, val, m = model_coo((3, 3, 12), nagata_H)
    whole = ca.SparseBilinear(ijk, val, m)
    n = 20011
    rng = np.random.default_rng(77)
    for ld in (n + 1, n + 2):  # even and odd
        bx = torch.zeros((m, ld), dtype=torch.float64, device="cuda")
        bv = torch.zeros((m, ld), dtype=torch.float64, device="cuda")
        bx[:, :n] = torch.from_numpy(0.1 * rng.standard_normal((m, n)))
        bv[:, :n] = torch.from_numpy(rng.standard_normal((m, n)))
        dX, dV = bx.t()[:n], bv.t()[:n]
        monkeypatch.delenv("CA_FORCE_WINDOW_MODES", raising=False)
        want = [whole(dX), whole.jvp(dX, dV), whole(dX, dV)]
        monkeypatch.setenv("CA_FORCE_WINDOW_MODES", "128")
        N = ca.SparseBilinear(ijk, val, m)
        first = [N(dX), N.jvp(dX, dV), N(dX, dV)]
        for a, b in zip(first, want):
            assert relerr(a.cpu().numpy(), b.cpu().numpy()) < TOL
        for _ in range(3):
            again = [N(dX), N.jvp(dX, dV), N(dX, dV)]
            for a, b in zip(first, again):
                assert torch.equal(a, b)


@pytest.mark.parametrize("wm", [8, 24, 96])
def test_windowed_kernel_forced_at_small_m(ca, nagata_H, wm, monkeypatch):
    """The large-m (windowed, cp.async double-buffered) kernel, forced onto small operators."""
    import torch
    from oracle.bilinear import SparseBilinear as Ref, eval_batched
    monkeypatch.setenv("CA_FORCE_WINDOW_MODES", str(wm))
    ijk, val, m = model_coo((2, 2, 3), nagata_H)
    N, ref = ca.SparseBilinear(ijk, val, m), Ref(ijk, val, m)
    rng = np.random.default_rng(21)
    X = 0.1 * rng.standard_normal((1003, m))
    Y = rng.standard_normal((1003, m))
    assert relerr(N(X), eval_batched(ref, X)) < TOL
    dX = torch.from_numpy(X.T.copy()).cuda().t()
    dY = torch.from_numpy(Y.T.copy()).cuda().t()
    idx = list(range(0, 1003, 59))
    assert relerr(N(dX, dY).cpu().numpy()[idx], np.stack([ref(X[s], Y[s]) for s in idx])) < TOL
    assert relerr(N.jvp(dX, dY).cpu().numpy()[idx], np.stack([ref.derivative(X[s]) @ Y[s] for s in idx])) < TOL
    # random operator with empty rows and duplicates
    m2, nnz = 53, 4000
    ijk2 = rng.integers(1, m2 + 1, size=(nnz, 3))
    ijk2[:, 0] = np.minimum(ijk2[:, 0], 40)
    v2 = rng.standard_normal(nnz)
    N2, ref2 = ca.SparseBilinear(ijk2, v2, m2), Ref(ijk2, v2, m2)
    X2 = rng.standard_normal((300, m2))
    got = N2(X2)
    assert relerr(got, eval_batched(ref2, X2)) < TOL
    assert np.all(got[:, 40:] == 0.0)


def test_windowed_model_rhs(ca, nagata_H, monkeypatch):
    from oracle.model import ODEModel
    monkeypatch.setenv("CA_FORCE_WINDOW_MODES", "32")
    ref = ODEModel(1.0, 2.0, 2, 2, 3, nagata_H)
    dev = ca.ModelOperators(ref.B, ref.A1, ref.A2, ref.N.ijk, ref.N.val)
    X = 0.1 * np.random.default_rng(3).standard_normal((700, 44))     # < 4096 states: DMMA path, not NVRTC
    for R in (200.0, 300.0):
        got = dev.f(X, R)
        want = np.stack([ref.f(x, R) for x in X[:30]])
        assert relerr(got[:30], want) < 1e-13 * max(1.0, np.linalg.cond(ref.B) / 10)


@pytest.mark.parametrize("seed", range(4))
def test_fuzz_all_kernel_paths_agree(ca, seed, monkeypatch):
    """Random operators (duplicates, empty rows, dense-ish and very sparse), ensembles on both sides of the
    streaming / DMMA threshold, padded leading dimensions: every path against the oracle loop."""
    import torch
    from oracle.bilinear import SparseBilinear as Ref, eval_batched
    rng = np.random.default_rng(100 + seed)
    m = int(rng.integers(2, 120))
    nnz = int(rng.integers(1, 6 * m * m))
    ijk = rng.integers(1, m + 1, size=(nnz, 3))
    if seed % 2:
        ijk[:, 0] = rng.integers(1, max(2, m // 3) + 1, size=nnz)
    val = rng.standard_normal(nnz)
    if seed == 3:
        monkeypatch.setenv("CA_FORCE_WINDOW_MODES", "16")
    N, ref = ca.SparseBilinear(ijk, val, m), Ref(ijk, val, m)
    for nstates in (1, 2, 5, 16, 17, 40, 301):
        X = rng.standard_normal((nstates, m))
        Y = rng.standard_normal((nstates, m))
        want = eval_batched(ref, X)
        assert relerr(N(X), want) < TOL, (m, nnz, nstates)
        ld = nstates + 3                                              # padded leading dimension on the device
        buf = torch.zeros((m, ld), dtype=torch.float64, device="cuda")
        buf[:, :nstates] = torch.from_numpy(X.T)
        dX = buf.t()[:nstates]
        assert relerr(N(dX).cpu().numpy(), want) < TOL
        bufy = torch.zeros((m, ld), dtype=torch.float64, device="cuda")
        bufy[:, :nstates] = torch.from_numpy(Y.T)
        dY = bufy.t()[:nstates]
        want2 = np.stack([ref(X[s], Y[s]) for s in range(min(nstates, 6))])
        assert relerr(N(dX, dY).cpu().numpy()[:6], want2) < TOL
        wantj = np.stack([ref.derivative(X[s]) @ Y[s] for s in range(min(nstates, 6))])
        assert relerr(N.jvp(dX, dY).cpu().numpy()[:6], wantj) < TOL
    x = rng.standard_normal(m)
    assert relerr(N.derivative(x), ref.derivative(x)) < TOL


def test_operator_without_terms(ca):
    N = ca.SparseBilinear(np.zeros((0, 3), dtype=np.int64), np.zeros(0), 5)
    X = np.random.default_rng(0).standard_normal((40, 5))
    assert np.all(N(X) == 0.0) and np.all(N(X[0]) == 0.0) and np.all(N.derivative(X[0]) == 0.0)
