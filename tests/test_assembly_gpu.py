"""Parity of the CUDA Galerkin assembly (ODEModels.jl:31-51) against the oracle.
L and Q within 1e-12 relative (max-abs difference over max-abs reference); sparsity pattern, order and
index sets bit-exact (north_star)."""
import numpy as np
import pytest

import __graft_entry__ as g

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def ca():
    return g.load_package()


def maxrel(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(b))


def lib_H(ca):
    sx, sy, sz, tx, tz = ca.halfbox_symmetries()
    return [sx * sy * sz, sz * tx * tz]


@pytest.mark.parametrize("JKL", [(1, 1, 3), (1, 2, 3), (2, 2, 3)], ids=["17d", "27d", "44d"])
@pytest.mark.parametrize("normalize", [False, True])
def test_small_models_against_float_restatement_and_exact(ca, nagata_H, JKL, normalize):
    from oracle.exact import assemble_N_exact
    from oracle.model import ODEModel as Ref
    ref = Ref(1.0, 2.0, *JKL, nagata_H, normalize=normalize)       # the reference's Float64 algorithm
    dev = ca.ODEModel(1.0, 2.0, *JKL, lib_H(ca), normalize=normalize)
    assert np.array_equal(dev.ijkl, ref.ijkl)
    for name in ("B", "A1", "A2"):
        assert maxrel(getattr(dev, name), getattr(ref, name)) < TOL, name
    # structural zeros of B (Fourier orthogonality, parity: BasisFunctions.jl:268) are exact zeros here too;
    # where only the polynomial integral vanishes the reference holds rounding noise
    assert np.all(np.abs(dev.B[ref.B == 0]) < 1e-14)
    assert np.all(np.abs(ref.B[dev.B == 0]) < 1e-14)
    diffmode = (np.abs(ref.ijkl[:, None, 1]) != np.abs(ref.ijkl[None, :, 1])) | \
               (np.abs(ref.ijkl[:, None, 2]) != np.abs(ref.ijkl[None, :, 2]))
    assert np.all(dev.B[diffmode] == 0)          # different wavenumbers: exactly orthogonal
    if not normalize:
        ijk, val = assemble_N_exact(1, 2, ref.ijkl)                  # exact pattern (396 / 1500 / 5654)
        assert np.array_equal(dev.ijk, ijk)
        assert maxrel(dev.val, np.array([float(v) for v in val])) < TOL
    # against the Float64 restatement: same values on the common pattern, extras of the restatement are noise
    got = {tuple(r): v for r, v in zip(dev.ijk.tolist(), dev.val)}
    scale = np.max(np.abs(ref.N.val))
    for r, v in zip(ref.N.ijk.tolist(), ref.N.val):
        assert abs(got.get(tuple(r), 0.0) - v) < TOL * scale
    assert len(got) <= ref.N.nnz


@pytest.mark.parametrize("case", ["no-symmetry", "one-generator", "nagata-wide-z", "nagata-alpha-gamma"])
def test_candidate_pruning_changes_nothing(ca, case, monkeypatch):
    """The assembly scans only the third factors of the wavenumber groups that the Fourier selection rules allow.  On
    bases outside the Nagata subspace too (no symmetry at all, a single generator, unequal J and K, other alpha and
    gamma) the result must be bit-identical to scanning every k, and agree with the float restatement of the reference."""
    from oracle.basis import basisIndices as ref_indices
    from oracle.fastfloat import assemble_N_float
    sx, sy, sz, tx, tz = ca.halfbox_symmetries()
    alpha, gamma = (1.0, 2.0)
    if case == "no-symmetry":
        ijkl = ca.basisIndices(1, 1, 2)
    elif case == "one-generator":
        ijkl = ca.basisIndices(2, 1, 3, [sx * sy * sz])
    elif case == "nagata-wide-z":
        ijkl = ca.basisIndices(1, 3, 4, lib_H(ca))
    else:
        ijkl, alpha, gamma = ca.basisIndices(2, 2, 4, lib_H(ca)), 0.75, 1.625
    slab = ca.AssemblySlab(alpha, gamma, ijkl)
    ijk, val = slab.coo_host()
    B, A1, A2 = slab.linear_host()
    monkeypatch.setenv("CA_ASSEMBLE_NOPRUNE", "1")
    full = ca.AssemblySlab(alpha, gamma, ijkl)
    ijk_f, val_f = full.coo_host()
    assert np.array_equal(ijk, ijk_f) and np.array_equal(val, val_f)
    for X, Y in zip((B, A1, A2), full.linear_host()):
        assert np.array_equal(X, Y)
    rijk, rval = assemble_N_float(alpha, gamma, np.asarray(ijkl))
    assert np.array_equal(ijk, rijk)
    assert maxrel(val, rval) < TOL


def test_nnz_of_the_baseline_configs(ca):
    H = lib_H(ca)
    for JKL, m, nnz in [((1, 1, 3), 17, 396), ((1, 2, 3), 27, 1500), ((2, 2, 3), 44, 5654)]:
        slab = ca.AssemblySlab(1.0, 2.0, ca.basisIndices(*JKL, H))
        assert (slab.m, slab.nnz) == (m, nnz)


def test_m307_against_exact_rational_rows(ca, nagata_H):
    """L = 12: the reference's Float64 is off by 4e-2 here (SURVEY exec. summary); truth is the
    exact-rational tier, checked on a sample of output rows, the vectorised tier on all of them."""
    from oracle.exact import assemble_N_exact, assemble_linear_exact
    from oracle.fastfloat import assemble_N_float
    ijkl = ca.basisIndices(3, 3, 12, lib_H(ca))
    slab = ca.AssemblySlab(1.0, 2.0, ijkl)
    ijk, val = slab.coo_host()
    assert slab.nnz == 1355658
    fijk, fval = assemble_N_float(1, 2, ijkl)
    assert np.array_equal(ijk, fijk)
    assert maxrel(val, fval) < TOL
    rows = [0, 5, 77, 150, 233, 306]
    eijk, eval_ = assemble_N_exact(1, 2, ijkl, rows=rows)
    sel = np.isin(ijk[:, 0] - 1, rows)
    assert np.array_equal(ijk[sel], eijk)
    ev = np.array([float(v) for v in eval_])
    assert np.max(np.abs(val[sel] - ev)) / np.max(np.abs(ev)) < TOL
    B, A1, A2 = slab.linear_host()
    Be, A1e, A2e = assemble_linear_exact(1, 2, ijkl[:60])  # nested basis: leading block of the operators
    conv = np.vectorize(float, otypes=[np.float64])
    assert maxrel(B[:60, :60], conv(Be)) < TOL
    assert maxrel(A1[:60, :60], conv(A1e)) < TOL
    assert maxrel(A2[:60, :60], conv(A2e)) < TOL


def test_row_slabs_concatenate_to_the_full_operator(ca):
    """Sharding by output mode: emulated ranks on one GPU."""
    ijkl = ca.basisIndices(2, 2, 3, lib_H(ca))
    full = ca.AssemblySlab(1.0, 2.0, ijkl)
    fB, fA1, fA2 = full.linear_host()
    fijk, fval = full.coo_host()
    for world in (2, 3, 8):
        bounds = ca.row_partition(44, world)
        slabs = [ca.AssemblySlab(1.0, 2.0, ijkl, False, bounds[r], bounds[r + 1]) for r in range(world)]
        ijk = np.concatenate([s.coo_host()[0] for s in slabs])
        val = np.concatenate([s.coo_host()[1] for s in slabs])
        assert np.array_equal(ijk, fijk) and np.array_equal(val, fval)
        B = sum(s.linear_host()[0] for s in slabs)
        assert np.array_equal(B, fB)
    empty = ca.AssemblySlab(1.0, 2.0, ijkl, False, 10, 10)
    assert empty.nnz == 0


def test_scaling_law_of_a_rescaled_basis(ca):
    """Psi_n -> s_n Psi_n gives N'_ijk = s_i s_j s_k N_ijk; normalisation is such a rescaling."""
    H = lib_H(ca)
    ijkl = ca.basisIndices(2, 2, 3, H)
    a = ca.AssemblySlab(1.0, 2.0, ijkl, False)
    b = ca.AssemblySlab(1.0, 2.0, ijkl, True)
    s = 1 / np.sqrt(np.diag(a.linear_host()[0]))
    (ijk, va), (ijk2, vb) = a.coo_host(), b.coo_host()
    assert np.array_equal(ijk, ijk2)
    want = va * s[ijk[:, 0] - 1] * s[ijk[:, 1] - 1] * s[ijk[:, 2] - 1]
    assert maxrel(vb, want) < 1e-13
    assert np.allclose(np.diag(b.linear_host()[0]), 1.0, atol=1e-14)


def test_full_model_17d_known_answers(ca, known):
    """test/runtests.jl:40-59 and find-eqb.ipynb:135,203 end to end on the device."""
    H = lib_H(ca)
    model = ca.ODEModel(1.0, 2.0, 1, 1, 3, H)
    assert len(model) == 17
    x = np.array(known["m17_unnormalized"]["xeqb"])
    assert np.linalg.norm(model.f(x, 200.0)) / np.linalg.norm(x) < 1e-10
    modeln = ca.ODEModel(1.0, 2.0, 1, 1, 3, H, normalize=True)
    assert abs(np.linalg.cond(modeln.B) / known["m17_normalized"]["cond_B"] - 1) < 1e-12
    assert np.linalg.norm(modeln.f(np.zeros(17), 400.0)) == 0.0
    xs = np.array(known["m17_normalized"]["xsoln_default_params"])
    assert np.linalg.norm(modeln.f(xs, 200.0)) / np.linalg.norm(xs) < 1e-11
    lam = np.linalg.eigvals(modeln.Df(xs, 200.0))
    for re, im in known["m17_normalized"]["eigs_top"]:
        assert np.min(np.abs(lam - complex(re, im))) < 1e-9


def test_m999_pattern_and_values_on_sampled_rows(ca, nagata_H):
    """BASELINE.json configs[3]: the ~1000-mode basis (J,K,L = 5,5,16).  Total count and a sample of output rows
    against the vectorised exact-table tier; two rows against all-rational arithmetic."""
    from oracle.exact import assemble_N_exact
    from oracle.fastfloat import FloatTables, assemble_N_float
    ijkl = ca.basisIndices(5, 5, 16, lib_H(ca))
    assert ijkl.shape[0] == 999
    slab = ca.AssemblySlab(1.0, 2.0, ijkl)
    assert slab.nnz == 23608304
    ijk, val = slab.coo_host()
    assert np.all(np.diff(ijk[:, 0]) >= 0)                       # lexicographic: i non-decreasing
    rows = list(range(0, 999, 53))
    fijk, fval = assemble_N_float(1, 2, ijkl, rows=rows, tables=FloatTables(1, 2, ijkl))
    sel = np.isin(ijk[:, 0] - 1, rows)
    assert np.array_equal(ijk[sel], fijk)
    assert maxrel(val[sel], fval) < TOL
    eijk, evals = assemble_N_exact(1, 2, ijkl, rows=[0, 998])
    sel = np.isin(ijk[:, 0] - 1, [0, 998])
    assert np.array_equal(ijk[sel], eijk)
    ev = np.array([float(v) for v in evals])
    assert np.max(np.abs(val[sel] - ev)) / np.max(np.abs(ev)) < TOL


@pytest.mark.parametrize("JKL,grid", [((1, 1, 3), (0, 0, 0)), ((2, 2, 3), (0, 0, 0)), ((2, 2, 3), (9, 12, 8)),
                                      ((3, 3, 12), (0, 0, 0))], ids=["17d", "44d", "44d-finer-grid", "307"])
def test_dense_quadrature_path_matches_separable_path(ca, JKL, grid):
    """The brute-force tensor-grid contraction (FP64 tensor pipe) reproduces the separable assembly: same COO
    pattern and order, values to 1e-12 of the largest entry."""
    ijkl = ca.basisIndices(*JKL, lib_H(ca))
    a = ca.AssemblySlab(1.0, 2.0, ijkl)
    b = ca.AssemblySlab(1.0, 2.0, ijkl, dense=True, grid=grid)
    (ia, va), (ib, vb) = a.coo_host(), b.coo_host()
    assert a.nnz == b.nnz and np.array_equal(ia, ib)
    assert maxrel(vb, va) < TOL
    info = b.dense_info()
    J, K, L = JKL
    assert info["grid"][0] > 3 * J and info["grid"][2] > 3 * K and info["grid"][1] >= (3 * L + 13) // 2
    assert info["contract_seconds"] > 0
    for x, y in zip(a.linear_host(), b.linear_host()):
        assert np.array_equal(x, y)
    with pytest.raises(ca.CloudAtlasError, match="not exact"):
        ca.AssemblySlab(1.0, 2.0, ijkl, dense=True, grid=(3 * J, 0, 0))


def test_dense_quadrature_row_slabs(ca):
    ijkl = ca.basisIndices(2, 2, 3, lib_H(ca))
    full = ca.AssemblySlab(1.0, 2.0, ijkl, dense=True).coo_host()
    parts = [ca.AssemblySlab(1.0, 2.0, ijkl, False, r0, r1, dense=True).coo_host() for r0, r1 in ((0, 13), (13, 44))]
    assert np.array_equal(np.concatenate([p[0] for p in parts]), full[0])
    assert np.array_equal(np.concatenate([p[1] for p in parts]), full[1])


@pytest.mark.parametrize("shape", ["R", "T", "W"])
def test_dense_contraction_shapes_agree_and_repeat(ca, shape, monkeypatch):
    """The three tile shapes of the dense contraction (R: three-stage mbarrier ring, the default; T, W: the two-buffer
    kernels) give the same operator on the 307-mode basis, rows 100:180 (two row tiles, ragged columns), and the same
    bits on a second run -- the ring kernel has no CTA barrier in its loop, a missing dependency would show here."""
    ijkl = ca.basisIndices(3, 3, 12, lib_H(ca))
    sep = ca.AssemblySlab(1.0, 2.0, ijkl, False, 100, 180).coo_host()
    monkeypatch.setenv("CA_DENSE_CFG", shape)
    a = ca.AssemblySlab(1.0, 2.0, ijkl, False, 100, 180, dense=True).coo_host()
    b = ca.AssemblySlab(1.0, 2.0, ijkl, False, 100, 180, dense=True).coo_host()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.array_equal(a[0], sep[0])
    assert np.max(np.abs(a[1] - sep[1])) < 1e-13 * np.max(np.abs(sep[1]))


def test_random_coefficient_basis_mode_scale(ca):
    """north_star: "random-coefficient bases" = the reference's data model c * Psi (BasisFunctions.jl:417) with
    s_n ~ U(0.5, 2), seed 20251029: N'_ijk = s_i s_j s_k N_ijk, B'_ij = s_i s_j B_ij, same pattern; both
    assembly paths; the model built on the scaled basis satisfies f'(x) = S^-1 f(S x)."""
    rng = np.random.default_rng(20251029)
    for JKL in ((2, 2, 3), (3, 3, 12)):
        ijkl = ca.basisIndices(*JKL, lib_H(ca))
        m = ijkl.shape[0]
        s = rng.uniform(0.5, 2.0, m)
        plain = ca.AssemblySlab(1.0, 2.0, ijkl)
        scaled = ca.AssemblySlab(1.0, 2.0, ijkl, mode_scale=s)
        (ijk, v), (ijk2, v2) = plain.coo_host(), scaled.coo_host()
        assert np.array_equal(ijk, ijk2)
        want = v * s[ijk[:, 0] - 1] * s[ijk[:, 1] - 1] * s[ijk[:, 2] - 1]
        assert maxrel(v2, want) < 1e-13                              # the assembly metric: max-abs over max-abs
        assert np.max(np.abs(v2 - want) / np.abs(want)) < 1e-10      # entrywise (entries are sums with cancellation)
        for P, Q in zip(plain.linear_host(), scaled.linear_host()):
            assert maxrel(Q, P * np.outer(s, s)) < 1e-13
        if m < 100:
            dense = ca.AssemblySlab(1.0, 2.0, ijkl, dense=True, mode_scale=s)
            assert np.array_equal(dense.coo_host()[0], ijk)
            assert maxrel(dense.coo_host()[1], want) < TOL
    with pytest.raises(ca.CloudAtlasError, match="non-zero"):
        ca.AssemblySlab(1.0, 2.0, ijkl, mode_scale=np.zeros(m))
    # the ODE of the scaled basis is the same flow in rescaled coordinates: x' = S^-1 x
    H = lib_H(ca)
    s = rng.uniform(0.5, 2.0, 44)
    a = ca.ODEModel(1.0, 2.0, 2, 2, 3, H)
    b = ca.ODEModel(1.0, 2.0, 2, 2, 3, H, mode_scale=s)
    x = 0.1 * rng.standard_normal(44)
    assert maxrel(b.f(x / s, 200.0) * s, a.f(x, 200.0)) < 1e-11
    assert abs(b.shear(x / s) - a.shear(x)) < 1e-13 and abs(b.dissipation(x / s) - a.dissipation(x)) < 1e-12


def test_mixed_basis_dense_quadrature(ca):
    """The dense quadrature path on a basis that does NOT factorise: Phi = Psi C, C = I + 0.1 G, G_ij ~ N(0,1)
    (SURVEY 8d).  N' = N x1 C x2 C x3 C and B' = C' B C against numpy contractions of the separable result."""
    rng = np.random.default_rng(20251029)
    for JKL, slab in (((1, 1, 3), (0, 17)), ((2, 2, 3), (0, 44)), ((2, 2, 3), (13, 30))):
        ijkl = ca.basisIndices(*JKL, lib_H(ca))
        m = ijkl.shape[0]
        Cm = np.eye(m) + 0.1 * rng.standard_normal((m, m))
        plain = ca.AssemblySlab(1.0, 2.0, ijkl)
        ijk, v = plain.coo_host()
        N = np.zeros((m, m, m))
        N[ijk[:, 0] - 1, ijk[:, 1] - 1, ijk[:, 2] - 1] = v
        want = np.einsum("ia,jb,kc,ijk->abc", Cm, Cm, Cm, N, optimize=True)
        mixed = ca.AssemblySlab(1.0, 2.0, ijkl, False, slab[0], slab[1], dense=True, mix=Cm)
        mijk, mv = mixed.coo_host()
        got = np.zeros((m, m, m))
        got[mijk[:, 0] - 1, mijk[:, 1] - 1, mijk[:, 2] - 1] = mv
        assert np.all(mijk[:, 0] - 1 >= slab[0]) and np.all(mijk[:, 0] - 1 < slab[1])
        assert np.max(np.abs(got[slab[0]:slab[1]] - want[slab[0]:slab[1]])) / np.max(np.abs(want)) < TOL
        for P, Q in zip(plain.linear_host(), mixed.linear_host()):
            assert maxrel(Q[slab[0]:slab[1]], (Cm.T @ P @ Cm)[slab[0]:slab[1]]) < 1e-13
    with pytest.raises(ca.CloudAtlasError, match="dense"):
        ca.AssemblySlab(1.0, 2.0, ijkl, mix=Cm)
