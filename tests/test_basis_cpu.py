"""Index sets and the tabular basis: bit-exact against the oracle (which is pinned to the reference's
stored outputs) -- host-only library calls, no device needed."""
import numpy as np
import pytest

import __graft_entry__ as g

ca = g.load_package()


def lib_H():
    sx, sy, sz, tx, tz = ca.halfbox_symmetries()
    return [sx * sy * sz, sz * tx * tz]


@pytest.mark.parametrize("JKL", [(0, 0, 0), (0, 0, 1), (1, 0, 2), (0, 2, 1), (1, 1, 3), (1, 2, 3), (2, 2, 3), (3, 3, 12), (5, 5, 16)])
def test_basis_indices_bit_exact(JKL, nagata_H):
    from oracle.basis import basisIndices as ref
    assert np.array_equal(ca.basisIndices(*JKL), ref(*JKL))
    assert np.array_equal(ca.basisIndices(*JKL, lib_H()), ref(*JKL, nagata_H))


def test_mode_counts_of_the_baseline_configs():
    H = lib_H()
    counts = {(1, 1, 3): 17, (1, 2, 3): 27, (2, 2, 3): 44, (3, 3, 12): 307, (5, 5, 16): 999, (8, 8, 20): 2962}
    for JKL, m in counts.items():
        assert ca.basisIndices(*JKL, H).shape[0] == m
    J, K, L = 1, 1, 3
    assert ca.basisIndices(J, K, L).shape[0] == (2 * J + 1) * (2 * K + 1) * (2 * L + 1) + 1 == 64  # find-eqb.ipynb:97


def test_every_generator_subset(nagata_H):
    from oracle.basis import basisIndices as ref, halfbox_symmetries as ref_hs
    gens, rgens = ca.halfbox_symmetries(), ref_hs()
    for mask in range(1, 32):
        H = [gens[b] for b in range(5) if mask >> b & 1]
        RH = [rgens[b] for b in range(5) if mask >> b & 1]
        assert np.array_equal(ca.basisIndices(2, 2, 3, H), ref(2, 2, 3, RH))
    # antisymmetric subspaces (s = -1)
    H = [ca.Symmetry(-1, -1, 1, 0, 0, -1)]
    from oracle.basis import Symmetry as RS
    assert np.array_equal(ca.basisIndices(2, 1, 3, H), ref(2, 1, 3, [RS(-1, -1, 1, 0, 0, -1)]))


def test_symmetric_single_rows(nagata_H):
    from oracle.basis import symmetric as ref_sym, basisIndices as ref
    H = lib_H()
    for row in ref(2, 2, 3):
        assert ca.symmetric(row, H) == ref_sym(tuple(int(v) for v in row), nagata_H)
    with pytest.raises(ca.CloudAtlasError, match="half-box"):
        ca.basisIndices(1, 1, 1, [ca.Symmetry(1, 1, 1, ca.basis.Fraction(1, 3), 0)])


def test_group_product():
    sx, sy, sz, tx, tz = ca.halfbox_symmetries()
    g = sz * tx * tz
    assert (g.sx, g.sy, g.sz, g.ax, g.az, g.s) == (1, 1, -1, ca.basis.Fraction(1, 2), ca.basis.Fraction(1, 2), 1)
    assert (tx * tx).ax == 0


@pytest.mark.parametrize("normalize", [False, True])
def test_components_match_oracle_basis_set(normalize, nagata_H):
    from oracle.basis import basisIndices as ref, basisSet
    ijkl = ref(2, 2, 3, nagata_H)
    Psi = basisSet(1.0, 2.0, ijkl, normalize=normalize)
    T = ca.basisComponents(1.0, 2.0, ijkl, normalize=normalize)
    for n, f in enumerate(Psi):
        for c in range(3):
            if f[c].coeff == 0:
                assert T["coef"][n, c] == 0
                continue
            assert abs(T["coef"][n, c] / f[c].coeff - 1) < 1e-14  # the reference norm cancels in monomial form
            assert T["jx"][n, c] == f[c].ejx.waveindex and T["kz"][n, c] == f[c].ekz.waveindex
            assert T["l"][n, c] == ijkl[n, 3]


def test_normalized_coefficients_of_the_notebook(known):
    # find-eqb.ipynb:163-167 (coefficients printed by psistr)
    k = known["m17_normalized"]
    ijkl = ca.basisIndices(1, 1, 3, lib_H())
    assert ijkl[:5].tolist() == k["ijkl_first5"]
    T = ca.basisComponents(1.0, 2.0, ijkl, normalize=True)
    for n in range(4):
        assert abs(T["coef"][n, 0] / k["psi_coeffs_first5"][n] - 1) < 1e-14
    assert abs(T["coef"][4, 1] / k["psi_coeffs_first5"][4][0] - 1) < 1e-14
    assert abs(T["coef"][4, 2] / k["psi_coeffs_first5"][4][1] - 1) < 1e-14


def test_symmetric_on_rows_outside_the_generated_set(nagata_H):
    """Symmetries.jl:25-60 is a sign computation on the row: it answers for any (i,j,k,l) with i in 1:6, also rows
    basisIndices never emits (family 1 with j != 0, l = 0 of the families that start at l = 1), and raises for i
    outside 1:6 where a parity table is consulted."""
    import pytest
    from oracle.basis import symmetric as ref_sym
    sx, sy, sz, tx, tz = ca.halfbox_symmetries()
    rsx, rsy, rsz, rtx, rtz = __import__("oracle.basis", fromlist=["halfbox_symmetries"]).halfbox_symmetries()
    gens = [(sx, rsx), (sy, rsy), (sz, rsz), (tx, rtx), (tz, rtz), (sx * sy * sz, rsx * rsy * rsz), (sz * tx * tz, rsz * rtx * rtz)]
    for i in range(1, 7):
        for j in (-2, -1, 0, 1, 3):
            for k in (-3, 0, 1, 2):
                for l in (0, 1, 2, 5):
                    for mine, theirs in gens:
                        assert ca.symmetric((i, j, k, l), mine) == ref_sym((i, j, k, l), theirs), (i, j, k, l)
    assert ca.symmetric((9, 0, 0, 2), sy) == ref_sym((9, 0, 0, 2), rsy)      # no parity table consulted: no error
    with pytest.raises(ca.CloudAtlasError, match="not in 1:6"):
        ca.symmetric((9, 0, 0, 2), sx)
