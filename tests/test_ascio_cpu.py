"""`.asc` I/O and changebasis: host utilities, checked on the reference's shipped data."""
import os

import numpy as np

import __graft_entry__ as g
from conftest import GOLDEN, read_asc

ca = g.load_package()


def lib_H():
    sx, sy, sz, tx, tz = ca.halfbox_symmetries()
    return [sx * sy * sz, sz * tx * tz]


def test_load_reads_the_shipped_projections():
    for name in ("xeq1projection-Re200-1-2-3-27d.asc", "xeq2projection-Re200-2-2-3-44d.asc"):
        x = ca.load(os.path.join(GOLDEN, name))
        assert np.array_equal(x, read_asc(name))


def test_save_load_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    x, A = rng.standard_normal(13), rng.standard_normal((5, 7))
    ca.save(x, str(tmp_path / "x"))
    ca.save(A, str(tmp_path / "A.asc"), cc="%")
    assert np.array_equal(ca.load(str(tmp_path / "x.asc")), x)       # repr() round-trips doubles exactly
    assert np.array_equal(ca.load(str(tmp_path / "A.asc")), A)
    assert open(tmp_path / "x.asc").readline() == "# 13\n"
    assert open(tmp_path / "A.asc").readline() == "% 5 7\n"
    ijkl = ca.basisIndices(1, 1, 3, lib_H())
    ca.ijkl2file(ijkl, str(tmp_path / "ijkl"))
    back = ca.load(str(tmp_path / "ijkl.asc")).astype(np.int64)
    assert np.array_equal(back, ijkl)


def test_changebasis_nested_bases():
    """The 27d basis is the head of the 44d basis (j is the outer loop), and the shipped 44d projection starts with
    the 27d one."""
    H = lib_H()
    i27, i44 = ca.basisIndices(1, 2, 3, H), ca.basisIndices(2, 2, 3, H)
    d = ca.basis_index_dict(i27, i44)
    assert d == {n: n for n in range(27)}
    x27 = read_asc("xeq1projection-Re200-1-2-3-27d.asc")
    x44 = read_asc("xeq1projection-Re200-2-2-3-44d.asc")
    up = ca.changebasis(x27, i27, i44)
    assert np.array_equal(up[:27], x27) and np.all(up[27:] == 0) and np.array_equal(up[:27], x44[:27])
    down = ca.changebasis(x44, i44, i27)
    assert np.array_equal(down, x27)
    # non-nested: 17d (1,1,3) into 27d (1,2,3)
    i17 = ca.basisIndices(1, 1, 3, H)
    d17 = ca.basis_index_dict(i17, i27)
    assert len(d17) == 17 and all(np.array_equal(i17[n], i27[m]) for n, m in d17.items())
