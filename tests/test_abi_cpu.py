"""CPU-only checks of the C-ABI boundary: the library loads, exports every declared symbol, rejects bad
arguments like the reference constructor does, refuses to compute without a device, and the host-side
packer reproduces its input exactly (structure round trip).  No compute calls."""
import ctypes as C

import numpy as np
import pytest

import __graft_entry__ as g

ca = g.load_package()
from cloudatlas_b200._lib import _sig, c_f64p, c_i64p, check, exported_symbols, lib  # noqa: E402

_selftest = _sig("ca_selftest_pack", c_i64p, c_f64p, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                 c_i64p, c_i64p, c_i64p, c_i64p, c_i64p, c_i64p, c_i64p)


def pack_stats(ijk, val, m, symmetric, window_modes=0):
    ijk = np.asfortranarray(ijk, dtype=np.int64)
    val = np.ascontiguousarray(val, dtype=np.float64)
    out = [C.c_int64() for _ in range(7)]
    check(_selftest(ijk.ctypes.data_as(c_i64p), val.ctypes.data_as(c_f64p), len(val), m, symmetric, window_modes,
                    *[C.byref(o) for o in out]))
    return dict(zip(("ntiles", "nterms", "padded_fma", "pair_products", "mismatches", "nwindows",
                     "max_window_modes"), (o.value for o in out)))


def test_every_declared_symbol_is_exported():
    names = exported_symbols()
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_constructor_errors_match_reference():
    # SparseBilinear.jl:8-9
    with pytest.raises(ca.CloudAtlasError, match="three columns"):
        ca.SparseBilinear(np.ones((4, 2), dtype=np.int64), np.ones(4), 3)
    with pytest.raises(ca.CloudAtlasError, match="same number of rows"):
        ca.SparseBilinear(np.ones((4, 3), dtype=np.int64), np.ones(5), 3)
    with pytest.raises(ca.CloudAtlasError, match="outside 1:3"):
        ca.SparseBilinear(np.array([[1, 4, 1]]), np.ones(1), 3)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(ca.CloudAtlasError) as e:
        ca.SparseBilinear(np.array([[1, 1, 1]]), np.ones(1), 3)
    assert e.value.code == 3 and "no CPU fallback" in str(e.value)


@pytest.mark.parametrize("JKL,nnz", [((1, 1, 3), 396), ((2, 2, 3), 5654)])
def test_pack_roundtrip_models(JKL, nnz, nagata_H):
    from oracle.basis import basisIndices
    from oracle.fastfloat import assemble_N_float
    ijkl = basisIndices(*JKL, nagata_H)
    ijk, val = assemble_N_float(1, 2, ijkl)
    assert len(val) == nnz
    for sym in (1, 0):
        st = pack_stats(ijk, val, ijkl.shape[0], sym)
        assert st["mismatches"] == 0
        assert st["padded_fma"] >= st["nterms"]


@pytest.mark.parametrize("seed", range(6))
def test_pack_roundtrip_random(seed):
    rng = np.random.default_rng(seed)
    m = int(rng.integers(1, 70))
    nnz = int(rng.integers(0, 4 * m * m))
    ijk = rng.integers(1, m + 1, size=(nnz, 3))
    if seed % 2:
        ijk[:, 0] = np.minimum(ijk[:, 0], max(1, m // 2))  # leaves rows without terms
    val = rng.standard_normal(nnz)
    val[rng.random(nnz) < 0.1] = 0.0                        # explicit zeros are dropped
    for sym in (1, 0):
        assert pack_stats(ijk, val, m, sym)["mismatches"] == 0
        for wm in (8, 24, 96):
            st = pack_stats(ijk, val, m, sym, wm)
            assert st["mismatches"] == 0 and st["max_window_modes"] <= wm


def test_pack_sort_paths_above_the_small_input_cutoff():
    """>= 65536 terms: (a) terms in random order (no runs of equal output row: element-wise counting sort),
    (b) whole rows in shuffled order (runs moved as blocks on worker threads).  Both must reproduce the input."""
    rng = np.random.default_rng(11)
    m, nnz = 60, 100_000
    ijk = rng.integers(1, m + 1, size=(nnz, 3))
    val = rng.integers(1, 9, size=nnz).astype(np.float64)   # duplicates abound: small integers sum exactly in any order
    for sym in (1, 0):
        assert pack_stats(ijk, val, m, sym)["mismatches"] == 0
    order = np.lexsort((ijk[:, 2], ijk[:, 1], ijk[:, 0]))                 # (i, j, k) order ...
    ijk_s, val_s = ijk[order], val[order]
    rows = [np.flatnonzero(ijk_s[:, 0] == i) for i in rng.permutation(np.arange(1, m + 1))]
    perm = np.concatenate(rows)                                           # ... with the rows shuffled as blocks
    for sym in (1, 0):
        for wm in (0, 24):
            assert pack_stats(ijk_s[perm], val_s[perm], m, sym, wm)["mismatches"] == 0


def test_pack_three_n_tiles_for_wide_row_clusters(monkeypatch):
    """Windowed packs whose rows mostly sit in clusters of 17-24 rows (the 3000-mode models) get one tile of three
    n-tiles per cluster instead of two half-empty tiles; narrower structures and whole-tile packs keep two."""
    rng = np.random.default_rng(5)
    m, rows_per_cluster = 66, 22
    ijk, val = [], []
    for c in range(3):
        pairs = rng.integers(1, m + 1, size=(90, 2))
        for i in range(c * rows_per_cluster, (c + 1) * rows_per_cluster):
            for j, k in pairs:
                ijk.append((i + 1, j, k))
                val.append(float(rng.integers(1, 9)))
    ijk, val = np.array(ijk), np.array(val)
    wide = pack_stats(ijk, val, m, 1, 32)
    assert wide["mismatches"] == 0 and wide["ntiles"] == 3
    monkeypatch.setenv("CA_NT3", "0")
    narrow = pack_stats(ijk, val, m, 1, 32)
    assert narrow["mismatches"] == 0 and narrow["ntiles"] == 6
    assert wide["padded_fma"] < narrow["padded_fma"] and wide["pair_products"] < narrow["pair_products"]
    monkeypatch.delenv("CA_NT3")
    assert pack_stats(ijk, val, m, 1, 0)["ntiles"] == 6        # whole-tile packs: two n-tiles at most


@pytest.mark.parametrize("JKL", [(1, 1, 3), (1, 2, 3), (2, 2, 3)], ids=["17d", "27d", "44d"])
def test_generated_small_model_kernel_compiles(JKL, nagata_H):
    """The NVRTC source generated for the reference's small models compiles for sm_100a (no device needed for that),
    with the measured CTA shapes: 128 threads x 2 stages at 17 modes, 256 x 2 up to 27, 128 x 2 beyond."""
    from oracle.basis import basisIndices
    from oracle.fastfloat import assemble_N_float
    ijkl = basisIndices(*JKL, nagata_H)
    m = ijkl.shape[0]
    ijk, val = assemble_N_float(1, 2, ijkl)
    ijk = np.asfortranarray(ijk, dtype=np.int64)
    fn = _sig("ca_selftest_small_kernel", c_i64p, c_f64p, C.c_int64, C.c_int64, c_i64p, C.POINTER(C.c_int32),
              C.POINTER(C.c_int32), c_i64p)
    cubin, src = C.c_int64(), C.c_int64()
    threads, stages = C.c_int32(), C.c_int32()
    check(fn(ijk.ctypes.data_as(c_i64p), np.ascontiguousarray(val).ctypes.data_as(c_f64p), len(val), m,
             C.byref(cubin), C.byref(threads), C.byref(stages), C.byref(src)))
    assert cubin.value > 10_000 and src.value > 1_000
    assert (threads.value, stages.value) == {17: (128, 2), 27: (256, 2), 44: (128, 2)}[m]


def test_sass_of_the_hot_kernels():
    """The library is compiled for sm_100a only, its ensemble kernels issue FP64 tensor instructions (DMMA.8x8x4) and
    stream through cp.async (LDGSTS); checked on the SASS of the built library, no device needed."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    libpath = ca.LIB_PATH
    arch = subprocess.run(["cuobjdump", "-lelf", libpath], capture_output=True, text=True).stdout
    assert "sm_100a" in arch and "sm_90" not in arch and "sm_80" not in arch
    sass = subprocess.run(["cuobjdump", "-sass", libpath], capture_output=True, text=True).stdout
    for pattern, need in (("bilinear_warptile_kernel", ("DMMA.8x8x4", "LDGSTS")),
                          # role-split kernels: cp.async completion arrives on mbarriers (ARRIVES.LDGSTSBAR), waits are
                          # SYNCS try-waits
                          ("bilinear_windowed_kernel", ("DMMA.8x8x4", "LDGSTS", "LDGDEPBAR", "ARRIVES.LDGSTSBAR", "SYNCS")),
                          ("dense_contract_ring_kernel", ("DMMA.8x8x4", "LDGSTS", "ARRIVES.LDGSTSBAR", "SYNCS")),
                          ("dense_contract_kernel", ("DMMA.8x8x4", "LDGSTS"))):
        blocks = [b for b in sass.split("Function : ")[1:] if pattern in b.split("\n", 1)[0]]
        assert blocks, pattern
        for b in blocks:
            for mnemonic in need:
                assert mnemonic in b, (pattern, mnemonic)


def test_pack_fill_at_307(nagata_H):
    """The clustering finds the wavenumber/parity classes: >= 75 % of the executed FMAs are real."""
    from oracle.basis import basisIndices
    from oracle.fastfloat import assemble_N_float
    ijkl = basisIndices(3, 3, 12, nagata_H)
    assert ijkl.shape[0] == 307
    ijk, val = assemble_N_float(1, 2, ijkl)
    assert len(val) == 1355658
    st = pack_stats(ijk, val, 307, 1)
    assert st["mismatches"] == 0
    assert st["nterms"] / st["padded_fma"] > 0.75
    # windowed packing (state rows gathered per window): same structure, a few per cent of padding more
    sw = pack_stats(ijk, val, 307, 1, 96)
    assert sw["mismatches"] == 0 and sw["max_window_modes"] <= 96
    assert sw["padded_fma"] < 1.05 * st["padded_fma"]
    # production window size: long windows are cut at the k-step cap of the kernel's pair staging (checked inside
    # the self test together with the per-window value / row offsets)
    s128 = pack_stats(ijk, val, 307, 1, 128)
    assert s128["mismatches"] == 0 and s128["max_window_modes"] <= 128
    assert s128["nwindows"] > sw["ntiles"] and s128["padded_fma"] < 1.05 * st["padded_fma"]


def test_parts_per_block_of_states():
    """csrc/bilinear.cu choose_parts: ensembles that fill whole passes of the machine keep one CTA per block of 32
    states; smaller ones divide the tiles among several CTAs per block (never more than allowed, never more than the
    slots), and the launch never exceeds the slots."""
    import ctypes as C
    from cloudatlas_b200._lib import lib, check
    lib.ca_selftest_choose_parts.argtypes = [C.c_int64, C.c_int64, C.c_int32, C.POINTER(C.c_int32)]
    lib.ca_selftest_choose_parts.restype = C.c_int32

    def parts(nblocks, slots, mx=32):
        p = C.c_int32(0)
        check(lib.ca_selftest_choose_parts(nblocks, slots, mx, C.byref(p)))
        return p.value

    for nblocks, slots in ((148, 148), (296, 148), (296, 296), (592, 148), (31250, 296), (0, 148)):
        assert parts(nblocks, slots) == 1                      # whole passes (the bench sizes), or nothing to do
    assert parts(1, 148) >= 30 and parts(1, 148, 5) == 5         # one block: (nearly) as many parts as allowed
    assert parts(16, 148) == 9 and parts(47, 148) == 3 and parts(93, 148) == 3
    for slots in (148, 296):
        for nblocks in range(1, 2 * slots):
            p = parts(nblocks, slots)
            assert 1 <= p <= 32 and (p == 1 or p * min(nblocks, slots // p) <= slots)
            if nblocks <= slots // 2:
                assert p >= 2                                    # small ensembles are always split
