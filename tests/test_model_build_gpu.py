"""The device model build (ca_model_create_dev, csrc/model_build.cu) against the host build (the round-1 path,
forced with CA_BUILD_HOST=1): both must produce the SAME pack, bit for bit -- tile descriptors, pair words, values,
row lists, windows, window mode lists, warp schedule, counters, L1, L2, linear slots, streaming form (twelve digests).
Reference behaviour replaced: Bfact = lu(B) and the closures of src/ODEModels.jl:51-58."""
import os

import numpy as np
import pytest

import __graft_entry__ as g

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ca():
    return g.load_package()


class env:
    """Temporarily set / unset environment variables the library reads with getenv() at call time."""

    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        for k, v in self.kv.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(v)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def oracle_operators(JKL, H, normalize=False):
    from oracle.model import ODEModel
    ref = ODEModel(1.0, 2.0, *JKL, H, normalize=normalize)
    return ref.B, ref.A1, ref.A2, np.asfortranarray(ref.N.ijk), ref.N.val


@pytest.fixture(scope="module")
def ops307(nagata_H):
    from oracle.basis import basisIndices
    from oracle.fastfloat import FloatTables, assemble_N_float, assemble_linear_float
    ijkl = np.array(basisIndices(3, 3, 12, nagata_H))
    T = FloatTables(1, 2, ijkl)
    ijk, val = assemble_N_float(1, 2, ijkl, tables=T)
    B, A1, A2 = assemble_linear_float(1, 2, ijkl, tables=T)
    return B, A1, A2, np.asfortranarray(ijk), val


def build_both(ca, ops, **force):
    with env(CA_BUILD_HOST="1", **force):
        host = ca.ModelOperators(*ops)
    with env(CA_BUILD_HOST=None, **force):
        dev = ca.ModelOperators(*ops)
    assert not host.build_info()["on_device"]
    assert dev.build_info()["on_device"]
    return host, dev


NAMES = ("tiles", "pairs", "vals", "rows", "windows", "modes", "schedule", "counters", "L1", "L2", "linear_slots", "stream")


def assert_same_build(host, dev):
    a, b = host.digest(), dev.digest()
    diff = [n for n, x, y in zip(NAMES, a, b) if x != y]
    assert not diff, f"device build differs from the host build in: {diff}"


@pytest.mark.parametrize("JKL,normalize", [((1, 1, 3), False), ((1, 1, 3), True), ((1, 2, 3), False), ((2, 2, 3), False)],
                         ids=["17d", "17d-normalized", "27d", "44d"])
def test_small_models_same_pack(ca, nagata_H, JKL, normalize):
    ops = oracle_operators(JKL, nagata_H, normalize)
    host, dev = build_both(ca, ops)
    assert_same_build(host, dev)
    x = 0.1 * np.random.default_rng(0).standard_normal(host.m)
    assert np.array_equal(host.f(x, 200.0), dev.f(x, 200.0))
    assert np.array_equal(host.Df(x, 200.0), dev.Df(x, 200.0))       # lazily fetched folded terms
    X = 0.1 * np.random.default_rng(1).standard_normal((5000, host.m))
    assert np.array_equal(host.f(X, 250.0), dev.f(X, 250.0))          # generated small kernel (host copies of L1, L2)


def test_no_symmetry_and_windowed_small(ca):
    """A basis without symmetry reduction (64 modes) and the windowed / three-n-tile packs forced onto it."""
    ops = oracle_operators((1, 1, 3), [])
    for force in ({}, {"CA_FORCE_WINDOW_MODES": "16"}, {"CA_FORCE_WINDOW_MODES": "24", "CA_NT3": "1"}):
        host, dev = build_both(ca, ops, **force)
        assert_same_build(host, dev)


@pytest.mark.parametrize("force", [{}, {"CA_FORCE_WINDOW_MODES": "32"}, {"CA_FORCE_WINDOW_MODES": "128"},
                                   {"CA_FORCE_WINDOW_MODES": "128", "CA_NT3": "1"}, {"CA_PACK_AMAJOR": "1"}],
                         ids=["whole-tile", "window32", "window128", "window128-nt3", "amajor"])
def test_307_same_pack(ca, ops307, force):
    host, dev = build_both(ca, ops307, **force)
    assert_same_build(host, dev)
    info = dev.info()
    assert info == host.info()
    X = 0.1 * np.random.default_rng(2).standard_normal((300, 307))
    with env(**force):
        assert np.array_equal(host.f(X, 200.0), dev.f(X, 200.0))
        assert np.array_equal(host.f(X[0], 200.0), dev.f(X[0], 200.0))  # streaming form (built on the device)


def test_unsorted_or_duplicated_coo_takes_the_host_build(ca, nagata_H):
    B, A1, A2, ijk, val = oracle_operators((1, 1, 3), nagata_H)
    perm = np.random.default_rng(3).permutation(len(val))
    shuffled = ca.ModelOperators(B, A1, A2, np.asfortranarray(ijk[perm]), val[perm])
    assert not shuffled.build_info()["on_device"]
    dup = ca.ModelOperators(B, A1, A2, np.asfortranarray(np.vstack([ijk, ijk[:5]])), np.concatenate([val, val[:5]]))
    assert not dup.build_info()["on_device"]
    plain = ca.ModelOperators(B, A1, A2, ijk, val)
    assert plain.build_info()["on_device"]
    x = 0.1 * np.random.default_rng(4).standard_normal(17)
    assert np.array_equal(shuffled.f(x, 200.0), plain.f(x, 200.0))


def test_singular_B_is_an_error(ca, nagata_H):
    B, A1, A2, ijk, val = oracle_operators((1, 1, 3), nagata_H)
    B = B.copy()
    B[:, 3] = 0.0
    B[3, :] = 0.0
    with pytest.raises(ca.CloudAtlasError, match="singular"):
        ca.ModelOperators(B, A1, A2, ijk, val)


def test_from_device_999_same_pack_and_speed(ca):
    """BASELINE configs[3] size: assembled on the GPU, built from the device arrays; same pack as the host build of
    the downloaded operator; the production windowed kernel gives bit-identical f and Jacobian actions."""
    import torch
    sx, sy, sz, tx, tz = ca.halfbox_symmetries()
    ijkl = ca.basisIndices(5, 5, 16, [sx * sy * sz, sz * tx * tz])
    lin, idx, val, _ = ca.assemble_device(1.0, 2.0, ijkl)
    assert lin.shape[1] == 999 and val.shape[0] == 23_608_304
    dev = ca.ModelOperators.from_device(lin[0].t(), lin[1].t(), lin[2].t(), idx, val)
    bi = dev.build_info()
    assert bi["on_device"]
    assert bi["seconds"]["total"] < 1.0, bi     # 3 s on the host (round 1); measured here: see bench.py model_build
    with env(CA_BUILD_HOST="1"):
        host = ca.ModelOperators(dev.B, dev.A1, dev.A2, dev.ijk, dev.val)
    assert not host.build_info()["on_device"]
    assert_same_build(host, dev)
    assert dev.kernel_path(4096)[0] == "dmma_row_tiles_windowed"
    rng = np.random.default_rng(5)
    X = torch.from_numpy(0.1 * rng.standard_normal((999, 2048))).cuda().t()
    V = torch.from_numpy(rng.standard_normal((999, 2048))).cuda().t()
    assert torch.equal(host.f(X, 200.0), dev.f(X, 200.0))
    assert torch.equal(host.Df_action(X, V, 200.0), dev.Df_action(X, V, 200.0))
    x = X[0].contiguous()
    assert torch.equal(host.f(x, 200.0), dev.f(x, 200.0))
