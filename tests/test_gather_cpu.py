"""The assembly all-gather on CPU: world_size-2 gloo, ragged slabs, order preserved."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    ca = g.load_package()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = 11
    bounds = ca.row_partition(m, world)
    rows = bounds[rank + 1] - bounds[rank]
    rng = np.random.default_rng(100 + rank)
    nnz = [7, 0, 19][rank % 3]
    lin = torch.from_numpy(rng.standard_normal((3, rows, m)))
    idx = torch.from_numpy(rng.integers(1, m + 1, size=(3, nnz)))
    val = torch.from_numpy(rng.standard_normal(nnz))
    L, I, V = ca.all_gather_slabs(lin, idx, val, m)
    q.put((rank, L.numpy(), I.numpy(), V.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_all_gather_slabs_gloo(world):
    import __graft_entry__ as g
    ca = g.load_package()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    m = 11
    bounds = ca.row_partition(m, world)
    assert bounds[0] == 0 and bounds[-1] == m and all(b1 >= b0 for b0, b1 in zip(bounds, bounds[1:]))
    # expected = concatenation in rank order of what each rank generated
    lins, idxs, vals = [], [], []
    for r in range(world):
        rng = np.random.default_rng(100 + r)
        rows = bounds[r + 1] - bounds[r]
        nnz = [7, 0, 19][r % 3]
        lins.append(rng.standard_normal((3, rows, m)))
        idxs.append(rng.integers(1, m + 1, size=(3, nnz)))
        vals.append(rng.standard_normal(nnz))
    L, I, V = np.concatenate(lins, axis=1), np.concatenate(idxs, axis=1), np.concatenate(vals)
    for _, l, i, v in results:
        assert np.array_equal(l, L) and np.array_equal(i, I) and np.array_equal(v, V)


def test_row_partition_alignment():
    import __graft_entry__ as g
    ca = g.load_package()
    for m, world, align in ((999, 8, 128), (999, 4, 128), (307, 8, 128), (44, 3, 1), (2962, 8, 128), (17, 8, 1)):
        b = ca.row_partition(m, world, align)
        assert len(b) == world + 1 and b[0] == 0 and b[-1] == m
        assert all(x <= y for x, y in zip(b, b[1:]))
        assert all(x % align == 0 or x == m for x in b)
    assert ca.row_partition(999, 8, 128) == [0, 128, 256, 384, 512, 640, 768, 896, 999]
