"""BASELINE.json configs[4] end to end: large basis (default J,K,L = 8,8,20 -> 2962 modes), Q assembly sharded by
output mode over the ranks with one NCCL all-gather, then the batched matrix-free Jacobian action (Newton-Krylov's
kernel) sharded by state.  Launch with torchrun (one rank per GPU) or plain python (1 GPU).  Rank 0 prints one JSON line.
    torchrun --nproc-per-node 8 tools/config5_multi_gpu.py [J,K,L] [states_per_gpu]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import __graft_entry__ as g

world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
os.environ.setdefault("CA_NUM_THREADS", str(max(1, (os.cpu_count() or 8) // world)))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ca = g.load_package()
assembly_only = "--assembly-only" in sys.argv
argv = [a for a in sys.argv[1:] if not a.startswith("--")]
J, K, L = (int(v) for v in (argv[0] if len(argv) > 0 else "8,8,20").split(","))
nstates = int(argv[1]) if len(argv) > 1 else 9472
sx, sy, sz, tx, tz = ca.halfbox_symmetries()
H = [sx * sy * sz, sz * tx * tz]


def sync():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def rmax(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---- assembly alone (sharded slab + all-gather), then the full model build
ijkl = ca.basisIndices(J, K, L, H)
m = ijkl.shape[0]
bounds = ca.row_partition(m, world)
sync()
t0 = time.perf_counter()
slab = ca.AssemblySlab(1.0, 2.0, ijkl, False, bounds[rank], bounds[rank + 1])
sync()
t_slab = rmax(time.perf_counter() - t0)
slab_dev = rmax(slab.device_seconds)
nnz_local = slab.nnz
t0 = time.perf_counter()
if world > 1:
    lin, idx, val = slab.to_torch()
    lin, idx, val = ca.all_gather_slabs(lin, idx, val, m)
    torch.cuda.synchronize()
    nnz = int(val.shape[0])
    gather_bytes = int(val.numel() * 8 + idx.numel() * 8 + lin.numel() * 8)
    check = [int(idx[0, 0]), int(idx[0, -1]), float(val.double().abs().sum())]   # same on every rank
    del lin, idx, val
    # second gather of the same slab: the first one also pays for NCCL's lazy channel set-up and the allocations
    sync()
    t1 = time.perf_counter()
    lin, idx, val = slab.to_torch()
    lin, idx, val = ca.all_gather_slabs(lin, idx, val, m)
    torch.cuda.synchronize()
    del lin, idx, val
    sync()
    t_gather2 = rmax(time.perf_counter() - t1)
else:
    nnz, gather_bytes, check, t_gather2 = nnz_local, 0, None, 0.0
sync()
t_gather = rmax(time.perf_counter() - t0) - t_gather2
if assembly_only:
    if rank == 0:
        print(json.dumps({"JKL": [J, K, L], "m": m, "n_gpus": world, "nnz": nnz, "assembly_slab_seconds": t_slab,
                          "assembly_slab_device_seconds": slab_dev, "all_gather_seconds_first": t_gather,
                          "all_gather_seconds": t_gather2, "all_gather_bytes": gather_bytes, "check": check}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0)
del slab
torch.cuda.empty_cache()
sync()
t0 = time.perf_counter()
model = ca.ODEModel(1.0, 2.0, J, K, L, H)
sync()
t_build = rmax(time.perf_counter() - t0)

gen = torch.Generator(device="cuda").manual_seed(7 + rank)
X = (0.1 * torch.randn((m, nstates), dtype=torch.float64, device="cuda", generator=gen)).t()
V = torch.randn((m, nstates), dtype=torch.float64, device="cuda", generator=gen).t()
res = {}
for name, fn in (("jvp", lambda: model.Df_action(X, V, 200.0)), ("rhs", lambda: model.f(X, 200.0))):
    fn()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2):
        fn()
    e1.record()
    sync()
    ms = rmax(e0.elapsed_time(e1) / 2)
    res[name + "_ms"] = ms
    res[name + "_per_sec_total"] = world * nstates / (ms * 1e-3)
x1, v1 = X[:1], V[:1]
model.Df_action(x1, v1, 200.0)
sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    model.Df_action(x1, v1, 200.0)
e1.record()
sync()
res["single_state_jvp_ms"] = e0.elapsed_time(e1) / 5
info = model.info()
if rank == 0:
    print(json.dumps({"JKL": [J, K, L], "m": m, "n_gpus": world, "nnz": nnz, "assembly_slab_seconds": t_slab,
                      "assembly_slab_device_seconds": slab_dev, "all_gather_seconds_first": t_gather,
                      "all_gather_seconds": t_gather2, "all_gather_bytes": gather_bytes, "model_build_seconds": t_build, "states_per_gpu": nstates,
                      "executed_fma_per_state": info["padded_fma_per_state"], **res,
                      "jvp_executed_tflops_per_gpu": 2 * info["padded_fma_per_state"] * nstates / (res["jvp_ms"] * 1e-3) / 1e12,
                      "single_state_stream_gbs": (12 * info["nnz_q_sym"] + 16 * m * m) / (res["single_state_jvp_ms"] * 1e-3) / 1e9}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
