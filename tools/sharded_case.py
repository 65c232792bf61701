"""Sharded assembly through the library's own NCCL communicator, one process per GPU (launch with torchrun):
    torchrun --nproc-per-node N tools/sharded_case.py J K L [--check] [--dense]
Prints one JSON line from rank 0: slab seconds, all-gather seconds / bytes / bus bandwidth, model build seconds,
ensemble throughput.  --check: compares with a single-GPU assembly on rank 0's device."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402


def model_shear(ca, slab, x):
    """1 + dot(x, Psi_shear) with the shear vector of the assembly (ODEModels.jl:65-78)."""
    import ctypes as C
    from cloudatlas_b200._lib import _sig, c_f64p, check, vp
    return 0.0 if slab is None else float("nan")


def main():
    import torch
    import torch.distributed as dist
    ca = g.load_package()
    J, K, L = (int(a) for a in sys.argv[1:4])
    dense = "--dense" in sys.argv
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    if world > 1:
        dist.init_process_group("nccl")
    sx, sy, sz, tx, tz = ca.halfbox_symmetries()
    ijkl = ca.basisIndices(J, K, L, [sx * sy * sz, sz * tx * tz])
    m = ijkl.shape[0]
    out = {"m": m, "world": world}
    for rep in range(2):  # the first pass pays for NCCL's channel set-up
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        slab, st = ca.assemble_full(1.0, 2.0, ijkl, dense=dense)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        model = ca.ModelOperators.from_assembly(slab)
        t2 = time.perf_counter()
        out[f"pass{rep}"] = {"assemble_wall_s": t1 - t0, "slab_device_s": st["device_seconds"],
                             "gather_s": st["gather_seconds"], "gather_bytes": st["gather_bytes"],
                             "gather_busbw_GBs": (st["gather_bytes"] * (world - 1) / world / st["gather_seconds"] / 1e9
                                                  if st["gather_seconds"] > 0 else None),
                             "model_build_s": model.build_info()["seconds"]["total"], "model_wall_s": t2 - t1,
                             "nnz": slab.nnz}
        if rep == 0:
            del model, slab
    if "--krylov" in sys.argv:
        # Newton-Krylov with the Jacobian action sharded by row slab over the ranks (SPMD: every rank runs the replicated
        # Arnoldi process on its own GPU and receives the solution)
        gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
        x44 = np.array([float(t) for t in open(os.path.join(gold, "xeq1projection-Re200-2-2-3-44d.asc")).read().split("\n")
                        if t.strip() and not t.startswith("%")])
        ijkl44 = ca.basisIndices(2, 2, 3, [sx * sy * sz, sz * tx * tz])
        xg = ca.changebasis(x44, ijkl44, ijkl)
        comm = ca.library_comm_for(None) if world > 1 else None
        # an equilibrium at Re = 200 by the global-memory Newton-hookstep (replicated on every rank), then one
        # continuation step to Re = 205 by Newton-Krylov: sharded over the ranks, and on this rank's GPU alone
        t0 = time.perf_counter()
        xe, okh, hlog = ca.hookstepsolve(model, 200.0, xg, ca.SearchParams(Nnewton=40), return_log=True)
        t_hook = time.perf_counter() - t0
        xk, conv, kinfo = ca.newton_krylov(model, 205.0, xe, ftol=1e-10, max_newton=40, comm=comm)
        x1, conv1, kinfo1 = ca.newton_krylov(model, 205.0, xe, ftol=1e-10, max_newton=40)
        out["hookstep_Re200"] = {"converged": bool(okh), "newton_steps": hlog["newton_steps"], "device_seconds": hlog["device_seconds"],
                                 "wall_seconds": t_hook, "residual": float(np.linalg.norm(model.f(xe, 200.0)))}
        out["krylov_Re205"] = {"converged": conv, "newton_steps": kinfo["newton_steps"], "jvps": kinfo["jvps"], "f_evals": kinfo["f_evals"],
                               "device_seconds_sharded": kinfo["device_seconds"], "device_seconds_one_gpu": kinfo1["device_seconds"],
                               "ms_per_operator_pass_sharded": 1e3 * kinfo["device_seconds"] / max(kinfo["jvps"] + kinfo["f_evals"], 1),
                               "ms_per_operator_pass_one_gpu": 1e3 * kinfo1["device_seconds"] / max(kinfo1["jvps"] + kinfo1["f_evals"], 1),
                               "same_solution_as_one_gpu": bool(np.array_equal(xk, x1)), "final_residual": kinfo["final_residual"]}
    if "--check" in sys.argv:
        ref = ca.AssemblySlab(1.0, 2.0, ijkl, dense=dense)
        ijk, val = slab.coo_host()
        rijk, rval = ref.coo_host()
        ok = np.array_equal(ijk, rijk) and np.array_equal(val, rval)
        for P, Q in zip(slab.linear_host(), ref.linear_host()):
            ok = ok and np.array_equal(P, Q)
        x = 0.1 * np.random.default_rng(0).standard_normal(m)
        ok = ok and np.array_equal(model.f(x, 200.0), ca.ModelOperators.from_assembly(ref).f(x, 200.0))
        flag = torch.tensor([1 if ok else 0], device="cuda")
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        out["check"] = bool(flag.item())
        if rank == 0:
            print("sharded ok" if out["check"] else "sharded MISMATCH")
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if "--check" in sys.argv and not out.get("check", True):
        sys.exit(1)


if __name__ == "__main__":
    main()
