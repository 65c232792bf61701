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
