#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json: batched RHS evals/sec + Q assembly
seconds vs basis size; % of roofline).

A *step* is one pass of the ensemble right-hand side  f(x_s, R) = B^-1 (A1 x_s + A2 x_s / R + N(x_s, x_s))
over one batch of synthetic states on the 307-mode Nagata-symmetric basis (J,K,L = 3,3,12), i.e.
BASELINE.json configs[2].  Each rank (one per GPU) owns its own batch -- states are independent, there is
no data-path collective -- so scaling is weak.  The model is built by the product's own GPU assembly,
sharded over the ranks by output mode with one NCCL all-gather.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--nstates S]
  python bench.py --impl reference ...     # the reference's CPU algorithm (oracle port), host cores

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

KERNEL_NAMES = {"dmma_row_tiles": "bilinear_warptile_kernel<4,0>", "dmma_row_tiles_ksplit": "bilinear_batched_kernel<4,0>",
                "dmma_row_tiles_windowed": "bilinear_windowed_kernel<0>"}
METRIC = "batched_rhs_evals_per_sec"
UNIT = "evals/s"
JKL = (3, 3, 12)
ALPHA, GAMMA, R = 1.0, 2.0, 200.0
SEED = 20251029


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nstates", type=int, default=1_000_000, help="states per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=0, help="states in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-extras", action="store_true", help="skip assembly sweep / HBM case / CPU baseline")
    return ap.parse_args()


def workload_name(nstates):
    return (f"ensemble RHS f(x)=B^-1(A1 x + A2 x/R + N(x,x)), {nstates} random states per GPU, "
            f"307-mode symmetric basis (J,K,L={JKL[0]},{JKL[1]},{JKL[2]}; alpha,gamma={ALPHA},{GAMMA}; R={R})")


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, f"/tmp/ca_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.fh,
                                         stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [t.strip() for t in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] or sm
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power)}


# ------------------------------------------------------------------------------------ reference arm
def oracle_model(jkl):
    """The reference's algorithm on the CPU (oracle port): operators + C loops."""
    from oracle.basis import basisIndices, halfbox_symmetries
    from oracle.cref import CRefModel
    from oracle.fastfloat import FloatTables, assemble_N_float, assemble_linear_float
    sx, sy, sz, tx, tz = halfbox_symmetries()
    ijkl = basisIndices(*jkl, [sx * sy * sz, sz * tx * tz])
    T = FloatTables(ALPHA, GAMMA, ijkl)
    B, A1, A2 = assemble_linear_float(ALPHA, GAMMA, ijkl, T)
    ijk, val = assemble_N_float(ALPHA, GAMMA, ijkl, tables=T)
    return CRefModel(B, A1, A2, ijk, val)


def host_states(nstates, m, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    return np.asfortranarray(0.1 * rng.standard_normal((nstates, m)))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    threads = os.cpu_count() or 1
    model = oracle_model(JKL)
    sample = args.cpu_sample or max(threads * 256, 2048)
    X = host_states(sample, model.m, SEED)
    for _ in range(max(args.warmup, 1)):
        model.f_ensemble(X[: max(threads * 8, 64)], R, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        model.f_ensemble(X, R, threads)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.nstates),
                   "note": "Julia is not installed: the reference's COO loop + dense A1,A2 + LU solve "
                           "(SparseBilinear.jl:56-65, ODEModels.jl:57) restated in C (oracle/csrc), "
                           "pthreads over states on all host cores; each step is a bounded sample"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} states per step of the same 307-mode workload"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------- b200 arm
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ca = g.load_package()
    sx, sy, sz, tx, tz = ca.halfbox_symmetries()
    H = [sx * sy * sz, sz * tx * tz]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- model: GPU assembly (sharded by output mode, one all-gather) + device packing
    comm_init_s = 0.0
    if world > 1:   # the library's own NCCL communicator (the id travels through torch.distributed): set up once
        barrier()
        t0 = time.perf_counter()
        ca.library_comm_for(None)
        barrier()
        comm_init_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    t0 = time.perf_counter()
    model = ca.ODEModel(ALPHA, GAMMA, *JKL, H)
    barrier()
    build_s = max_over_ranks(time.perf_counter() - t0)
    m = model.m
    info = model.info()

    # ---- synthetic states, resident in HBM (2.46 GB per rank at the default size: >> L2, no flush needed)
    n = args.nstates
    gen = torch.Generator(device="cuda").manual_seed(SEED + rank)
    X = (0.1 * torch.randn((m, n), dtype=torch.float64, device="cuda", generator=gen)).t()
    OUT = ca.ensemble_empty(n, m)

    sampler = ClockSampler(local)
    for _ in range(args.warmup):
        model.f(X, R, out=OUT)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = ca.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        model.f(X, R, out=OUT)
    e1.record()
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = ca.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    value = world * n * args.steps / (dev_ms * 1e-3)

    # ---- end to end through the host-pointer C-ABI call: pinned host buffers, H2D + kernel + D2H per step
    Xh = torch.empty((m, n), dtype=torch.float64, pin_memory=True)
    Oh = torch.empty((m, n), dtype=torch.float64, pin_memory=True)
    Xh.copy_(X.t())
    torch.cuda.synchronize()
    Xh_np, Oh_np = Xh.numpy().T, Oh.numpy().T          # nstates x m, column-major views
    model.f_into(Xh_np, R, Oh_np)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        model.f_into(Xh_np, R, Oh_np)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * n * e2e_steps / e2e_s
    e2e_check = float(torch.linalg.norm(Oh[:, :1000].cuda() - OUT.t()[:, :1000]) /
                      torch.linalg.norm(OUT.t()[:, :1000]))
    # ---- the ceiling of that path: the same bytes as plain page-locked copies, H2D and D2H concurrently, every
    # rank at once (what the host's memory system and the PCIe links give N ranks together)
    copy_bytes = min(8 * m * n, 1 << 30)
    dbuf = torch.empty(copy_bytes // 8, dtype=torch.float64, device="cuda")
    dbuf2 = torch.empty(copy_bytes // 8, dtype=torch.float64, device="cuda")
    hin, hout = Xh.view(-1)[: copy_bytes // 8], Oh.view(-1)[: copy_bytes // 8]
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    def copies():
        with torch.cuda.stream(s_in):
            dbuf.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s_out):
            hout.copy_(dbuf2, non_blocking=True)
    copies()
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        copies()
    barrier()
    copy_s = max_over_ranks(time.perf_counter() - t0) / 3
    ceiling = world * (copy_bytes / (8 * m)) / copy_s          # evals/s if the step were nothing but its copies
    e2e_ceiling = {"value": ceiling, "unit": UNIT, "per_rank_GBs_each_way": copy_bytes / copy_s / 1e9,
                   "bytes_each_way": copy_bytes,
                   "how": "page-locked cudaMemcpyAsync H2D and D2H on two streams, all ranks concurrently, barrier on both sides"}
    del Xh, Oh, Xh_np, Oh_np, dbuf, dbuf2, hin, hout

    # ---- roofline of the dominant (only) kernel of the step: the DMMA ensemble kernel in MODE_QUAD
    ijk = model.ijk
    a, b = np.minimum(ijk[:, 1], ijk[:, 2]), np.maximum(ijk[:, 1], ijk[:, 2])
    key = (ijk[:, 0] * (m + 1) + a) * (m + 1) + b
    nnz_sym = int(np.unique(key).size)
    npairs = int(np.unique(a * (m + 1) + b).size)
    nnzB = int(np.count_nonzero(model.B))
    flops_state = 2 * nnz_sym + npairs + 2 * m * m + 2 * nnzB          # SURVEY.md 8(d), minimal count
    bytes_state = 16 * m
    kernel_ms = dev_ms / args.steps                                       # one launch per step
    dmma_peak, dfma_peak = ca.measure_fp64_peak()
    achieved = flops_state * n / (kernel_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_peak_source = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get("rhs_m307_bytes_per_launch")
    except (OSError, ValueError):
        pass
    roofline = {
        "bound": "tensor", "pipe": "fp64 (mma.sync.m8n8k4.f64 -> DMMA.8x8x4)",
        "achieved": achieved, "peak": dmma_peak, "unit": "TFLOP/s", "frac": achieved / dmma_peak,
        "peak_source": "measured in this run (ca_measure_fp64_peak: DMMA loop); MEASURED_PEAKS.json has no FP64 entry",
        "dfma_peak": dfma_peak, "traffic": traffic,
        "traffic_source": "stored ncu --set full capture (profiles/ncu_traffic.json, profiles/r02_rhs307_final_ncu_raw.csv): "
                          "dram__bytes_read.sum + dram__bytes_write.sum of one launch scaled to this launch's states; not re-measured in this run",
        "kernel": KERNEL_NAMES.get(model.kernel_path(n)[0], model.kernel_path(n)[0]), "kernel_ms": kernel_ms,
        "algorithmic_flops_per_state": flops_state, "executed_fma_per_state": info["padded_fma_per_state"],
        "algorithmic_bytes_per_state": bytes_state,
        "hbm_view": {"achieved": bytes_state * n / (kernel_ms * 1e-3) / 1e9, "peak": hbm_peak,
                     "peak_source": hbm_peak_source, "unit": "GB/s",
                     "frac": bytes_state * n / (kernel_ms * 1e-3) / 1e9 / hbm_peak,
                     "note": "arithmetic intensity ~340 flop/B: this config is FP64-bound, not HBM-bound"},
    }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n), "m": m, "nnz_N": int(ijk.shape[0]), "states_per_gpu": n,
                   "l2": "inputs (2.46 GB/GPU at the default size) exceed L2; no flush needed",
                   "layout": "nstates x m column-major (state index fastest)"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * m * n, "d2h_bytes_per_step": 8 * m * n,
                "steps": e2e_steps, "path": "ca_model_rhs_batched (host pointers, page-locked, chunks of whole kernel waves on three streams)",
                "agrees_with_device_path": e2e_check, "copy_ceiling": e2e_ceiling,
                "frac_of_copy_ceiling": e2e_value / ceiling, "frac_of_device_rate": e2e_value / value},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "model_build": {"seconds_total": build_s, "assembly_device_seconds": model.assembly_stats["device_seconds"],
                        "gather_seconds": model.assembly_stats.get("gather_seconds", 0.0),
                        "library_comm_init_seconds": comm_init_s,
                        "model_create": model.build_info(), "world": world,
                        "note": "ODEModel(alpha,gamma,J,K,L,H): index set + basis tables + GPU assembly (+ library-side NCCL "
                                "all-gather) + device model build (B^-1 folding, symmetrisation, DMMA packing on the GPU); "
                                "first call in the process (first launches, allocations); the NCCL communicator set-up is timed separately"},
        "kernel_path": model.kernel_path(n)[0],
    }

    if not args.no_extras:
        del X, OUT
        torch.cuda.empty_cache()
        dense = dense_assembly_case(ca, H, world, rank, max_over_ranks, barrier, dmma_peak)
        c4 = large_basis_case(ca, H, (5, 5, 16), world, rank, max_over_ranks, barrier, dmma_peak, hbm_peak)
        c5 = large_basis_case(ca, H, (8, 8, 20), world, rank, max_over_ranks, barrier, dmma_peak, hbm_peak)
        if rank == 0:
            line["roofline_assembly"] = dense
            line["config4"] = c4
            line["config5"] = c5
    if rank == 0 and world == 1 and not args.no_extras:
        line["assembly"] = assembly_sweep(ca, H)
        line["roofline_assembly_mixed"] = mixed_basis_case(ca, H, dmma_peak)
        line["roofline_hbm_case"] = hbm_case(ca, H, hbm_peak)
        line["hookstep"] = hookstep_cases(ca, H)
        line["cpu_baseline"] = cpu_baseline(model, args)
        line["cpu_baseline_assembly"] = cpu_baseline_assembly()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def dense_assembly_case(ca, H, world, rank, max_over_ranks, barrier, dmma_peak):
    """BASELINE.json configs[3]: Q-tensor assembly of the ~1000-mode basis (J,K,L = 5,5,16) by dense quadrature on the
    16 x 33 x 16 tensor grid, sharded by output mode over the ranks (the FP64-roofline formulation of the assembly;
    the separable path gives the same tensor in ~25 ms, see `assembly`)."""
    ijkl = ca.basisIndices(5, 5, 16, H)
    m = ijkl.shape[0]
    bounds = ca.row_partition(m, world, 128)          # whole 128-row output tiles per rank
    barrier()
    t0 = time.perf_counter()
    slab = ca.AssemblySlab(ALPHA, GAMMA, ijkl, False, bounds[rank], bounds[rank + 1], dense=True, grid=(16, 33, 16))
    barrier()
    wall = max_over_ranks(time.perf_counter() - t0)
    info = slab.dense_info()
    sec = max_over_ranks(info["contract_seconds"])
    flops_total = 2.0 * m * m * m * 3 * 16 * 33 * 16
    nnz = slab.nnz
    del slab
    return {"workload": "dense quadrature assembly of N, m=999 (J,K,L=5,5,16), grid 16x33x16, rows sharded over ranks",
            "bound": "tensor", "pipe": "fp64 DMMA", "achieved": flops_total / sec / 1e12 / world, "peak": dmma_peak,
            "unit": "TFLOP/s per GPU", "frac": flops_total / sec / 1e12 / world / dmma_peak,
            "contract_seconds": sec, "wall_seconds": wall, "algorithmic_flops": flops_total, "n_gpus": world,
            "nnz_rank0": int(nnz), "kernel": "dense_contract_ring_kernel"}


NVLINK_BUSBW_MEASURED = 725.0   # GB/s, all-gather bus bandwidth measured on this pool's 8 x B200 boxes


def large_basis_case(ca, H, jkl, world, rank, max_over_ranks, barrier, dmma_peak, hbm_peak, nstates=9472):
    """BASELINE.json configs[3] / [4] end to end at N GPUs: assembly sharded by output mode through the library's own
    NCCL communicator (one grouped all-gather), device model build from the gathered operator, then ensemble f and
    Jacobian actions (the Newton-Krylov evaluation) with states sharded over the ranks, and single-state f / Jacobian
    action (HBM-bound stream of the operator)."""
    import numpy as np
    import torch
    ijkl = ca.basisIndices(*jkl, H)
    m = ijkl.shape[0]
    passes = []
    for rep in range(2):        # the first pass pays for NCCL channel set-up and first-touch allocations
        barrier()
        t0 = time.perf_counter()
        slab, st = ca.assemble_full(ALPHA, GAMMA, ijkl)
        barrier()
        t1 = time.perf_counter()
        model = ca.ModelOperators.from_assembly(slab)
        x1 = torch.zeros(m, dtype=torch.float64, device="cuda")
        model.f(x1, R)
        barrier()
        t2 = time.perf_counter()
        bi = model.build_info()
        passes.append({"assembly_wall_seconds": max_over_ranks(t1 - t0), "slab_device_seconds": max_over_ranks(st["device_seconds"]),
                       "all_gather_seconds": max_over_ranks(st["gather_seconds"]), "all_gather_bytes": st["gather_bytes"],
                       "model_create_seconds": max_over_ranks(bi["seconds"]["total"]),
                       "model_create_phases_rank0": {k: round(v, 4) for k, v in bi["seconds"].items()},
                       "time_to_first_eval_seconds": max_over_ranks(t2 - t0)})
        nnz = slab.nnz
        if rep == 0:
            del model, slab
    gs = passes[1]["all_gather_seconds"]
    busbw = passes[1]["all_gather_bytes"] * (world - 1) / world / gs / 1e9 if world > 1 and gs > 0 else None
    info = model.info()
    gen = torch.Generator(device="cuda").manual_seed(SEED + 100 + rank)
    X = (0.1 * torch.randn((m, nstates), dtype=torch.float64, device="cuda", generator=gen)).t()
    V = torch.randn((m, nstates), dtype=torch.float64, device="cuda", generator=gen).t()

    def timed(fn, reps):
        fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / reps

    rhs_ms = timed(lambda: model.f(X, R), 2)
    jvp_ms = timed(lambda: model.Df_action(X, V, R), 2)
    x0, v0 = X[0].contiguous(), V[0].contiguous()
    rhs1_ms = timed(lambda: model.f(x0, R), 10)
    jvp1_ms = timed(lambda: model.Df_action(X[:1], V[:1], R), 10)
    # ---- solvers at this size: Newton-hookstep with the device LU from the prolonged 44-mode equilibrium (replicated
    # on every rank), then one continuation step in Re by preconditioned Newton-Krylov, Jacobian actions sharded by
    # row slab over the ranks
    solvers = None
    try:
        golden = os.path.join(ROOT, "tests", "golden", "xeq1projection-Re200-2-2-3-44d.asc")
        x44g = np.array([float(t) for t in open(golden).read().split("\n") if t.strip() and not t.startswith("%")])
        ijkl44 = ca.basisIndices(2, 2, 3, H)
        xg = ca.changebasis(x44g, ijkl44, ijkl)   # the shipped DNS projection (44 modes), zero-padded to this basis
        barrier()
        t0 = time.perf_counter()
        xe, okh, hlog = ca.hookstepsolve(model, R, xg, ca.SearchParams(Nnewton=40), return_log=True)
        barrier()
        hook_wall = max_over_ranks(time.perf_counter() - t0)
        comm = ca.library_comm_for(None) if world > 1 else None
        xk, okk, klog = ca.newton_krylov(model, R + 5.0, xe, ftol=1e-10, max_newton=30, comm=comm)
        barrier()
        solvers = {"hookstep": {"converged": bool(okh), "newton_steps": hlog["newton_steps"], "device_seconds": hlog["device_seconds"],
                                "wall_seconds": hook_wall, "from": "the reference's 44-mode DNS projection prolonged with changebasis, Re = 200",
                                "solver": "global-memory Newton-hookstep: Jacobian by ensemble action, blocked LU on the device"},
                   "newton_krylov": {"converged": bool(okk), "newton_steps": klog["newton_steps"], "jacobian_actions": klog["jvps"],
                                     "f_evals": klog["f_evals"], "device_seconds": klog["device_seconds"], "world": klog["world"],
                                     "ms_per_operator_pass": 1e3 * klog["device_seconds"] / max(klog["jvps"] + klog["f_evals"], 1),
                                     "what": "Re 200 -> 205 from the hookstep solution; GMRES preconditioned with L1 + L2/R; "
                                             "row-slab sharded Jacobian action + one all-gather of m doubles per Krylov step at N > 1"}}
    except Exception as exc:   # a solver hiccup must not lose the throughput numbers of the line
        solvers = {"error": str(exc)[:300]}
    fma = info["padded_fma_per_state"]
    exec_tf = 2.0 * fma * nstates / (rhs_ms * 1e-3) / 1e12
    stream_bytes = 12 * info["nnz_q_sym"] + 16 * m * m       # streamed operator: 12 B per entry + the dense linear part
    out = {"workload": f"m={m} (J,K,L={jkl[0]},{jkl[1]},{jkl[2]}): sharded assembly + all-gather + device model build + "
                       f"{nstates} states per GPU of f and Jacobian actions",
           "m": int(m), "nnz_N": int(nnz), "n_gpus": world, "first_pass": passes[0], "steady_pass": passes[1],
           "all_gather_busbw_GBs": busbw, "nvlink_busbw_measured_GBs": NVLINK_BUSBW_MEASURED,
           "all_gather_frac_of_measured": (busbw / NVLINK_BUSBW_MEASURED) if busbw else None,
           "collective": "ca_assemble_sharded: slab sizes (8-byte all-gather), then every piece of every slab broadcast in place "
                         "inside one ncclGroup (csrc/comm.cu)" if world > 1 else "none (one GPU)",
           "rhs_evals_per_sec": world * nstates / (rhs_ms * 1e-3), "jvp_per_sec": world * nstates / (jvp_ms * 1e-3),
           "rhs_ms": rhs_ms, "jvp_ms": jvp_ms, "kernel_path": model.kernel_path(nstates)[0],
           "executed_fma_per_state": fma, "rhs_executed_tflops_per_gpu": exec_tf, "rhs_executed_frac_of_dmma_peak": exec_tf / dmma_peak,
           "jvp_executed_frac_of_dmma_peak": 2.0 * fma * nstates / (jvp_ms * 1e-3) / 1e12 / dmma_peak,
           "solvers": solvers,
           "single_state": {"rhs_ms": rhs1_ms, "jvp_ms": jvp1_ms, "bound": "hbm", "algorithmic_bytes": stream_bytes,
                            "rhs_GBs": stream_bytes / (rhs1_ms * 1e-3) / 1e9,
                            "rhs_frac_of_hbm": stream_bytes / (rhs1_ms * 1e-3) / 1e9 / hbm_peak}}
    del model, slab, X, V
    torch.cuda.empty_cache()
    return out


def mixed_basis_case(ca, H, dmma_peak):
    """The dense quadrature path on a basis that does not factorise in x, y, z -- Phi = Psi C, C = I + 0.1 G,
    G_ij ~ N(0,1), seed 20251029, on top of mode scales s_n ~ U(0.5, 2): one 128-row slab of the
    999-mode basis on the 16 x 33 x 16 grid.  The separable path cannot assemble such a basis at all."""
    import numpy as np
    ijkl = ca.basisIndices(5, 5, 16, H)
    m = ijkl.shape[0]
    rng = np.random.default_rng(SEED)
    s = rng.uniform(0.5, 2.0, m)
    Cm = np.eye(m) + 0.1 * rng.standard_normal((m, m))
    t0 = time.perf_counter()
    slab = ca.AssemblySlab(ALPHA, GAMMA, ijkl, False, 384, 512, dense=True, grid=(16, 33, 16), mode_scale=s, mix=Cm)
    wall = time.perf_counter() - t0
    info = slab.dense_info()
    tf = info["contract_flops"] / info["contract_seconds"] / 1e12
    out = {"workload": "dense quadrature assembly of N on a mixed random-coefficient basis Phi = (Psi diag(s)) C, m=999, rows 384:512, grid 16x33x16",
           "bound": "tensor", "pipe": "fp64 DMMA", "achieved": tf, "peak": dmma_peak, "unit": "TFLOP/s", "frac": tf / dmma_peak,
           "contract_seconds": info["contract_seconds"], "wall_seconds": wall, "nnz_slab": int(slab.nnz),
           "dense_fraction_of_slab": slab.nnz / (128.0 * m * m), "kernel": "dense_contract_ring_kernel"}
    del slab
    return out


def hookstep_cases(ca, H):
    """BASELINE.json configs[0] / [1]: the Newton-hookstep search of the find-eqb / continue-eqb notebooks, device
    resident, beside the same search on the host (oracle/hookstep.py: Hookstep.jl restated in numpy on the oracle's
    CPU model -- LAPACK solves, one core)."""
    import numpy as np
    from oracle.hookstep import hookstepsolve as ref_solve
    golden = os.path.join(ROOT, "tests", "golden")
    known = json.load(open(os.path.join(golden, "reference_known_answers.json")))

    def read_asc(name):
        return np.array([float(t) for t in open(os.path.join(golden, name)).read().split("\n") if t.strip() and not t.startswith("%")])

    out = []
    x44 = None
    for name, jkl, guess in (("17d", (1, 1, 3), np.array(known["m17_unnormalized"]["xguess"])),
                             ("44d", (2, 2, 3), read_asc(known["m44"]["guess_file"])), ("307", (3, 3, 12), None)):
        model = ca.ODEModel(ALPHA, GAMMA, *jkl, H)
        if guess is None:
            guess = ca.changebasis(x44, small_ijkl, model.ijkl)
        params = ca.SearchParams(Nnewton=30)
        ca.hookstepsolve(model, R, guess, params)                                   # warm-up (lazy structures, first launch)
        t0 = time.perf_counter()
        x, ok, log = ca.hookstepsolve(model, R, guess, params, return_log=True)
        wall = time.perf_counter() - t0
        if name == "44d":
            x44, small_ijkl = x, model.ijkl
        from oracle.cref import CRefModel
        ref = CRefModel(model.B, model.A1, model.A2, model.ijk, model.val)
        from oracle.hookstep import SearchParams as RP
        Bf = np.linalg.inv(model.B)
        fcpu = lambda v: ref.f(v, R)
        Dfcpu = lambda v: Bf @ (model.A1 + model.A2 / R + ref.derivative(v))
        t0 = time.perf_counter()
        xr, okr = ref_solve(fcpu, Dfcpu, guess, RP(Nnewton=30))
        cpu = time.perf_counter() - t0
        entry = {"model": name, "m": int(model.m), "converged": bool(ok), "newton_steps": log["newton_steps"],
                 "device_seconds": log["device_seconds"], "wall_seconds": wall,
                 "solver": "one persistent CTA (shared memory)" if model.m <= 95 else "global-memory solver (blocked LU, csrc/hookstep_large.cu)",
                 "cpu_baseline": {"seconds": cpu, "cores": 1, "kind": "port",
                                  "what": "Hookstep.jl:109-352 restated in numpy, f / derivative by the reference's COO loops in C, LAPACK solves"},
                 "agrees_with_cpu": float(np.linalg.norm(x - xr) / np.linalg.norm(xr))}
        if model.m <= 95:   # batches of independent searches: what the one-CTA solver is for
            nb = 128
            Rs = np.linspace(R - 5, R + 5, nb)
            t0 = time.perf_counter()
            X, good = ca.hookstepsolve(model, Rs, np.tile(x, (nb, 1)), params)
            entry["batch_of_128"] = {"wall_seconds": time.perf_counter() - t0, "converged": int(good.sum()),
                                     "seconds_per_search": (time.perf_counter() - t0) / nb}
        out.append(entry)
        del model
    return out


def cpu_baseline_assembly():
    """The reference's assembly algorithm on one host core (Julia unavailable): the oracle's restatement of
    ODEModels.jl:31-51 at the sizes it finishes in seconds, an m^3 extrapolation (labelled) beyond, and the oracle's
    vectorised exact-table tier at m = 307."""
    from oracle.basis import basisIndices, halfbox_symmetries
    from oracle.fastfloat import FloatTables, assemble_N_float
    from oracle.model import ODEModel as RefModel
    import numpy as np
    sx, sy, sz, tx, tz = halfbox_symmetries()
    Hs = [sx * sy * sz, sz * tx * tz]
    rows, per = [], []
    for jkl in ((1, 1, 3), (1, 2, 3), (2, 2, 3)):
        t0 = time.perf_counter()
        mdl = RefModel(ALPHA, GAMMA, *jkl, Hs)
        dt = time.perf_counter() - t0
        mm = int(mdl.B.shape[0])
        rows.append({"m": mm, "seconds": dt, "measured": True})
        per.append(dt / mm ** 3)
    c = float(np.median(per))
    for mm in (307, 999, 2962):
        rows.append({"m": mm, "seconds": c * mm ** 3, "measured": False, "how": "extrapolated ~ m^3 from the measured sizes"})
    ijkl = np.array(basisIndices(3, 3, 12, Hs))
    t0 = time.perf_counter()
    assemble_N_float(ALPHA, GAMMA, ijkl, tables=FloatTables(ALPHA, GAMMA, ijkl))
    vec = time.perf_counter() - t0
    return {"cores": 1, "kind": "port", "rows": rows,
            "vectorised_tier_m307_seconds": vec,
            "what": "oracle/model.py (the reference's m^2 + m^3 inner-product loops, ODEModels.jl:31-51) on one core; "
                    "oracle/fastfloat.py is the oracle's numpy-vectorised separable formulation"}


def assembly_sweep(ca, H):
    """Q-assembly seconds vs basis size (second half of BASELINE.json's metric) on random-coefficient bases
    (Psi_n -> s_n Psi_n, s_n ~ U(0.5, 2), seed 20251029: SURVEY.md 8d), and what a user waits for: assembly wall time
    plus the device model build (ca_model_create_from_assembly)."""
    import numpy as np
    import torch
    out = []
    rng = np.random.default_rng(SEED)
    for jkl in [(1, 1, 3), (2, 2, 3), (3, 3, 12), (5, 5, 16), (8, 8, 20)]:
        ijkl = ca.basisIndices(*jkl, H)
        s = rng.uniform(0.5, 2.0, ijkl.shape[0])
        ca.AssemblySlab(ALPHA, GAMMA, ijkl, mode_scale=s)        # warm-up (first launch, allocator)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        slab = ca.AssemblySlab(ALPHA, GAMMA, ijkl, mode_scale=s)
        t1 = time.perf_counter()
        model = ca.ModelOperators.from_assembly(slab)
        t2 = time.perf_counter()
        out.append({"JKL": list(jkl), "m": int(slab.m), "nnz": int(slab.nnz), "basis": "random-coefficient (mode scales)",
                    "device_seconds": slab.device_seconds, "wall_seconds_incl_tables": t1 - t0,
                    "model_create_seconds": t2 - t1, "wall_seconds_incl_model": t2 - t0,
                    "model_create_phases": model.build_info()["seconds"]})
        del model, slab
        torch.cuda.empty_cache()
    return out


def hbm_case(ca, H, hbm_peak):
    """The HBM-bound regime of the same kernel: 17-mode model, 10^7 states."""
    import torch
    model = ca.ODEModel(ALPHA, GAMMA, 1, 1, 3, H)
    n, m = 10_000_000, model.m
    X = (0.1 * torch.randn((m, n), dtype=torch.float64, device="cuda")).t()
    OUT = ca.ensemble_empty(n, m)
    for _ in range(3):
        model.f(X, R, out=OUT)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        model.f(X, R, out=OUT)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    gbs = 16 * m * n / (ms * 1e-3) / 1e9
    path, _, jitlog = model.kernel_path(n)
    return {"workload": "ensemble RHS, 17-mode model (J,K,L=1,1,3), 10^7 states", "bound": "hbm", "kernel": path,
            "note": "m <= 48: model-specialised straight-line kernel (NVRTC, sm_100a), states through a per-thread cp.async ring",
            "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
            "evals_per_sec": n / (ms * 1e-3), "kernel_ms": ms, "algorithmic_bytes_per_state": 16 * m}


def cpu_baseline(model, args):
    """The reference's loops (oracle/csrc) on this box's host cores, bounded sample of the same workload."""
    import numpy as np
    from oracle.cref import CRefModel
    threads = os.cpu_count() or 1
    ref = CRefModel(model.B, model.A1, model.A2, model.ijk, model.val)
    X1 = host_states(256, model.m, SEED)
    t0 = time.perf_counter()
    ref.f_ensemble(X1, R, 1)
    single = 256 / (time.perf_counter() - t0)
    sample = args.cpu_sample or int(max(1024, min(65536, single * threads * 12)))
    X = host_states(sample, model.m, SEED)
    t0 = time.perf_counter()
    out = ref.f_ensemble(X, R, threads)
    dt = time.perf_counter() - t0
    got = model.f(X[:64], R)
    agree = float(np.linalg.norm(got - out[:64]) / np.linalg.norm(out[:64]))
    return {"value": sample / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{sample} of the states of the same 307-mode workload, pthreads over states",
            "single_thread_value": single, "gpu_vs_port_rel_diff_on_64_states": agree}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
