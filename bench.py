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
    del Xh, Oh, Xh_np, Oh_np

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
                "agrees_with_device_path": e2e_check},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "model_build": {"seconds_total": build_s, "assembly_device_seconds": model.assembly_stats["device_seconds"],
                        "world": world, "note": "index set + GPU assembly (+ all-gather) + host B^-1 folding + DMMA packing"},
        "kernel_path": model.kernel_path(n)[0],
    }

    if not args.no_extras:
        dense = dense_assembly_case(ca, H, world, rank, max_over_ranks, barrier, dmma_peak)
        if rank == 0:
            line["roofline_assembly"] = dense
    if rank == 0 and world == 1 and not args.no_extras:
        line["assembly"] = assembly_sweep(ca, H)
        line["roofline_hbm_case"] = hbm_case(ca, H, hbm_peak)
        line["cpu_baseline"] = cpu_baseline(model, args)
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
            "nnz_rank0": int(nnz), "kernel": "dense_contract_kernel"}


def assembly_sweep(ca, H):
    """Q-assembly seconds vs basis size (second half of BASELINE.json's metric)."""
    out = []
    for jkl in [(1, 1, 3), (2, 2, 3), (3, 3, 12), (5, 5, 16)]:
        ijkl = ca.basisIndices(*jkl, H)
        ca.AssemblySlab(ALPHA, GAMMA, ijkl)                      # warm-up (first launch, allocator)
        t0 = time.perf_counter()
        slab = ca.AssemblySlab(ALPHA, GAMMA, ijkl)
        wall = time.perf_counter() - t0
        out.append({"JKL": list(jkl), "m": int(slab.m), "nnz": int(slab.nnz),
                    "device_seconds": slab.device_seconds, "wall_seconds_incl_tables": wall})
        del slab
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
