"""Large-basis check (BASELINE.json configs[3] and [4]): GPU assembly at m ~ 1000 / ~3000, model packing,
ensemble RHS and the matrix-free Jacobian action.  Prints one JSON line per basis.  Usage:
    python tools/large_basis.py 5,5,16 [8,8,20] [--nstates 4096]"""
import json
import sys
import time

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np
import torch

import __graft_entry__ as g

ca = g.load_package()
sx, sy, sz, tx, tz = ca.halfbox_symmetries()
H = [sx * sy * sz, sz * tx * tz]
nstates_list = [4096]
args = [a for a in sys.argv[1:] if not a.startswith("--")]
if "--nstates" in sys.argv:
    tok = sys.argv[sys.argv.index("--nstates") + 1]
    nstates_list = [int(v) for v in tok.split(",")]
    args = [a for a in args if a != tok]
for spec in args:
    J, K, L = (int(v) for v in spec.split(","))
    t0 = time.perf_counter()
    model = ca.ODEModel(1.0, 2.0, J, K, L, H)
    build = time.perf_counter() - t0
    m = model.m
    info = model.info()
    for nstates in nstates_list:
        gen = torch.Generator(device="cuda").manual_seed(1)
        X = (0.1 * torch.randn((m, nstates), dtype=torch.float64, device="cuda", generator=gen)).t()
        V = torch.randn((m, nstates), dtype=torch.float64, device="cuda", generator=gen).t()
        out = {"nstates": nstates, "JKL": [J, K, L], "m": m, "nnz": int(model.ijk.shape[0]), "build_seconds": build,
               "assembly_device_seconds": model.assembly_stats["device_seconds"], **info}
        out["stream_bytes_per_pass"] = 12 * info["nnz_q_sym"] + 16 * m * m
        for name, fn in (("rhs", lambda: model.f(X, 200.0)), ("jvp", lambda: model.Df_action(X, V, 200.0))):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                r = fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            out[name + "_ms"] = ms
            out[name + "_per_sec"] = nstates / (ms * 1e-3)
            out[name + "_executed_tflops"] = 2 * info["padded_fma_per_state"] * nstates / (ms * 1e-3) / 1e12
            if nstates <= 16:  # streaming kernel: one pass of the operator per <= 8 states
                out[name + "_stream_gbs"] = out["stream_bytes_per_pass"] * ((nstates + 7) // 8) / (ms * 1e-3) / 1e9
        if nstates == nstates_list[0] and m <= 1200:   # dense Jacobian Df(x,R): sliced-ELL stream (K3)
            x1 = X[0].contiguous()
            model.Df(x1, 200.0)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                D = model.Df(x1, 200.0)
            e1.record()
            torch.cuda.synchronize()
            out["jac_ms"] = e0.elapsed_time(e1) / 3
            out["jac_vs_jvp"] = float(torch.linalg.norm(D @ V[0] - model.Df_action(X[:1], V[:1], 200.0)[0]) /
                                      torch.linalg.norm(D @ V[0]))
        # consistency: JVP against a central difference of the RHS on a few states
        eps = 1e-6
        fd = (model.f((X[:64] + eps * V[:64]).t().contiguous().t(), 200.0) -
              model.f((X[:64] - eps * V[:64]).t().contiguous().t(), 200.0)) / (2 * eps)
        jv = model.Df_action(X, V, 200.0)[:64]
        out["jvp_vs_central_difference"] = float(torch.linalg.norm(fd - jv) / torch.linalg.norm(jv))
        print(json.dumps(out), flush=True)
    del model
