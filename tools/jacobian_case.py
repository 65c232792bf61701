"""Dense Jacobians Df(x_s, R) of an ensemble (ca_model_jac_batched_dev): ms and achieved GB/s of the output stream.
    python tools/jacobian_case.py J,K,L nstates [J,K,L nstates ...]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
ca = g.load_package()
sx, sy, sz, tx, tz = ca.halfbox_symmetries()
args = sys.argv[1:]
for spec, ns in zip(args[0::2], args[1::2]):
    J, K, L = (int(v) for v in spec.split(","))
    n = int(ns)
    model = ca.ODEModel(1.0, 2.0, J, K, L, [sx * sy * sz, sz * tx * tz])
    m = model.m
    X = (0.1 * torch.randn((m, n), dtype=torch.float64, device="cuda")).t()
    for _ in range(2):
        D = model.Df(X, 200.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        D = model.Df(X, 200.0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    nbytes = 8 * (m * m + m) * n
    print(json.dumps({"JKL": [J, K, L], "m": m, "nstates": n, "ms": round(ms, 4), "jacobians_per_sec": round(n / ms * 1e3),
                      "algorithmic_bytes": nbytes, "gbs": round(nbytes / ms / 1e6, 1)}))
    del D, X
