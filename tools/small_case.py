"""Ensemble RHS of a small model through the generated (NVRTC) kernel: ms, GB/s and DFMA TFLOP/s.
    python tools/small_case.py J,K,L nstates[,nstates...]      (CA_SMALL_SHAPE=threads,stages,ctas to tune)"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
ca = g.load_package()
sx, sy, sz, tx, tz = ca.halfbox_symmetries()
J, K, L = (int(v) for v in sys.argv[1].split(","))
model = ca.ODEModel(1.0, 2.0, J, K, L, [sx * sy * sz, sz * tx * tz])
m = model.m
info = model.info()
flops = 2 * (info["nnz_q_sym"] + info["nnz_lin"]) + info["npairs"]
for n in (int(v) for v in sys.argv[2].split(",")):
    X = (0.1 * torch.randn((m, n), dtype=torch.float64, device="cuda")).t()
    OUT = ca.ensemble_empty(n, m)
    for _ in range(3):
        model.f(X, 200.0, out=OUT)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        model.f(X, 200.0, out=OUT)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(json.dumps({"shape": os.environ.get("CA_SMALL_SHAPE", "default"), "m": m, "nstates": n, "path": model.kernel_path(n)[0],
                      "ms": round(ms, 4), "gbs": round(16 * m * n / ms / 1e6, 1), "tflops": round(flops * n / ms / 1e9, 2)}))
    del X, OUT
