"""Few-state streaming kernel (K2s): single-state f and Jacobian action at a large basis, with the CTAs-per-SM hook.
    python tools/stream_case.py J,K,L"""
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
stream_bytes = 12 * (info["nnz_q_sym"]) + 16 * m * m
for ctas in ("", "1", "2", "3"):
    if ctas:
        os.environ["CA_STREAM_CTAS"] = ctas
    for n in (1, 8):
        X = (0.1 * torch.randn((m, n), dtype=torch.float64, device="cuda")).t()
        V = torch.randn((m, n), dtype=torch.float64, device="cuda").t()
        for name, fn in (("rhs", lambda: model.f(X, 200.0)), ("jvp", lambda: model.Df_action(X, V, 200.0))):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(json.dumps({"m": m, "ctas_per_sm": ctas or "max", "nstates": n, "op": name, "ms": round(ms, 4),
                              "stream_gbs": round(stream_bytes / ms / 1e6, 1)}))
