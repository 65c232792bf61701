"""f and Jacobian action of a large model on SMALL ensembles (fewer blocks of 32 states than SMs): ms per call.
    python tools/small_ensemble_case.py J K L n1 n2 ...      (CA_WIN_NO_SPLIT=1: one CTA per block of states walks all tiles)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
ca = g.load_package()
sx, sy, sz, tx, tz = ca.halfbox_symmetries()
J, K, L = (int(a) for a in sys.argv[1:4])
model = ca.ODEModel(1.0, 2.0, J, K, L, [sx * sy * sz, sz * tx * tz])
m = model.m
out = {"m": m, "split": not os.environ.get("CA_WIN_NO_SPLIT")}
for n in (int(a) for a in sys.argv[4:]):
    X = (0.1 * torch.randn((m, n), dtype=torch.float64, device="cuda")).t()
    V = torch.randn((m, n), dtype=torch.float64, device="cuda").t()
    for name, fn in (("f", lambda: model.f(X, 200.0)), ("jvp", lambda: model.Df_action(X, V, 200.0))):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record(); torch.cuda.synchronize()
        out[f"{name}_{n}_ms"] = round(e0.elapsed_time(e1) / 5, 3)
print(json.dumps(out))
