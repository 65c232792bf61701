"""Ensemble RHS of the 307-mode model (the headline workload at a fifth of its size): ms per launch.
    python tools/rhs307_case.py [nstates]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
ca = g.load_package()
sx, sy, sz, tx, tz = ca.halfbox_symmetries()
model = ca.ODEModel(1.0, 2.0, 3, 3, 12, [sx * sy * sz, sz * tx * tz])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
m = model.m
X = (0.1 * torch.randn((m, n), dtype=torch.float64, device="cuda")).t()
OUT = ca.ensemble_empty(n, m)
for _ in range(3):
    model.f(X, 200.0, out=OUT)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    model.f(X, 200.0, out=OUT)
e1.record()
torch.cuda.synchronize()
print(json.dumps({"m": m, "nstates": n, "ms": e0.elapsed_time(e1) / 10, "path": model.kernel_path(n)[0], "info": model.info(),
                  "lpt_w1": os.environ.get("CA_LPT_W1")}))
