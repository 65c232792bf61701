"""Small single-purpose runs for ncu captures of the secondary kernels.
    python tools/profile_cases.py small17 | stream999 | windowed999 | windowed999jvp"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
ca = g.load_package()
sx, sy, sz, tx, tz = ca.halfbox_symmetries()
H = [sx * sy * sz, sz * tx * tz]
case = sys.argv[1]
if case == "small17":
    model = ca.ODEModel(1.0, 2.0, 1, 1, 3, H)
    n = 10_000_000
elif case == "stream999":
    model = ca.ODEModel(1.0, 2.0, 5, 5, 16, H)
    n = 1
else:
    model = ca.ODEModel(1.0, 2.0, 5, 5, 16, H)
    n = 9472
m = model.m
X = (0.1 * torch.randn((m, n), dtype=torch.float64, device="cuda")).t()
OUT = ca.ensemble_empty(n, m)
if case == "windowed999jvp":
    V = torch.randn((m, n), dtype=torch.float64, device="cuda").t()
    for _ in range(4):
        OUT = model.Df_action(X, V, 200.0)
else:
    for _ in range(4):
        model.f(X, 200.0, out=OUT)
torch.cuda.synchronize()
print(case, model.kernel_path(n)[0], float(OUT.abs().max()))
