"""notebooks/continue-eqb.ipynb on the B200 path (BASELINE.json configs[1]): the 27-mode (or 44-mode) model, the
equilibrium from the shipped DNS projection, then continuation in Re (natural-parameter sweep as one batched launch per
step, and pseudo-arclength around the saddle-node).  Usage: python examples/continue_eqb.py [27|44]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as g

ca = g.load_package()
which = sys.argv[1] if len(sys.argv) > 1 else "27"
JKL, fname = {"27": ((1, 2, 3), "xeq1projection-Re200-1-2-3-27d.asc"),
              "44": ((2, 2, 3), "xeq1projection-Re200-2-2-3-44d.asc")}[which]
sx, sy, sz, tx, tz = ca.halfbox_symmetries()
model = ca.ODEModel(1.0, 2.0, *JKL, [sx * sy * sz, sz * tx * tz], verbose=True)
xguess = ca.load(os.path.join(ROOT, "tests", "golden", fname))
print("|f(xguess, 200)| =", np.linalg.norm(model.f(xguess, 200.0)), "(notebook, 27d: 0.009524056134927641)")
x, ok = ca.hookstepsolve(model, 200.0, xguess)
print("equilibrium: |f|/|x| =", np.linalg.norm(model.f(x, 200.0)) / np.linalg.norm(x), " shear =", model.shear(x),
      "(notebook, 27d: 1.646052703437098)")
br = ca.pseudo_arclength_continuation(model, x, 200.0, ds=0.02, nsteps=300, direction=-1, R_max=300.0, ds_max=0.3)
k = int(np.argmin(br["R"]))
print(f"branch: {len(br['R'])} points, saddle-node near Re = {br['R'][k]:.2f} (shear {br['shear'][k]:.4f}), "
      f"max residual {br['residual'].max():.1e}")
for R, I in list(zip(br["R"], br["shear"]))[:: max(1, len(br["R"]) // 12)]:
    print(f"  Re = {R:8.3f}   I = {I:.5f}")
