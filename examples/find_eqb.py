"""notebooks/find-eqb.ipynb on the B200 path (BASELINE.json configs[0]): build the 17-mode normalised Nagata model,
find the equilibrium at Re = 200 with the device-resident Newton-hookstep, print its leading eigenvalues, and compare
with the values stored in the reference notebook."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g

ca = g.load_package()
sx, sy, sz, tx, tz = ca.halfbox_symmetries()
H = [sx * sy * sz, sz * tx * tz]                       # generators of the Nagata-equilibrium subspace
model = ca.ODEModel(1.0, 2.0, 1, 1, 3, H, normalize=True, verbose=True)
print("m =", len(model), " cond(B) =", np.linalg.cond(model.B), "(notebook: 22.597316672026967)")
xguess = np.array([0.117, -0.0402, 0.201, -0.00779, -0.0158, 0.00719, -0.00551, 0.0126, -0.0281, -0.00619, -0.0102,
                   -0.0121, -0.0144, -0.0222, -0.00386, 0.01876, 0.0229])
x, ok, log = ca.hookstepsolve(model, 200.0, xguess, return_log=True)
print("success =", ok, " |f(x)|/|x| =", np.linalg.norm(model.f(x, 200.0)) / np.linalg.norm(x),
      " |x| =", np.linalg.norm(x), "(notebook: 0.3340773241869399)")
print("newton steps:", log["newton_steps"], " residuals:", ["%.2e" % r for r in log["residual"]],
      " device time: %.3f ms" % (1e3 * log["device_seconds"]))
lam = np.linalg.eigvals(model.Df(x, 200.0))
lam = lam[np.argsort(-lam.real)]
print("leading eigenvalues:", lam[:3], "(notebook: 0.09395778499834803 -/+ 0.1638027353197821im, 0.04460084133030933)")
print("shear rate I =", model.shear(x), " dissipation D =", model.dissipation(x))
