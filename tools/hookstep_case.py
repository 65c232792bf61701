"""Device time of one Newton-hookstep search (K4) for the reference's 17d / 27d / 44d guesses, and of a batch.
    python tools/hookstep_case.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
ca = g.load_package()
sx, sy, sz, tx, tz = ca.halfbox_symmetries()
H = [sx * sy * sz, sz * tx * tz]
GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def read_asc(name):
    return np.array([float(t) for t in open(os.path.join(GOLD, name)).read().split("\n") if t.strip() and not t.startswith("%")])


x17 = np.array([0.117, -0.0402, 0.201, -0.00779, -0.0158, 0.00719, -0.00551, 0.0126, -0.0281, -0.00619, -0.0102,
                -0.0121, -0.0144, -0.0222, -0.00386, 0.01876, 0.0229])
cases = [((1, 1, 3), True, x17), ((1, 2, 3), False, read_asc("xeq1projection-Re200-1-2-3-27d.asc")),
         ((2, 2, 3), False, read_asc("xeq1projection-Re200-2-2-3-44d.asc"))]
for jkl, norm, xg in cases:
    model = ca.ODEModel(1.0, 2.0, *jkl, H, normalize=norm)
    best = None
    for _ in range(3):
        x, ok, log = ca.hookstepsolve(model, 200.0, xg, return_log=True)
        best = log["device_seconds"] if best is None else min(best, log["device_seconds"])
    print(json.dumps({"m": len(model), "success": bool(ok), "newton_steps": log["newton_steps"], "f_evals": log["f_evals"],
                      "hooksteps": [int(h) for h in log["hooksteps"][:log["newton_steps"]]],
                      "device_ms": round(1e3 * best, 4), "residual": float(np.linalg.norm(model.f(x, 200.0)))}))
