"""One Newton-hookstep search with the global-memory solver (m > 95) from the prolonged 44-mode DNS projection:
    python tools/hookstep_large_case.py J K L      (run under `ncu --metrics gpu__time_duration.sum` for the launch list)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
ca = g.load_package()
sx, sy, sz, tx, tz = ca.halfbox_symmetries()
H = [sx * sy * sz, sz * tx * tz]
GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
x44 = np.array([float(t) for t in open(os.path.join(GOLD, "xeq1projection-Re200-2-2-3-44d.asc")).read().split("\n")
                if t.strip() and not t.startswith("%")])
J, K, L = (int(a) for a in sys.argv[1:4])
small = ca.basisIndices(2, 2, 3, H)
model = ca.ODEModel(1.0, 2.0, J, K, L, H)
guess = ca.changebasis(x44, small, model.ijkl)
x, ok, log = ca.hookstepsolve(model, 200.0, guess, return_log=True)
print(json.dumps({"m": model.m, "ok": bool(ok), "newton_steps": log["newton_steps"], "device_seconds": log["device_seconds"],
                  "hooksteps": [int(h) for h in log["hooksteps"][:log["newton_steps"]]]}))
