"""The reference's bifurcation diagram (notebooks/continue-eqb.ipynb:654-668, 757-761): wall shear I of the lower-branch
Nagata equilibrium against Re, ODE models of growing size beside the DNS curve the reference ships
(notebooks/data/eq1ReD-DNS.asc = tests/golden/eq1ReD-DNS.asc).  Every equilibrium is found by the device-resident
Newton-hookstep (one-CTA solver up to 95 modes, global-memory solver beyond), continued in Re by natural continuation.
    python tools/dns_curve_case.py [JKL ...]   e.g.  1,2,3 2,2,3 3,3,4 3,3,12"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def dns_lower_branch():
    d = np.array([[float(t) for t in l.split("#")[0].split()] for l in open(os.path.join(GOLD, "eq1ReD-DNS.asc"))
                  if l.strip() and not l.startswith("#")])
    i = int(np.argmin(d[:, 0]))
    lower = d[: i + 1][::-1]
    return lambda R: float(np.interp(R, lower[:, 0], lower[:, 1]))


def main():
    ca = g.load_package()
    sx, sy, sz, tx, tz = ca.halfbox_symmetries()
    H = [sx * sy * sz, sz * tx * tz]
    jkls = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]] or [(1, 2, 3), (2, 2, 3), (3, 3, 4), (3, 3, 12)]
    dns = dns_lower_branch()
    Rs = np.array([200.0, 225.0, 250.0, 275.0, 300.0])
    x44 = np.array([float(t) for t in open(os.path.join(GOLD, "xeq1projection-Re200-2-2-3-44d.asc")).read().split("\n")
                    if t.strip() and not t.startswith("%")])
    base = ca.ODEModel(1.0, 2.0, 2, 2, 3, H)
    xb, ok = ca.hookstepsolve(base, 200.0, x44, ca.SearchParams(Nnewton=40))
    out = []
    for jkl in jkls:
        model = ca.ODEModel(1.0, 2.0, *jkl, H)
        xg = ca.changebasis(xb, base.ijkl, model.ijkl) if model.m >= 44 else xb[: model.m]
        row = {"JKL": list(jkl), "m": int(model.m), "Re": [], "I_model": [], "I_dns": [], "converged": []}
        x = xg
        for R in Rs:
            x, ok, log = ca.hookstepsolve(model, float(R), x, ca.SearchParams(Nnewton=40), return_log=True)
            row["Re"].append(float(R))
            row["I_model"].append(float(model.shear(x)))
            row["I_dns"].append(dns(float(R)))
            row["converged"].append(bool(ok))
        row["max_rel_dev_from_dns"] = float(np.max(np.abs(np.array(row["I_model"]) / np.array(row["I_dns"]) - 1)))
        out.append(row)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
