"""Device model build at a given basis size: assembly on the GPU, ca_model_create_dev, phase times, first evaluations.
usage: python tools/model_build_case.py J K L [--host] [--check]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402


def main():
    import torch
    ca = g.load_package()
    J, K, L = (int(a) for a in sys.argv[1:4])
    sx, sy, sz, tx, tz = ca.halfbox_symmetries()
    ijkl = ca.basisIndices(J, K, L, [sx * sy * sz, sz * tx * tz])
    m = ijkl.shape[0]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lin, idx, val, st = ca.assemble_device(1.0, 2.0, ijkl)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    dev = ca.ModelOperators.from_device(lin[0].t(), lin[1].t(), lin[2].t(), idx, val)
    t2 = time.perf_counter()
    out = {"m": m, "nnz": int(val.shape[0]), "assembly_wall_s": t1 - t0, "assembly_device_s": st["device_seconds"],
           "model_create_wall_s": t2 - t1, "build": dev.build_info(), "info": dev.info()}
    rng = np.random.default_rng(0)
    n = 9472
    X = torch.from_numpy(0.1 * rng.standard_normal((m, n))).cuda().t()
    V = torch.from_numpy(rng.standard_normal((m, n))).cuda().t()
    for name, fn in (("rhs", lambda: dev.f(X, 200.0)), ("jvp", lambda: dev.Df_action(X, V, 200.0)),
                     ("rhs1", lambda: dev.f(X[0].contiguous(), 200.0))):
        if name == "jvp" and os.environ.get("CA_SKIP_JVP"):
            out["jvp_ms"] = float("nan")
            continue
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out[name + "_ms"] = e0.elapsed_time(e1) / 3
    if "--host" in sys.argv:
        os.environ["CA_BUILD_HOST"] = "1"
        t3 = time.perf_counter()
        host = ca.ModelOperators(dev.B, dev.A1, dev.A2, dev.ijk, dev.val)
        out["host_build_wall_s"] = time.perf_counter() - t3
        out["same_pack"] = host.digest() == dev.digest()
        out["same_rhs"] = bool(torch.equal(host.f(X, 200.0), dev.f(X, 200.0)))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
