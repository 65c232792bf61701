"""Dense quadrature assembly (ca_assemble_dense) at a given basis: time, FP64 TFLOP/s, agreement with the separable path.
    python tools/dense_assembly.py 5,5,16 [Nx,Ny,Nz] [row_begin,row_end]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
ca = g.load_package()
J, K, L = (int(v) for v in sys.argv[1].split(","))
grid = tuple(int(v) for v in sys.argv[2].split(",")) if len(sys.argv) > 2 else (0, 0, 0)
sx, sy, sz, tx, tz = ca.halfbox_symmetries()
ijkl = ca.basisIndices(J, K, L, [sx * sy * sz, sz * tx * tz])
m = ijkl.shape[0]
r0, r1 = (int(v) for v in sys.argv[3].split(",")) if len(sys.argv) > 3 else (0, m)
dmma, dfma = ca.measure_fp64_peak()
t0 = time.perf_counter()
d = ca.AssemblySlab(1.0, 2.0, ijkl, False, r0, r1, dense=True, grid=grid)
wall = time.perf_counter() - t0
info = d.dense_info()
s = ca.AssemblySlab(1.0, 2.0, ijkl, False, r0, r1)
(di, dv), (si, sv) = d.coo_host(), s.coo_host()
out = {"JKL": [J, K, L], "m": m, "rows": [r0, r1], **info, "fp64_peak_dmma_tflops": dmma,
       "frac_of_fp64_peak": info["tflops"] / dmma, "total_device_seconds": d.device_seconds, "wall_seconds": wall,
       "nnz": int(d.nnz), "separable_device_seconds": s.device_seconds, "same_pattern": bool(np.array_equal(di, si)),
       "maxabs_rel_diff": float(np.max(np.abs(dv - sv)) / np.max(np.abs(sv))) if d.nnz == s.nnz else None}
print(json.dumps(out))
