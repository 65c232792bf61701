"""Separable Galerkin assembly (K1b): device seconds and entry count per basis; CA_ASSEMBLE_NOPRUNE=1 scans every k.
    python tools/assembly_case.py 3,3,12 5,5,16 8,8,20"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
ca = g.load_package()
sx, sy, sz, tx, tz = ca.halfbox_symmetries()
H = [sx * sy * sz, sz * tx * tz]
for spec in sys.argv[1:]:
    J, K, L = (int(v) for v in spec.split(","))
    ijkl = ca.basisIndices(J, K, L, H)
    ca.AssemblySlab(1.0, 2.0, ijkl)
    best = None
    for _ in range(3):
        slab = ca.AssemblySlab(1.0, 2.0, ijkl)
        best = slab.device_seconds if best is None else min(best, slab.device_seconds)
        nnz = int(slab.nnz)
        del slab
    print(json.dumps({"JKL": [J, K, L], "m": int(ijkl.shape[0]), "nnz": nnz, "device_ms": round(1e3 * best, 4),
                      "prune": "CA_ASSEMBLE_NOPRUNE" not in os.environ}))
