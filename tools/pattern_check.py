"""Sparsity-pattern audit of the GPU assembly at a large basis against the oracle's vectorised tier, with the
entries on which they differ re-evaluated in exact rational arithmetic (oracle/exact.py)."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
from oracle.basis import basisIndices as ref_indices, halfbox_symmetries as ref_hs
from oracle.exact import assemble_N_exact
from oracle.fastfloat import FloatTables, assemble_N_float

ca = g.load_package()
J, K, L = (int(v) for v in sys.argv[1].split(","))
sx, sy, sz, tx, tz = ca.halfbox_symmetries()
ijkl = ca.basisIndices(J, K, L, [sx * sy * sz, sz * tx * tz])
slab = ca.AssemblySlab(1.0, 2.0, ijkl)
ijk, val = slab.coo_host()
m = ijkl.shape[0]
rows = list(range(0, m, max(1, m // 40)))
T = FloatTables(1, 2, ijkl)
fijk, fval = assemble_N_float(1, 2, ijkl, rows=rows, tables=T)
sel = np.isin(ijk[:, 0] - 1, rows)
gset = {tuple(r): v for r, v in zip(ijk[sel].tolist(), val[sel])}
fset = {tuple(r): v for r, v in zip(fijk.tolist(), fval)}
only_g = sorted(set(gset) - set(fset))
only_f = sorted(set(fset) - set(gset))
common = sorted(set(gset) & set(fset))
scale = max(abs(v) for v in fset.values())
maxdiff = max(abs(gset[k] - fset[k]) for k in common) / scale
out = {"JKL": [J, K, L], "m": m, "nnz_gpu": int(slab.nnz), "rows_checked": len(rows), "only_gpu": len(only_g),
       "only_oracle_float": len(only_f), "maxabs_rel_diff_common": maxdiff}
bad_rows = sorted({k[0] - 1 for k in only_g + only_f})[:6]
if bad_rows:
    eijk, evals = assemble_N_exact(1, 2, ijkl, rows=bad_rows)
    eset = {tuple(r): float(v) for r, v in zip(eijk.tolist(), evals)}
    out["disputed"] = [{"ijk": list(k), "gpu": gset.get(k), "oracle_float": fset.get(k), "exact": eset.get(k, 0.0)}
                       for k in (only_g + only_f) if k[0] - 1 in bad_rows][:20]
print(json.dumps(out))
