"""SparseBilinear COO container and its two kernels (test oracle only).

Follows /root/reference/src/SparseBilinear.jl: constructor checks :7-11, dense->COO in
lexicographic (i,j,k) order keeping ``!= 0`` :24-49, ``N(x,y)`` :56-65, ``N(x)`` :70,
``derivative(N,x)`` :75-91.  Indices are 1-based int64 exactly as ``N.ijk``.
The accumulation order (one running sum per output, in COO order) is the reference's.
"""
from __future__ import annotations

import numpy as np


class SparseBilinear:
    def __init__(self, ijk, val, m: int):
        ijk = np.asarray(ijk, dtype=np.int64)
        if ijk.ndim != 2 or ijk.shape[1] != 3:
            raise ValueError("ijk must have three columns")
        if ijk.shape[0] != len(val):
            raise ValueError("ijk and val must have same number of rows")
        self.ijk = ijk
        self.val = val if isinstance(val, list) else np.asarray(val, dtype=np.float64)
        self.m = int(m)

    @classmethod
    def from_dense(cls, N):
        """SparseBilinear.jl:24-49 (two passes; C-order nonzero() is lexicographic i,j,k)."""
        N = np.asarray(N)
        idx = np.nonzero(N != 0)
        ijk = np.stack(idx, axis=1).astype(np.int64) + 1
        val = N[idx]
        if N.dtype == object:
            val = list(val)
        return cls(ijk, val, N.shape[0])

    @property
    def nnz(self) -> int:
        return self.ijk.shape[0]

    def __call__(self, x, y=None):
        """rtn[i] += val * x[j] * y[k] in COO order  --  SparseBilinear.jl:56-65."""
        if y is None:
            y = x
        exact = isinstance(self.val, list)
        out = [0] * self.m if exact else np.zeros(self.m)
        i = self.ijk[:, 0] - 1
        j = self.ijk[:, 1] - 1
        k = self.ijk[:, 2] - 1
        if exact or self.nnz < 4096:
            for r in range(self.nnz):
                out[i[r]] = out[i[r]] + self.val[r] * x[j[r]] * y[k[r]]
            return out
        # large float case: the same products, summed per output by np.add.at in COO order
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        np.add.at(out, i, self.val * x[j] * y[k])
        return out

    def derivative(self, x):
        """DN[i,j] += val*x[k]; DN[i,k] += val*x[j]  --  SparseBilinear.jl:75-91."""
        exact = isinstance(self.val, list)
        i = self.ijk[:, 0] - 1
        j = self.ijk[:, 1] - 1
        k = self.ijk[:, 2] - 1
        if exact:
            DN = [[0] * self.m for _ in range(self.m)]
            for r in range(self.nnz):
                DN[i[r]][j[r]] += self.val[r] * x[k[r]]
                DN[i[r]][k[r]] += self.val[r] * x[j[r]]
            return DN
        DN = np.zeros((self.m, self.m))
        x = np.asarray(x, dtype=np.float64)
        if self.nnz < 4096:
            for r in range(self.nnz):
                DN[i[r], j[r]] += self.val[r] * x[k[r]]
                DN[i[r], k[r]] += self.val[r] * x[j[r]]
            return DN
        np.add.at(DN, (i, j), self.val * x[k])
        np.add.at(DN, (i, k), self.val * x[j])
        return DN


def eval_batched(N: SparseBilinear, X: np.ndarray) -> np.ndarray:
    """N(x_s) for every row s of X (nstates x m); vectorised over states, COO order kept."""
    X = np.asarray(X, dtype=np.float64)
    out = np.zeros((X.shape[0], N.m))
    i = N.ijk[:, 0] - 1
    j = N.ijk[:, 1] - 1
    k = N.ijk[:, 2] - 1
    # chunk the nnz axis so the temporary stays small
    step = max(1, (1 << 24) // max(1, X.shape[0]))
    for a in range(0, N.nnz, step):
        b = min(N.nnz, a + step)
        contrib = N.val[a:b][None, :] * X[:, j[a:b]] * X[:, k[a:b]]
        np.add.at(out.T, i[a:b], contrib.T)
    return out
