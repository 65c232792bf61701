"""Plain-text `.asc` I/O and basis changes -- mirror of src/BasisFunctions.jl:925-958 (ijkl2file, save) and
:987-1021 (basis_index_dict, changebasis), plus a reader for the format (the reference's notebooks define
their own, `myreaddlm`, continue-eqb.ipynb:48-54).  Host-side utilities: this is the wire format shared with the
channelflow tools (chflow-programs/basisfunctions.cpp:71-166) and with the shipped DNS projections."""
from __future__ import annotations

import numpy as np

from ._lib import CloudAtlasError


def _filename(filebase):
    return filebase if ".asc" in filebase else filebase + ".asc"


def _fmt(v):
    return repr(float(v))  # shortest round-trip representation, as Julia prints Float64


def ijkl2file(ijkl, filebase, comment_char="#"):
    """BasisFunctions.jl:925-936."""
    ijkl = np.asarray(ijkl)
    if ijkl.ndim != 2 or ijkl.shape[1] != 4:
        raise CloudAtlasError(1, f"ijkl matrix should have 4 cols, but it has N={ijkl.shape[-1]}")
    with open(_filename(filebase), "w") as io:
        io.write(f"{comment_char} {ijkl.shape[0]}\n")
        for row in ijkl:
            io.write(" ".join(str(int(v)) for v in row) + "\n")


def save(A, filebase, cc="#"):
    """save(A::Matrix) / save(x::Vector) -- BasisFunctions.jl:938-958."""
    A = np.asarray(A)
    with open(_filename(filebase), "w") as io:
        if A.ndim == 2:
            io.write(f"{cc} {A.shape[0]} {A.shape[1]}\n")
            for row in A:
                io.write(" ".join(_fmt(v) for v in row) + "\n")
        elif A.ndim == 1:
            io.write(f"{cc} {A.shape[0]}\n")
            for v in A:
                io.write(_fmt(v) + "\n")
        else:
            raise CloudAtlasError(1, "save: expected a vector or a matrix")


def load(filename, cc="%#"):
    """Read a vector or matrix written by `save` / the channelflow tools (comment lines start with % or #)."""
    rows = []
    with open(filename) as fh:
        for line in fh:
            line = line.strip()
            if not line or line[0] in cc:
                continue
            rows.append([float(t) for t in line.split()])
    X = np.array(rows)
    return X[:, 0] if X.ndim == 2 and X.shape[1] == 1 else X


def basis_index_dict(ijklPsi, ijklPhi):
    """n -> m such that ijklPsi[n] == ijklPhi[m] (0-based)  --  BasisFunctions.jl:987-997."""
    where = {tuple(int(v) for v in row): m for m, row in reversed(list(enumerate(np.asarray(ijklPhi))))}
    out = {}
    for n, row in enumerate(np.asarray(ijklPsi)):
        m = where.get(tuple(int(v) for v in row))
        if m is not None:
            out[n] = m
    return out


def changebasis(x, ijklPsi, ijklPhi):
    """Coefficients x in basis Psi -> coefficients in basis Phi by matching (i,j,k,l); modes missing from Phi
    are dropped, modes missing from Psi are zero  --  BasisFunctions.jl:1000-1021."""
    y = np.zeros(np.asarray(ijklPhi).shape[0])
    for n, m in basis_index_dict(ijklPsi, ijklPhi).items():
        y[m] = x[n]
    return y
