"""CPU oracle for the CloudAtlas hot path  --  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithms of johnfgibson/CloudAtlas.jl that
build and evaluate  dx/dt = B^-1 (A1 x + A2 x / R + N(x,x)).  It exists to *check*
the CUDA path in ``cloudatlas_b200``; nothing in the product imports it.  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import or execute anything under ``oracle/``.

Parity status: PINNED.  Julia is not installed in the build container, so the
reference itself cannot be executed; instead every known-answer value the reference
ships for this path is reproduced by this restatement (``tests/test_oracle_golden.py``):
``test/runtests.jl:29-107`` (2-D hookstep, 17d equilibrium residual, hookstep to the
17-digit equilibria with and without normalisation) and the stored notebook outputs
``notebooks/find-eqb.ipynb`` (m, cond(B), normalised Psi coefficients, xsoln,
eigenvalues) and ``notebooks/continue-eqb.ipynb`` (27d residual norm from the shipped
DNS projection).  Entry-level values of B/A1/A2/N beyond those joint checks are not
pinned by the reference; they are pinned by the exact-rational tier of this oracle.

Modules
  algebra.py    type-generic (float | fractions.Fraction) restatement of the symbolic
                FourierMode / BasisComponent / BasisFunction algebra
  basis.py      index sets, symmetry filter, basisSet
  model.py      Galerkin assembly of B, A1, A2, N and the closures f, Df
  bilinear.py   SparseBilinear COO container, N(x,y), derivative(N,x)
  hookstep.py   Newton-hookstep state machine
  exact.py      fast exact-rational assembly by separable tabulation (any L)
  fastfloat.py  the same tables rounded once to Float64, vectorised (N at m = 307 / 999 in seconds)
  cref.py       ctypes front end of csrc/
  csrc/         plain-C restatement of the evaluation loops (COO N(x), derivative, f with LU solve; pthreads over
                states) -- timed as the CPU baseline by bench.py
  model.py also restates cnab2, shear_function and dissipation_function (ODEModels.jl:65-181).
"""
