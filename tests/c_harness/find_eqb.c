/* find_eqb.c -- the call sequence of the Julia binding (cloudatlas.jl_b200/julia/CloudAtlasB200.jl), and through it
 * of the reference's notebooks/find-eqb.ipynb and test/runtests.jl:40-107, driven from plain C with HOST buffers:
 *
 *   ijkl   = basisIndices(J,K,L,H)                 ca_basis_indices            (BasisFunctions.jl:566-574)
 *   model  = ODEModel(alpha,gamma,J,K,L,H)         ca_assemble + ca_model_create_from_assembly   (ODEModels.jl:22-61)
 *   model.f(xeqb, R)                               ca_model_rhs                (ODEModels.jl:57)
 *   model.Df(x, R)                                 ca_model_jac                (ODEModels.jl:58)
 *   model.f(X, R) for an ensemble                  ca_pin + ca_model_rhs_batched
 *   hookstepsolve(model, R, xguess)                ca_hookstep_solve_model     (Hookstep.jl:109-352)
 *
 * No CUDA, no Python: this is what a non-Python host sees of the library.  Exit code 0 and the line "harness ok" mean
 * every check passed.  Build:  gcc -O2 -I include tests/c_harness/find_eqb.c -L cloudatlas.jl_b200/lib -lcloudatlas_b200 -lm */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cloudatlas_b200.h"

#define CHECK(call)                                                                  \
  do {                                                                               \
    int32_t st__ = (call);                                                           \
    if (st__ != CA_OK) {                                                             \
      fprintf(stderr, "%s failed (%d): %s\n", #call, (int)st__, ca_last_error());   \
      return 1;                                                                      \
    }                                                                                \
  } while (0)

static double norm2(const double* v, int n) {
  double s = 0;
  for (int i = 0; i < n; ++i) s += v[i] * v[i];
  return sqrt(s);
}

/* test/runtests.jl:49-53 and :70-71 */
static const double xeqb[17] = {0.16764501289529937, -0.018552429231415948, 0.48640800040772025, -0.09759477649001111,
                                -0.03411986464200022, -0.010665050862032609, -0.01731713276665348, 0.020113317487885345,
                                -0.03572908824700301, -0.008809478655993888, -0.018431734698126298, -0.04736482532692928,
                                -0.04609045287990289, -0.03157916691896989, 0.0010374375712356061, 0.01601957685174594,
                                0.07702704667134751};
static const double xguess[17] = {0.105, -0.0539, 0.388, -0.0172, -0.0133, 0.0113, -0.00706, 0.0240, -0.0344, -0.00868,
                                  -0.01264, -0.0234, -0.0320, -0.0180, -0.00464, 0.0106, 0.0235};

int main(void) {
  const double alpha = 1.0, gamma = 2.0, R = 200.0;
  /* H = [sx*sy*sz, sz*tx*tz]  (runtests.jl:44; Symmetries.jl:68-75) */
  const ca_symmetry H[2] = {{-1, -1, -1, 0, 0, 1}, {1, 1, -1, 1, 1, 1}};
  int64_t m = 0;
  CHECK(ca_basis_indices(1, 1, 3, H, 2, NULL, &m));
  if (m != 17) { fprintf(stderr, "expected 17 modes, got %lld\n", (long long)m); return 1; }
  int64_t* ijkl = (int64_t*)malloc(sizeof(int64_t) * 4 * (size_t)m);
  CHECK(ca_basis_indices(1, 1, 3, H, 2, ijkl, &m));

  ca_assembly* a = NULL;
  CHECK(ca_assemble(alpha, gamma, ijkl, m, 0, 0, m, &a));
  int64_t nnz = 0;
  CHECK(ca_assembly_info(a, NULL, NULL, NULL, &nnz, NULL));
  if (nnz != 396) { fprintf(stderr, "expected 396 entries of N, got %lld\n", (long long)nnz); return 1; }
  ca_model* model = NULL;
  CHECK(ca_model_create_from_assembly(a, &model));
  double* B = (double*)calloc((size_t)(m * m), sizeof(double));
  double* A1 = (double*)calloc((size_t)(m * m), sizeof(double));
  double* A2 = (double*)calloc((size_t)(m * m), sizeof(double));
  CHECK(ca_assembly_get_linear(a, B, A1, A2));
  CHECK(ca_assembly_destroy(a));

  /* @test norm(model.f(xeqb, R))/norm(xeqb) < 1e-10     (runtests.jl:57) */
  double f[17];
  CHECK(ca_model_rhs(model, xeqb, R, f));
  const double res = norm2(f, 17) / norm2(xeqb, 17);
  printf("|f(xeqb)|/|xeqb| = %.3e\n", res);
  if (!(res < 1e-10)) return 1;

  /* Df by finite differences of f (Hookstep.jl:77-93) */
  double* Df = (double*)malloc(sizeof(double) * 17 * 17);
  CHECK(ca_model_jac(model, xeqb, R, Df));
  double worst = 0;
  for (int j = 0; j < 17; ++j) {
    double xp[17], fp[17];
    memcpy(xp, xeqb, sizeof xp);
    xp[j] += 1e-6;
    CHECK(ca_model_rhs(model, xp, R, fp));
    for (int i = 0; i < 17; ++i) {
      const double fd = (fp[i] - f[i]) / 1e-6, d = fabs(fd - Df[i + 17 * j]);
      if (d > worst) worst = d;
    }
  }
  printf("max |Df - finite difference| = %.3e\n", worst);
  if (!(worst < 1e-4)) return 1;

  /* ensemble through the host-pointer call, page-locked buffers: row s of X is xeqb scaled by (1 + s * 1e-3) */
  const int64_t ns = 4099;
  double* X = (double*)malloc(sizeof(double) * (size_t)(ns * m));
  double* F = (double*)malloc(sizeof(double) * (size_t)(ns * m));
  for (int64_t s = 0; s < ns; ++s)
    for (int j = 0; j < 17; ++j) X[s + ns * j] = xeqb[j] * (1.0 + 1e-3 * (double)s);
  CHECK(ca_pin(X, (int64_t)sizeof(double) * ns * m));
  CHECK(ca_pin(F, (int64_t)sizeof(double) * ns * m));
  CHECK(ca_model_rhs_batched(model, X, ns, R, F));
  CHECK(ca_unpin(X));
  CHECK(ca_unpin(F));
  double diff = 0;
  for (int j = 0; j < 17; ++j) diff = fmax(diff, fabs(F[0 + ns * j] - f[j]));
  double x7[17], f7[17];
  for (int j = 0; j < 17; ++j) x7[j] = X[7 + ns * j];
  CHECK(ca_model_rhs(model, x7, R, f7));
  for (int j = 0; j < 17; ++j) diff = fmax(diff, fabs(F[7 + ns * j] - f7[j]));
  printf("ensemble vs single-state calls: %.3e\n", diff);
  if (!(diff < 1e-13)) return 1;

  /* xsoln, success = hookstepsolve(model, R, xguess);  @test norm(xsoln - xeqb)/norm(xeqb) < 1e-08   (runtests.jl:78-81) */
  ca_search_params p = {1e-8, 1e-8, 0.02, 20, 4, 6, 0};
  double xs[17];
  int32_t success = 0;
  ca_solve_log log;
  CHECK(ca_hookstep_solve_model(model, &R, xguess, 1, &p, xs, &success, &log));
  double d[17];
  for (int j = 0; j < 17; ++j) d[j] = xs[j] - xeqb[j];
  const double err = norm2(d, 17) / norm2(xeqb, 17);
  printf("hookstep: %d Newton steps, residual %.3e, |x - xeqb|/|xeqb| = %.3e, exit %d\n", (int)log.newton_steps,
         log.final_residual, err, (int)log.exit_reason);
  if (!(err < 1e-8)) return 1;

  /* observables installed by ca_model_create_from_assembly */
  double obs[2];
  CHECK(ca_model_observables_batched(model, xeqb, 1, obs));
  printf("shear %.12f dissipation %.12f\n", obs[0], obs[1]);
  if (!(obs[0] > 1.0 && obs[1] > 1.0)) return 1;

  CHECK(ca_model_destroy(model));
  free(ijkl); free(B); free(A1); free(A2); free(Df); free(X); free(F);
  printf("harness ok\n");
  return 0;
}
