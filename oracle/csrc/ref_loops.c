/* CPU restatement of the reference's evaluation loops -- TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * Same structure as the reference's single-threaded Julia (parity status: PINNED through
 * tests/test_oracle_csrc.py, which checks these loops against oracle/bilinear.py and oracle/model.py,
 * themselves pinned to the reference's known answers):
 *   ref_bilinear_eval   rtn[i] += val * x[j] * y[k] over COO, Int64 indices   src/SparseBilinear.jl:56-65
 *   ref_jacobian        DN[i,j] += val*x[k]; DN[i,k] += val*x[j]              src/SparseBilinear.jl:75-91
 *   ref_rhs             Bfact \ (A1*x + (1/R)*(A2*x) + N(x))                  src/ODEModels.jl:57
 * The LU solve uses a pre-factored B (ref_lu_factor = partial-pivot LU like LinearAlgebra.lu,
 * src/ODEModels.jl:54).  ref_rhs_ensemble runs ref_rhs for every state, optionally with OpenMP over
 * states (the reference itself is single-threaded; the threaded figure is the generous baseline).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

void ref_bilinear_eval(const int64_t* ijk, const double* val, int64_t nnz, int64_t m, const double* x,
                       const double* y, double* out) {
  memset(out, 0, (size_t)m * sizeof(double));
  for (int64_t r = 0; r < nnz; ++r)
    out[ijk[r] - 1] += val[r] * x[ijk[r + nnz] - 1] * y[ijk[r + 2 * nnz] - 1];
}

void ref_jacobian(const int64_t* ijk, const double* val, int64_t nnz, int64_t m, const double* x,
                  double* DN /* m x m column-major */) {
  memset(DN, 0, (size_t)m * m * sizeof(double));
  for (int64_t r = 0; r < nnz; ++r) {
    const int64_t i = ijk[r] - 1, j = ijk[r + nnz] - 1, k = ijk[r + 2 * nnz] - 1;
    DN[i + m * j] += val[r] * x[k];
    DN[i + m * k] += val[r] * x[j];
  }
}

/* in-place LU with partial pivoting of a column-major m x m matrix; returns 0 on success */
int ref_lu_factor(double* A, int64_t m, int64_t* piv) {
  for (int64_t c = 0; c < m; ++c) {
    int64_t p = c;
    for (int64_t r = c + 1; r < m; ++r)
      if (fabs(A[r + m * c]) > fabs(A[p + m * c])) p = r;
    piv[c] = p;
    if (A[p + m * c] == 0.0) return 1;
    if (p != c)
      for (int64_t k = 0; k < m; ++k) {
        double t = A[c + m * k];
        A[c + m * k] = A[p + m * k];
        A[p + m * k] = t;
      }
    for (int64_t r = c + 1; r < m; ++r) A[r + m * c] /= A[c + m * c];
    for (int64_t k = c + 1; k < m; ++k) {
      const double f = A[c + m * k];
      if (f != 0.0)
        for (int64_t r = c + 1; r < m; ++r) A[r + m * k] -= A[r + m * c] * f;
    }
  }
  return 0;
}

void ref_lu_solve(const double* LU, const int64_t* piv, int64_t m, double* b) {
  for (int64_t c = 0; c < m; ++c)
    if (piv[c] != c) {
      double t = b[c];
      b[c] = b[piv[c]];
      b[piv[c]] = t;
    }
  for (int64_t c = 0; c < m; ++c) {
    const double f = b[c];
    if (f != 0.0)
      for (int64_t r = c + 1; r < m; ++r) b[r] -= LU[r + m * c] * f;
  }
  for (int64_t c = m - 1; c >= 0; --c) {
    b[c] /= LU[c + m * c];
    const double f = b[c];
    if (f != 0.0)
      for (int64_t r = 0; r < c; ++r) b[r] -= LU[r + m * c] * f;
  }
}

/* f(x,R) for one state; work: m doubles of scratch */
void ref_rhs(const double* LU, const int64_t* piv, const double* A1, const double* A2, const int64_t* ijk,
             const double* val, int64_t nnz, int64_t m, const double* x, double R, double* out, double* work) {
  ref_bilinear_eval(ijk, val, nnz, m, x, x, out);
  memset(work, 0, (size_t)m * sizeof(double));
  for (int64_t j = 0; j < m; ++j) { /* A1*x and A2*x as column sweeps (dense gemv) */
    const double xj = x[j];
    for (int64_t i = 0; i < m; ++i) {
      out[i] += A1[i + m * j] * xj;
      work[i] += A2[i + m * j] * xj;
    }
  }
  const double invR = 1.0 / R;
  for (int64_t i = 0; i < m; ++i) out[i] += invR * work[i];
  ref_lu_solve(LU, piv, m, out);
}

/* X, OUT: nstates x m column-major (state index fastest); threads <= 1 means serial.
 * Plain pthreads (static split of the states) -- the image's gcc wrapper has no OpenMP spec file. */
typedef struct {
  const double *LU, *A1, *A2, *val, *X;
  const int64_t *piv, *ijk;
  int64_t nnz, m, nstates, s0, s1;
  double R;
  double* OUT;
} ens_job;

static void* ens_worker(void* arg) {
  ens_job* jb = (ens_job*)arg;
  const int64_t m = jb->m, ns = jb->nstates;
  double* x = (double*)malloc((size_t)m * 3 * sizeof(double));
  double *o = x + m, *w = x + 2 * m;
  for (int64_t s = jb->s0; s < jb->s1; ++s) {
    for (int64_t j = 0; j < m; ++j) x[j] = jb->X[s + ns * j];
    ref_rhs(jb->LU, jb->piv, jb->A1, jb->A2, jb->ijk, jb->val, jb->nnz, m, x, jb->R, o, w);
    for (int64_t j = 0; j < m; ++j) jb->OUT[s + ns * j] = o[j];
  }
  free(x);
  return 0;
}

void ref_rhs_ensemble(const double* LU, const int64_t* piv, const double* A1, const double* A2,
                      const int64_t* ijk, const double* val, int64_t nnz, int64_t m, const double* X,
                      int64_t nstates, double R, double* OUT, int threads) {
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  pthread_t tid[256];
  ens_job jobs[256];
  for (int t = 0; t < threads; ++t) {
    ens_job jb = {LU, A1, A2, val, X, piv, ijk, nnz, m, nstates, nstates * t / threads,
                  nstates * (t + 1) / threads, R, OUT};
    jobs[t] = jb;
    if (threads == 1) ens_worker(&jobs[t]);
    else pthread_create(&tid[t], 0, ens_worker, &jobs[t]);
  }
  if (threads > 1)
    for (int t = 0; t < threads; ++t) pthread_join(tid[t], 0);
}
