// Device-resident Newton-hookstep for an ODEModel: one persistent CTA per search.
//
// Replaces, for f = model.f(.,R) and Df = model.Df(.,R), the host loop of src/Hookstep.jl:109-352
// (hookstepsolve) and :21-74 (hookstep).  On a GPU the 17/27/44-dimensional searches of the reference's
// tests and notebooks are launch-latency-bound if every f, Df, LU and solve is its own kernel, so the
// whole search runs inside ONE kernel: x, f(x), Df(x), H = Df'Df and the LU workspace live in shared
// memory, f and Df are evaluated in-kernel from the folded operators (L1 + L2/R, Q = B^-1 N; CSR rows
// for f, the sliced-ELL Jacobian structure for Df), and every reduction has a fixed order, so a search
// is bit-reproducible.  Decision tree, constants (mu0 = 1e-6, dtol = 1e-4, factors 3/2, 1/2, 2/3, 4/3,
// 0.01 / 0.10 / 0.8 / 1.2) and exit flags are the reference's.
#include <algorithm>
#include <vector>

#include "model.h"

using namespace ca;

namespace {

constexpr int kHookThreads = 256;

struct SolverArgs {
  int m;
  // folded operators
  const double* L1;
  const double* L2;
  const int32_t* q_rowptr;
  const uint32_t* q_ab;
  const double* q_val;
  // Jacobian structure of Q
  const int64_t* jac_slice_off;
  const int32_t* jac_out_pos;
  const uint16_t* jac_src;
  const double* jac_val;
  int64_t jac_nslices;
  // problem
  const double* R;
  const double* Xguess;  // nsolve x m column-major
  int64_t nsolve;
  ca_search_params p;
  double* Xout;
  int32_t* success;
  ca_solve_log* logs;
};

// ---- block-wide primitives (all threads call; results identical in every thread) -----------------
__device__ double block_dot(const double* a, const double* b, int m, double* scratch) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double s = 0.0;
  for (int i = tid; i < m; i += kHookThreads) s = fma(a[i], b[i], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __syncthreads();  // scratch may still be read from a previous call
  if (lane == 0) scratch[warp] = s;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < kHookThreads / 32; ++w) t += scratch[w];
  return t;
}

// y = A x, A m x m column-major in shared memory; one warp per row, fixed summation order
__device__ void block_gemv(const double* A, const double* x, double* y, int m, bool transpose) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = warp; i < m; i += kHookThreads / 32) {
    double s = 0.0;
    for (int j = lane; j < m; j += 32) s = fma(transpose ? A[j + m * i] : A[i + m * j], x[j], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) y[i] = s;
  }
  __syncthreads();
}

// Small systems (m <= kWarpLuMax): one warp factors and solves with __syncwarp only -- a block barrier costs
// more than a whole column step at these sizes (4 barriers per column made a 17 x 17 LU ~20x slower).
constexpr int kWarpLuMax = 64;

// argmax of the lanes' (best, at) with the smallest row winning ties, in three redux ops: the bit pattern of a
// non-negative double is monotonic in its value.  best < 0 marks a lane without rows.
__device__ __forceinline__ void warp_argmax(double best, int at, double& gbest, int& gat) {
  const unsigned long long key = best < 0.0 ? 0ull : (unsigned long long)__double_as_longlong(best) + 1ull;
  const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
  const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
  const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
  gat = (int)__reduce_min_sync(0xffffffffu, (hi == mhi && lo == mlo) ? (unsigned)at : 0xffffffffu);
  const unsigned long long gk = ((unsigned long long)mhi << 32) | mlo;
  gbest = gk == 0ull ? -1.0 : __longlong_as_double((long long)(gk - 1ull));
}

// perm: the row permutation of the factorisation as one gather (b <- b[perm]), kept next to LAPACK-style piv
__device__ void warp_lu(double* W, int* piv, int* perm, int m, int* flag) {
  const int lane = threadIdx.x & 31;
  for (int i = lane; i < m; i += 32) perm[i] = i;
  __syncwarp();
  for (int k = 0; k < m; ++k) {
    double best = -1.0;
    int at = k;
    for (int r = k + lane; r < m; r += 32) {
      const double v = fabs(W[r + m * k]);
      if (v > best) { best = v; at = r; }
    }
    warp_argmax(best, at, best, at);
    if (best == 0.0) {
      if (lane == 0) *flag = 1;
      return;
    }
    if (lane == 0) {
      piv[k] = at;
      const int t = perm[k];
      perm[k] = perm[at];
      perm[at] = t;
    }
    if (at != k)
      for (int j = lane; j < m; j += 32) {
        const double t = W[k + m * j];
        W[k + m * j] = W[at + m * j];
        W[at + m * j] = t;
      }
    __syncwarp();
    // multipliers of this lane's (at most two, m <= 64) rows stay in registers; one pass over the columns
    const double pivot = W[k + m * k];
    const int r0 = k + 1 + lane, r1 = r0 + 32;
    double l0 = 0.0, l1 = 0.0;
    if (r0 < m) { l0 = W[r0 + m * k] / pivot; W[r0 + m * k] = l0; }
    if (r1 < m) { l1 = W[r1 + m * k] / pivot; W[r1 + m * k] = l1; }
    // four columns per pass, all loads before the stores: row k is never written here, but the compiler cannot know
    // that and would otherwise serialise every column behind the previous column's store
    const bool h0 = r0 < m, h1 = r1 < m;
    int c = k + 1;
    for (; c + 4 <= m; c += 4) {
      double u[4], w0[4], w1[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        u[q] = W[k + m * (c + q)];  // broadcast read
        w0[q] = h0 ? W[r0 + m * (c + q)] : 0.0;
        w1[q] = h1 ? W[r1 + m * (c + q)] : 0.0;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (h0) W[r0 + m * (c + q)] = fma(-l0, u[q], w0[q]);
        if (h1) W[r1 + m * (c + q)] = fma(-l1, u[q], w1[q]);
      }
    }
    for (; c < m; ++c) {
      const double u = W[k + m * c];
      if (h0) W[r0 + m * c] = fma(-l0, u, W[r0 + m * c]);
      if (h1) W[r1 + m * c] = fma(-l1, u, W[r1 + m * c]);
    }
    __syncwarp();
  }
}

// b <- (LU)^-1 b with b in registers (lane owns rows lane and lane + 32): the same operations in the same order as
// the shared-memory loops of block_lu_solve, minus their per-column round trips and barriers.
__device__ void warp_lu_solve(const double* W, const int* perm, double* b, int m) {
  const int lane = threadIdx.x & 31;
  const int r0 = lane, r1 = lane + 32;
  double b0 = r0 < m ? b[perm[r0]] : 0.0, b1 = r1 < m ? b[perm[r1]] : 0.0;
  __syncwarp();
  for (int c = 0; c < m; ++c) {
    const double f = __shfl_sync(0xffffffffu, c < 32 ? b0 : b1, c & 31);
    if (r0 > c && r0 < m) b0 = fma(-W[r0 + m * c], f, b0);
    if (r1 > c && r1 < m) b1 = fma(-W[r1 + m * c], f, b1);
  }
  for (int c = m - 1; c >= 0; --c) {
    double mine = c < 32 ? b0 : b1;
    if (lane == (c & 31)) mine /= W[c + m * c];
    const double f = __shfl_sync(0xffffffffu, mine, c & 31);
    if (r0 == c) b0 = f;
    if (r1 == c) b1 = f;
    if (r0 < c) b0 = fma(-W[r0 + m * c], f, b0);
    if (r1 < c && r1 < m) b1 = fma(-W[r1 + m * c], f, b1);
  }
  if (r0 < m) b[r0] = b0;
  if (r1 < m) b[r1] = b1;
  __syncwarp();
}

// In-place LU with partial pivoting (first maximal |entry| wins, like LAPACK's idamax).  Returns false
// on an exactly zero pivot.  m <= kWarpLuMax: one warp (above).  Larger systems use the whole CTA, three barriers per
// column:
//   warp 0 finds the pivot and publishes its index, its signed value and the old diagonal entry            | barrier
//   all threads swap rows k and p in the columns j != k; column k is written from the published scalars
//   (diagonal <- pivot, multipliers l_r = W[r,k] / pivot with the swapped entry taken from the scalar)       | barrier
//   rank-one update of the trailing block: warp w takes the columns k+1+w, k+1+w+8, ..., a lane its rows     | barrier
// Same operations on the same values as warp_lu (l = w / pivot, w <- fma(-l, u, w)): the factors are bit-identical.
// (Round 1 ran the 44 x 44 factorisations of the reference's configs[1] on ONE warp: 70 % of that search.)
#ifndef CA_WARP_LU_MAX
#define CA_WARP_LU_MAX 24
#endif
constexpr int kWarpLuOnly = CA_WARP_LU_MAX;  // up to here the one-warp factorisation is the faster one
__device__ bool block_lu(double* W, int* piv, int* perm, int m, int* flag) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (m <= kWarpLuOnly) {
    if (tid < 32) warp_lu(W, piv, perm, m, flag);
    __syncthreads();
    return *flag == 0;
  }
  __shared__ double pivot_and_diag[2];
  for (int i = tid; i < m; i += kHookThreads) perm[i] = i;
  __syncthreads();
  for (int k = 0; k < m; ++k) {
    if (tid < 32) {
      double best = -1.0;
      int at = k;
      for (int r = k + lane; r < m; r += 32) {
        const double v = fabs(W[r + m * k]);
        if (v > best) { best = v; at = r; }
      }
      warp_argmax(best, at, best, at);
      if (lane == 0) {
        piv[k] = at;
        if (best == 0.0) {
          *flag = 1;
        } else {
          const int t = perm[k];
          perm[k] = perm[at];
          perm[at] = t;
          pivot_and_diag[0] = W[at + m * k];
          pivot_and_diag[1] = W[k + m * k];
        }
      }
    }
    __syncthreads();
    if (*flag) return false;
    const int p = piv[k];
    const double pivot = pivot_and_diag[0], diag = pivot_and_diag[1];
    if (p != k)
      for (int j = tid; j < m; j += kHookThreads)
        if (j != k) {
          const double t = W[k + m * j];
          W[k + m * j] = W[p + m * j];
          W[p + m * j] = t;
        }
    for (int r = k + tid; r < m; r += kHookThreads)
      W[r + m * k] = r == k ? pivot : (r == p ? diag : W[r + m * k]) / pivot;
    __syncthreads();
    for (int c = k + 1 + warp; c < m; c += kHookThreads / 32) {
      const double u = W[k + m * c];  // broadcast
      for (int r = k + 1 + lane; r < m; r += 32) W[r + m * c] = fma(-W[r + m * k], u, W[r + m * c]);
    }
    __syncthreads();
  }
  return true;
}

// b <- (LU)^-1 b
__device__ void block_lu_solve(const double* W, const int* piv, const int* perm, double* b, int m) {
  const int tid = threadIdx.x;
  if (m <= kWarpLuMax) {  // two rows per lane in registers, shuffles instead of barriers: faster than the CTA loops below
    if (tid < 32) warp_lu_solve(W, perm, b, m);
    __syncthreads();
    return;
  }
  if (tid == 0)
    for (int k = 0; k < m; ++k) {
      const int p = piv[k];
      if (p != k) { const double t = b[k]; b[k] = b[p]; b[p] = t; }
    }
  __syncthreads();
  for (int c = 0; c < m; ++c) {
    const double f = b[c];
    for (int r = c + 1 + tid; r < m; r += kHookThreads) b[r] = fma(-W[r + m * c], f, b[r]);
    __syncthreads();
  }
  for (int c = m - 1; c >= 0; --c) {
    if (tid == 0) b[c] /= W[c + m * c];
    __syncthreads();
    const double f = b[c];
    for (int r = tid; r < c; r += kHookThreads) b[r] = fma(-W[r + m * c], f, b[r]);
    __syncthreads();
  }
}

// ---- the hookstep's mu search in a basis where H = Df'Df is tridiagonal -------------------------------------------
// Hookstep.jl:21-74 solves (H + mu I) dx = -Df'f  and  (H + mu I) t = dx  for ~20 values of mu per hookstep (Newton
// iteration on mu for |dx(mu)| = delta), each a dense factorisation in the reference: ~400 factorisations of a 44 x 44
// matrix in the 44-mode search, 70 % of its time in round 1 and still 53 % with a CTA-wide Cholesky.  Here H is
// reduced ONCE per Newton step to H = Q T Q' (Householder, T tridiagonal); in the coordinates z = Q'dx every solve is
// a pivoted LU of the tridiagonal T + mu I (O(m)), |dx| = |z| and dx.t = z.(Q't), so a mu iteration is O(m) scalar
// work without a barrier.  Q is never formed: the reflectors stay in H's columns and one warp applies them to the
// three vectors a hookstep needs (Q'g, Q'dx_newton, Q z).
struct Tridiag {            // arrays of m entries, carved from the (then idle) factorisation workspace W
  double *a, *b;            // diagonal and sub-diagonal of T
  double *v0, *beta;        // first component and 2 / |v|^2 of reflector k (the rest of v_k is H[k+2.., k]); beta = 0: none
  double *qg, *z, *t;       // Q'g; coordinates of the current dx; a solve's result
  double *dl, *rd, *du, *du2;  // factors of T + mu I: multipliers, 1 / diagonal of U, its two super-diagonals
  double *v, *p;            // p: Householder work vector (v: spare)
  int* piv;                 // 1: rows i and i+1 were interchanged
  static constexpr int kDoubles = 14;  // 13 arrays of doubles + one of ints, in units of m doubles
  __device__ void carve(double* W, int m) {
    a = W; b = a + m; v0 = b + m; beta = v0 + m; qg = beta + m; z = qg + m; t = z + m;
    dl = t + m; rd = dl + m; du = rd + m; du2 = du + m; v = du2 + m; p = v + m;
    piv = reinterpret_cast<int*>(p + m);
  }
};

// H (m x m, symmetric, column-major) -> T.a, T.b and the reflectors (v0, beta; tails in H's columns).
// Two barriers per column: the norms and dot products are computed by every warp for itself (same data, same order of
// operations, hence the same value in every thread) instead of through block-wide reductions.
__device__ void block_tridiagonalise(double* H, Tridiag& T, int m) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  auto warp_sum = [](double s) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
  };
  for (int k = 0; k + 2 < m; ++k) {
    const int n = m - k - 1;                    // trailing block H22 = H[k+1.., k+1..]
    const double* x = H + (k + 1) + m * k;      // the column below the diagonal: v = (v0, x[1..])
    double part = 0.0;
    for (int i = 1 + lane; i < n; i += 32) part = fma(x[i], x[i], part);
    const double sigma = warp_sum(part);
    const double x0 = x[0];
    if (sigma == 0.0) {                          // already tridiagonal in this column (same value in every thread)
      if (tid == 0) { T.a[k] = H[k + m * k]; T.b[k] = x0; T.v0[k] = 0.0; T.beta[k] = 0.0; }
      continue;
    }
    const double alpha = -copysign(sqrt(x0 * x0 + sigma), x0);
    const double v0 = x0 - alpha;
    const double beta = 2.0 / (v0 * v0 + sigma);  // reflector I - beta v v'
    if (tid == 0) { T.a[k] = H[k + m * k]; T.b[k] = alpha; T.v0[k] = v0; T.beta[k] = beta; }
    // p = beta H22 v: four lanes per row (columns j = g, g + 4, ...), combined by shuffles; H22 is symmetric, so the
    // lanes of a row group walk down columns and consecutive rows are consecutive addresses
    for (int i0 = 0; i0 < n; i0 += kHookThreads / 4) {
      const int i = i0 + (tid >> 2), g = tid & 3;
      double sum = 0.0, sum2 = 0.0;
      if (i < n) {
        const double* hrow = H + (k + 1 + i) + m * (k + 1);
        int j = g;
        for (; j + 4 < n; j += 8) {  // two independent chains
          sum = fma(hrow[m * j], j == 0 ? v0 : x[j], sum);
          sum2 = fma(hrow[m * (j + 4)], x[j + 4], sum2);
        }
        if (j < n) sum = fma(hrow[m * j], j == 0 ? v0 : x[j], sum);
        sum += sum2;
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      if (i < n && g == 0) T.p[i] = beta * sum;
    }
    __syncthreads();
    part = 0.0;
    for (int i = lane; i < n; i += 32) part = fma(i == 0 ? v0 : x[i], T.p[i], part);
    const double K = 0.5 * beta * warp_sum(part);
    // H22 -= v w' + w v' with w = p - K v: a warp per column, lanes over the rows
    double vi[3], wi[3];  // this lane's rows (m <= 95: at most three)
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int i = lane + 32 * r;
      vi[r] = i < n ? (i == 0 ? v0 : x[i]) : 0.0;
      wi[r] = i < n ? T.p[i] - K * vi[r] : 0.0;
    }
    for (int j = warp; j < n; j += kHookThreads / 32) {
      const double vj = j == 0 ? v0 : x[j], wj = T.p[j] - K * vj;
      double* hcol = H + (k + 1) + m * (k + 1 + j);
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int i = lane + 32 * r;
        if (i < n) hcol[i] -= vi[r] * wj + wi[r] * vj;
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    if (m >= 2) { T.a[m - 2] = H[(m - 2) + m * (m - 2)]; T.b[m - 2] = H[(m - 1) + m * (m - 2)]; }
    T.a[m - 1] = H[(m - 1) + m * (m - 1)];
  }
  __syncthreads();
}

// One warp: y <- Q' y (transpose) or y <- Q y, Q = P_0 P_1 ... P_{m-3}, P_k = I - beta_k v_k v_k' on components k+1..
__device__ void warp_apply_reflectors(const double* H, const Tridiag& T, double* y, int m, bool transpose) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  for (int q = 0; q + 2 < m; ++q) {
    const int k = transpose ? q : m - 3 - q;
    const double beta = T.beta[k];
    if (beta == 0.0) continue;
    const int n = m - k - 1;
    const double* tail = H + (k + 1) + m * k;  // tail[i] = v_i for i >= 1
    double dot = 0.0;
    for (int i = lane; i < n; i += 32) dot = fma(i == 0 ? T.v0[k] : tail[i], y[k + 1 + i], dot);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    const double c = beta * dot;
    for (int i = lane; i < n; i += 32) y[k + 1 + i] -= c * (i == 0 ? T.v0[k] : tail[i]);
    __syncwarp();
  }
}

// One thread: LU with partial pivoting of the tridiagonal T + mu I (the mu iteration may step to negative mu, where
// the matrix is indefinite; the reference's `\` pivots too).  Gaussian elimination with row interchanges between
// neighbours, fill-in on a second super-diagonal (LAPACK dgttrf's scheme): dl = multipliers, rd = 1 / diagonal of U,
// du, du2 = super-diagonals of U divided by their row's pivot.  False on an exactly zero pivot.
__device__ bool tridiag_factor(const Tridiag& T, double mu, int m) {
  double d = T.a[0] + mu;           // current diagonal entry
  double u = m > 1 ? T.b[0] : 0.0;  // current super-diagonal entry (T is symmetric: the same b)
#pragma unroll 4  // (the loads of T.a, T.b do not depend on the chain d -> 1/d -> d: let them run ahead)
  for (int i = 0; i + 1 < m; ++i) {
    const double sub = T.b[i], dnext = T.a[i + 1] + mu, unext = i + 2 < m ? T.b[i + 1] : 0.0;
    if (fabs(d) >= fabs(sub)) {
      if (d == 0.0) return false;
      const double r = 1.0 / d, f = sub * r;
      T.dl[i] = f;
      T.rd[i] = r;
      T.du[i] = u * r;  // the super-diagonals are stored divided by the pivot: one fma per row on the solve's chain
      T.du2[i] = 0.0;
      T.piv[i] = 0;
      d = dnext - f * u;
      u = unext;
    } else {  // interchange rows i and i+1
      const double r = 1.0 / sub, f = d * r;
      T.dl[i] = f;
      T.rd[i] = r;
      T.du[i] = dnext * r;
      T.du2[i] = unext * r;
      T.piv[i] = 1;
      d = u - f * dnext;
      u = -f * unext;
    }
  }
  if (d == 0.0) return false;
  T.rd[m - 1] = 1.0 / d;
  return true;
}
// out = (T + mu I)^-1 rhs from the factors (out != rhs)
__device__ void tridiag_solve(const Tridiag& T, const double* rhs, double* out, int m) {
  double cur = rhs[0];
#pragma unroll 4
  for (int i = 0; i + 1 < m; ++i) {
    const double nxt = rhs[i + 1];
    if (T.piv[i]) {
      out[i] = nxt;
      cur = cur - T.dl[i] * nxt;
    } else {
      out[i] = cur;
      cur = nxt - T.dl[i] * cur;
    }
  }
  double x1 = cur * T.rd[m - 1], x2 = 0.0;
  out[m - 1] = x1;
#pragma unroll 4
  for (int i = m - 2; i >= 0; --i) {
    const double x0 = fma(-T.du[i], x1, fma(-T.du2[i], x2, out[i] * T.rd[i]));
    out[i] = x0;
    x2 = x1;
    x1 = x0;
  }
}

// out = (L1 + L2/R) x + Q(x,x)   -- ODEModels.jl:57 with B folded in
__device__ void eval_f(const SolverArgs& a, double invR, const double* x, double* out) {
  const int m = a.m, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = warp; i < m; i += kHookThreads / 32) {
    double s = 0.0;
    for (int j = lane; j < m; j += 32) s = fma(fma(a.L2[i + m * j], invR, a.L1[i + m * j]), x[j], s);
    for (int e = a.q_rowptr[i] + lane; e < a.q_rowptr[i + 1]; e += 32) {
      const uint32_t ab = a.q_ab[e];
      s = fma(a.q_val[e], x[ab & 0xffff] * x[ab >> 16], s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[i] = s;
  }
  __syncthreads();
}

// Df = L1 + L2/R + DQ(x)   -- ODEModels.jl:58 with B folded in; Df m x m column-major in shared memory
__device__ void eval_Df(const SolverArgs& a, double invR, const double* x, double* Df) {
  const int m = a.m, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < m * m; e += kHookThreads) Df[e] = fma(a.L2[e], invR, a.L1[e]);
  __syncthreads();
  for (int64_t sl = warp; sl < a.jac_nslices; sl += kHookThreads / 32) {
    const int64_t b = a.jac_slice_off[sl], e = a.jac_slice_off[sl + 1];
    double acc = 0.0;
    for (int64_t p = b + lane; p < e; p += 32) acc = fma(a.jac_val[p], x[a.jac_src[p]], acc);
    const int32_t o = a.jac_out_pos[sl * 32 + lane];
    if (o >= 0) Df[o] += acc;  // every output entry belongs to exactly one lane of one slice
  }
  __syncthreads();
}

#ifdef CA_HOOK_PROFILE
#define TIC long long t__ = clock64()
#define TOC(slot) prof[slot] += clock64() - t__
#else
#define TIC
#define TOC(slot)
#endif
__global__ void __launch_bounds__(kHookThreads, 1) hookstep_kernel(SolverArgs a) {
#ifdef CA_HOOK_PROFILE
  long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long tstart = clock64();
#endif
  extern __shared__ __align__(16) double sm[];
  const int m = a.m, tid = threadIdx.x;
  const int64_t s = blockIdx.x;
  double* Df = sm;
  double* H = Df + m * m;
  double* W = H + m * m;
  double* x = W + (m * m > Tridiag::kDoubles * m ? m * m : Tridiag::kDoubles * m);
  double* fx = x + m;
  double* dx = fx + m;
  double* dxn = dx + m;   // Newton step
  double* xh = dxn + m;   // x + dx
  double* fh = xh + m;    // f(x + dx)
  double* g = fh + m;     // Df' fx
  double* t1 = g + m;
  double* t2 = t1 + m;
  double* scratch = t2 + m;          // 8 doubles
  int* piv = (int*)(scratch + 8);    // m ints
  int* perm = piv + m;               // m ints
  int* flag = perm + m;

  Tridiag T;       // the mu search's workspace lives in W (idle between two Newton factorisations)
  T.carve(W, m);
  __shared__ int mu_status;  // 0 ok, 1: a tridiagonal factorisation met a zero pivot
  const double invR = 1.0 / a.R[s];
  const double ftol = a.p.ftol, xtol = a.p.xtol;
  double delta = a.p.delta;
  ca_solve_log* L = (a.logs && tid == 0) ? &a.logs[s] : nullptr;  // thread 0 writes the per-step log
  if (L)
    for (int n = 0; n < CA_LOG_STEPS; ++n) { L->residual[n] = 0.0; L->delta[n] = 0.0; L->hooksteps[n] = 0; }
  int exit_reason = CA_EXIT_MAX_NEWTON, nsteps = 0, nf = 0, ndf = 0;
  double last_res = 0.0;

  for (int i = tid; i < m; i += kHookThreads) x[i] = a.Xguess[s + a.nsolve * i];
  if (tid == 0) *flag = 0;
  __syncthreads();

  for (int n = 0; n < a.p.Nnewton; ++n) {
    { TIC; eval_f(a, invR, x, fx); TOC(0); }
    ++nf;
    const double rx = 0.5 * block_dot(fx, fx, m, scratch);
    last_res = sqrt(2 * rx);
    if (L && n < CA_LOG_STEPS) { L->residual[n] = last_res; L->delta[n] = delta; }
    nsteps = n;
    if (last_res <= ftol) { exit_reason = CA_EXIT_CONVERGED; break; }
    nsteps = n + 1;

    { TIC; eval_Df(a, invR, x, Df); TOC(1); }
    ++ndf;
    // Newton step: dx = -Df \ fx
    for (int e = tid; e < m * m; e += kHookThreads) W[e] = Df[e];
    for (int i = tid; i < m; i += kHookThreads) dx[i] = -fx[i];
    __syncthreads();
    { TIC; const bool okk = block_lu(W, piv, perm, m, flag); TOC(2); if (!okk) { exit_reason = CA_EXIT_SINGULAR; break; } }
    { TIC; block_lu_solve(W, piv, perm, dx, m); TOC(3); }
    const double norm_dx_newt = sqrt(block_dot(dx, dx, m, scratch));
    block_gemv(Df, dx, t1, m, false);  // Df dx
    if (block_dot(fx, t1, m, scratch) >= 0.0) { exit_reason = CA_EXIT_RESIDUAL_UP; break; }
    if (norm_dx_newt < xtol) {
      for (int i = tid; i < m; i += kHookThreads) x[i] += dx[i];
      __syncthreads();
      exit_reason = CA_EXIT_XTOL_NEWTON;
      break;
    }
    for (int i = tid; i < m; i += kHookThreads) { dxn[i] = dx[i]; xh[i] = x[i] + dx[i]; }
    __syncthreads();

    bool have_H = false, have_qg = false;
    bool xtol_exit = false;
    int nh = 0;
    for (int h = 0; h < a.p.Nhook; ++h) {
      ++nh;
      if (norm_dx_newt > delta) {
        // ---- hookstep(fx, Dfx, delta, dx_newt): Newton iteration on mu for norm(dx(mu)) = delta
        if (!have_H) {
          TIC;
          for (int e = tid; e < m * m; e += kHookThreads) {
            const int r = e % m, c = e / m;
            double sum = 0.0;
            for (int k = 0; k < m; ++k) sum = fma(Df[k + m * r], Df[k + m * c], sum);
            H[e] = sum;
          }
          __syncthreads();
          block_gemv(Df, fx, g, m, true);  // Df' fx
          TOC(4);
          { TIC; block_tridiagonalise(H, T, m); TOC(2); }  // H = Q T Q'; H now holds the reflectors
          have_H = true;
        }
        // ---- Newton iteration on mu, in the coordinates z = Q' dx: warp 0 alone, O(m) scalar work per iteration
        {
          TIC;
          if (tid < 32) {
            if (!have_qg) {
              for (int i = tid; i < m; i += 32) T.qg[i] = g[i];
              warp_apply_reflectors(H, T, T.qg, m, true);  // Q' g
            }
            for (int i = tid; i < m; i += 32) T.z[i] = dxn[i];
            warp_apply_reflectors(H, T, T.z, m, true);      // Q' dx_newton
            if (tid == 0) {
              double* t = T.t;
              double mu = 1e-6;
              double norm_dx = norm_dx_newt;
              int status = 0;
              // factors of H + mu I for the current mu; the reference factors twice per iteration (same matrix)
              if (!tridiag_factor(T, mu, m)) status = 1;
              for (int it = 0; it < a.p.Nmusearch && status == 0; ++it) {
                const double phi = norm_dx - delta;
                tridiag_solve(T, T.z, t, m);  // (H + mu I) \ dx
                double zt = 0.0;
#pragma unroll 4
                for (int i = 0; i < m; ++i) zt = fma(T.z[i], t[i], zt);
                const double phiprime = -(1.0 / norm_dx) * zt;
                mu = mu - (norm_dx / delta) * (phi / phiprime);
                if (!tridiag_factor(T, mu, m)) { status = 1; break; }
                for (int i = 0; i < m; ++i) t[i] = -T.qg[i];
                tridiag_solve(T, t, T.z, m);  // dx = -(H + mu I) \ (Df' fx)
                double zz = 0.0;
#pragma unroll 4
                for (int i = 0; i < m; ++i) zz = fma(T.z[i], T.z[i], zz);
                norm_dx = sqrt(zz);
                if (fabs(norm_dx - delta) / delta <= 1e-4) break;
              }
              mu_status = status;
            }
            __syncwarp();
            warp_apply_reflectors(H, T, T.z, m, false);      // dx = Q z
            for (int i = tid; i < m; i += 32) dx[i] = T.z[i];
          }
          have_qg = true;
          __syncthreads();
          TOC(3);
        }
        const bool singular = mu_status != 0;
        if (singular) { exit_reason = CA_EXIT_SINGULAR; xtol_exit = true; break; }
      }
      block_gemv(Df, dx, t1, m, false);  // Df dx
      const double norm_dx_cur = sqrt(block_dot(dx, dx, m, scratch));
      if (norm_dx_cur < xtol) {
        for (int i = tid; i < m; i += kHookThreads) x[i] += dx[i];
        __syncthreads();
        exit_reason = CA_EXIT_XTOL_HOOK;
        xtol_exit = true;
        break;
      }
      const bool within = norm_dx_newt <= delta;
      for (int i = tid; i < m; i += kHookThreads) xh[i] = x[i] + dx[i];
      __syncthreads();
      { TIC; eval_f(a, invR, xh, fh); TOC(0); }
      ++nf;
      const double r_hook = 0.5 * block_dot(fh, fh, m, scratch);
      const double r_linear = rx + block_dot(fx, t1, m, scratch);
      for (int i = tid; i < m; i += kHookThreads) t2[i] = fx[i] + t1[i];
      __syncthreads();
      const double r_quadratic = 0.5 * block_dot(t2, t2, m, scratch);
      const double dr_hook = rx - r_hook, dr_linear = rx - r_linear, dr_quadratic = rx - r_quadratic;
      if (dr_hook > dr_linear) {
        if (within) break;
        delta = 3 * delta / 2;
        continue;
      } else if (dr_hook < 0.01 * dr_quadratic) {
        delta = delta / 2;
        continue;
      } else if (dr_hook < 0.10 * dr_quadratic) {
        delta = delta / 2;
        break;
      } else if (0.8 * dr_quadratic < dr_hook && dr_hook < 1.2 * dr_quadratic) {
        const double mm = -dr_linear / delta;
        const double c = (r_hook - rx - mm * delta) / (delta * delta);
        const double dnew = -mm / (2 * c);
        if (within) {
        } else if (dnew < 2 * delta / 3) {
          delta = 2 * delta / 3;
        } else if (dnew > 4 * delta / 3) {
          delta = 4 * delta / 3;
        } else {
          delta = dnew;
        }
        break;
      } else {
        break;
      }
    }
    if (L && n < CA_LOG_STEPS) { L->delta[n] = delta; L->hooksteps[n] = nh; }
    if (xtol_exit) break;
    for (int i = tid; i < m; i += kHookThreads) x[i] = xh[i];
    __syncthreads();
  }

#ifdef CA_HOOK_PROFILE
  if (tid == 0 && s == 0)
    printf("[hook profile] total %lld cycles: f %lld  Df %lld  LU %lld  solve %lld  H+g %lld\n", clock64() - tstart, prof[0],
           prof[1], prof[2], prof[3], prof[4]);
#endif
  for (int i = tid; i < m; i += kHookThreads) a.Xout[s + a.nsolve * i] = x[i];
  if (tid == 0) {
    a.success[s] = exit_reason == CA_EXIT_CONVERGED ? 1 : 0;
    if (L) {
      L->exit_reason = exit_reason;
      L->newton_steps = nsteps;
      L->f_evals = nf;
      L->df_evals = ndf;
      L->final_residual = last_res;
      L->final_delta = delta;
      L->device_seconds = 0.0;
    }
  }
}

}  // namespace

void hookstep_solve_large(ca_model* h, const double* R, const double* Xguess, int64_t nsolve, const ca_search_params& p,
                          double* X_out, int32_t* success, ca_solve_log* logs);

extern "C" int32_t ca_hookstep_solve_model(ca_model* h, const double* R, const double* Xguess, int64_t nsolve,
                                           const ca_search_params* params, double* X_out, int32_t* success,
                                           ca_solve_log* logs) {
  return guarded([&] {
    if (!h || !R || !Xguess || !X_out || !success) fail(CA_ERR_INVALID, "ca_hookstep_solve_model: null argument");
    if (nsolve <= 0) fail(CA_ERR_INVALID, "nsolve must be positive");
    ca_search_params p{1e-8, 1e-8, 0.02, 20, 4, 6, 0};  // Hookstep.jl:6-14 defaults
    if (params) p = *params;
    if (p.Nnewton < 0 || p.Nhook < 0 || p.Nmusearch < 0 || !(p.delta > 0)) fail(CA_ERR_INVALID, "bad SearchParams");
    for (int64_t s = 0; s < nsolve; ++s)
      if (!(R[s] != 0.0)) fail(CA_ERR_INVALID, "R must be non-zero");
    h->bind();
    const int m = (int)h->m;
    // Df, H, W (the factorisation workspace; it also holds the mu search's vectors: Tridiag::kDoubles = 14 per mode)
    const size_t wdoubles = std::max((size_t)m * m, (size_t)14 * m);
    const size_t smem = ((size_t)2 * m * m + wdoubles + 9 * (size_t)m + 8) * sizeof(double) + (2 * (size_t)m + 4) * sizeof(int);
    int optin = 0;
    CA_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    cudaFuncAttributes fattr;
    CA_CUDA(cudaFuncGetAttributes(&fattr, hookstep_kernel));  // the kernel's static shared memory counts against the same limit
    if (smem + fattr.sharedSizeBytes > (size_t)optin || getenv("CA_HOOK_LARGE")) {  // beyond the one-CTA solver: global-memory solver, still device-resident
      hookstep_solve_large(h, R, Xguess, nsolve, p, X_out, success, logs);
      return;
    }
    CA_CUDA(cudaFuncSetAttribute(hookstep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    DeviceBuffer<double> dR, dX, dOut;
    DeviceBuffer<int32_t> dSucc;
    DeviceBuffer<ca_solve_log> dLogs;
    dR.upload(R, (size_t)nsolve);
    dX.upload(Xguess, (size_t)nsolve * m);
    dOut.alloc((size_t)nsolve * m);
    dSucc.alloc((size_t)nsolve);
    if (logs) dLogs.alloc((size_t)nsolve);
    SolverArgs a{};
    a.m = m;
    a.L1 = h->L1.ptr;
    a.L2 = h->L2.ptr;
    a.q_rowptr = h->q_rowptr.ptr;
    a.q_ab = h->q_ab.ptr;
    a.q_val = h->q_val.ptr;
    const JacobianPack& J = h->jacobian();
    a.jac_slice_off = J.slice_off.ptr;
    a.jac_out_pos = J.out_pos.ptr;
    a.jac_src = J.src.ptr;
    a.jac_val = J.val.ptr;
    a.jac_nslices = J.nslices;
    a.R = dR.ptr;
    a.Xguess = dX.ptr;
    a.nsolve = nsolve;
    a.p = p;
    a.Xout = dOut.ptr;
    a.success = dSucc.ptr;
    a.logs = logs ? dLogs.ptr : nullptr;
    cudaEvent_t e0, e1;
    CA_CUDA(cudaEventCreate(&e0));
    CA_CUDA(cudaEventCreate(&e1));
    CA_CUDA(cudaEventRecord(e0));
    hookstep_kernel<<<(unsigned)nsolve, kHookThreads, smem>>>(a);
    CA_CUDA(cudaGetLastError());
    count_launch();
    CA_CUDA(cudaEventRecord(e1));
    CA_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    CA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    CA_CUDA(cudaMemcpy(X_out, dOut.ptr, (size_t)nsolve * m * sizeof(double), cudaMemcpyDeviceToHost));
    CA_CUDA(cudaMemcpy(success, dSucc.ptr, (size_t)nsolve * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (logs) {
      CA_CUDA(cudaMemcpy(logs, dLogs.ptr, (size_t)nsolve * sizeof(ca_solve_log), cudaMemcpyDeviceToHost));
      for (int64_t s = 0; s < nsolve; ++s) logs[s].device_seconds = ms * 1e-3;
    }
  });
}
