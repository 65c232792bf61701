// Dense LU with partial pivoting on the device, for column-major m x m systems in global memory (m up to ~10^4):
// blocked right-looking factorisation -- a panel of kLuNb columns by one CTA (pivot search, interchanges inside the
// panel, rank-1 updates), row interchanges of the other columns, triangular solve of the block row, trailing update by
// the tiled DGEMM of dgemm.cuh -- and blocked forward / backward substitution by one CTA.  Used by the large-model
// Newton-hookstep (hookstep_large.cu: dx = -Df \ fx, (H + mu I) \ g) and as the preconditioner of the Newton-Krylov
// solver (newton_krylov.cu: (L1 + L2/R)^-1).  Checked against LAPACK through ca_selftest_lu_solve.
#pragma once
#include <algorithm>
#include <cstdint>

#include "common.h"
#include "dgemm.cuh"

namespace ca {
namespace lu {

constexpr int kLuNb = 16;         // panel width of the blocked LU
constexpr int kLuThreads = 1024;  // the panel and the substitution kernels are single-CTA
constexpr int kSolveNb = 32;      // block size of the triangular substitutions

// ---- panel factorisation: columns [k0, k0+nb) of rows [k0, m), partial pivoting, unit lower factor stored in place
static __global__ void __launch_bounds__(kLuThreads)
lu_panel_kernel(double* __restrict__ A, int64_t ld, int m, int k0, int nb, int32_t* __restrict__ piv,
                int32_t* __restrict__ singular, int pivoting) {
  __shared__ double red_val[32];
  __shared__ int red_idx[32];
  __shared__ double prow[kLuNb];
  __shared__ int psel;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int c = 0; c < nb; ++c) {
    const int col = k0 + c;
    if (col >= m) break;
    double* Ac = A + ld * col;
    // pivot: first row of maximal magnitude at or below the diagonal
    double best = -1.0;
    int bi = col;
    if (pivoting) {
      for (int r = col + tid; r < m; r += kLuThreads) {
        const double v = fabs(Ac[r]);
        if (v > best) { best = v; bi = r; }
      }
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, d);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      if (lane == 0) { red_val[warp] = best; red_idx[warp] = bi; }
      __syncthreads();
      if (warp == 0) {
        best = red_val[lane];
        bi = red_idx[lane];
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
          const double ov = __shfl_xor_sync(0xffffffffu, best, d);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
          if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) psel = bi;
      }
      __syncthreads();
    } else if (tid == 0) {
      psel = col;
    }
    if (!pivoting) __syncthreads();
    const int p = psel;
    if (tid == 0) piv[col] = p;
    // swap rows col and p inside the panel, cache the pivot row
    if (tid < nb) {
      double* a = A + ld * (k0 + tid);
      const double up = a[p];
      if (p != col) {
        a[p] = a[col];
        a[col] = up;
      }
      prow[tid] = up;
    }
    __syncthreads();
    const double pv = prow[c];
    if (pv == 0.0) {
      if (tid == 0) *singular = 1;
      continue;  // (uniform: every thread reads the same pivot)
    }
    const int ncols = min(nb, m - k0);
    for (int r = col + 1 + tid; r < m; r += kLuThreads) {
      const double l = Ac[r] / pv;
      Ac[r] = l;
      for (int j = c + 1; j < ncols; ++j) A[ld * (k0 + j) + r] -= l * prow[j];
    }
    __syncthreads();
  }
}

// row interchanges of the panel applied to the columns outside it
static __global__ void lu_swap_kernel(double* __restrict__ A, int64_t ld, int m, int k0, int nb, const int32_t* __restrict__ piv) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m - nb) return;
  if (j >= k0) j += nb;  // skip the panel's own columns
  double* a = A + ld * j;
  for (int c = 0; c < nb && k0 + c < m; ++c) {
    const int p = piv[k0 + c];
    if (p != k0 + c) {
      const double t = a[k0 + c];
      a[k0 + c] = a[p];
      a[p] = t;
    }
  }
}

// block row of U: A12 <- L11^-1 A12 (unit lower L11), one thread per column right of the panel
static __global__ void lu_trsm_kernel(double* __restrict__ A, int64_t ld, int m, int k0, int nb) {
  __shared__ double L[kLuNb][kLuNb + 1];
  for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) L[e % nb][e / nb] = A[ld * (k0 + e / nb) + k0 + e % nb];
  __syncthreads();
  const int j = k0 + nb + blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  double* a = A + ld * j + k0;
  double x[kLuNb];
#pragma unroll
  for (int r = 0; r < kLuNb; ++r) x[r] = r < nb ? a[r] : 0.0;
#pragma unroll
  for (int c = 0; c < kLuNb; ++c)
#pragma unroll
    for (int r = c + 1; r < kLuNb; ++r)
      if (r < nb) x[r] -= L[r][c] * x[c];
#pragma unroll
  for (int r = 0; r < kLuNb; ++r)
    if (r < nb) a[r] = x[r];
}

// b <- U^-1 L^-1 P b, blocked substitutions by one CTA (the vectors are L2-resident)
static __global__ void __launch_bounds__(kLuThreads)
lu_solve_kernel(const double* __restrict__ A, int64_t ld, int m, const int32_t* __restrict__ piv, double* __restrict__ B,
                int64_t ldb, int nrhs, double scale) {
  __shared__ double T[kSolveNb][kSolveNb + 1];
  __shared__ double xs[kSolveNb];
  __shared__ int32_t pchunk[kLuThreads];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int rhs = blockIdx.x; rhs < nrhs; rhs += gridDim.x) {  // one right-hand side at a time per CTA
  double* b = B + (int64_t)rhs * ldb;
  __syncthreads();
  for (int c0 = 0; c0 < m; c0 += kLuThreads) {  // the interchanges are sequential; the pivot indices are staged in bulk
    if (c0 + tid < m) pchunk[tid] = piv[c0 + tid];
    __syncthreads();
    if (tid == 0)
      for (int c = c0; c < min(m, c0 + kLuThreads); ++c) {
        const int p = pchunk[c - c0];
        if (p != c) {
          const double t = b[c];
          b[c] = b[p];
          b[p] = t;
        }
      }
    __syncthreads();
  }
  // forward: unit lower
  for (int k0 = 0; k0 < m; k0 += kSolveNb) {
    const int nb = min(kSolveNb, m - k0);
    {
      const int r = tid % kSolveNb, c = tid / kSolveNb;
      T[r][c] = (r < nb && c < nb) ? A[ld * (k0 + c) + k0 + r] : 0.0;
    }
    __syncthreads();
    if (warp == 0) {
      double x = lane < nb ? b[k0 + lane] : 0.0;
      for (int c = 0; c < nb; ++c) {
        const double xc = __shfl_sync(0xffffffffu, x, c);
        if (lane > c) x -= T[lane][c] * xc;
      }
      xs[lane] = x;
      if (lane < nb) b[k0 + lane] = x;
    }
    __syncthreads();
    for (int r = k0 + nb + tid; r < m; r += kLuThreads) {
      double acc = 0.0;
      for (int c = 0; c < nb; ++c) acc = fma(A[ld * (k0 + c) + r], xs[c], acc);
      b[r] -= acc;
    }
    __syncthreads();
  }
  // backward: upper with diagonal
  for (int k1 = m; k1 > 0; k1 -= kSolveNb) {
    const int k0 = max(0, k1 - kSolveNb), nb = k1 - k0;
    {
      const int r = tid % kSolveNb, c = tid / kSolveNb;
      T[r][c] = (r < nb && c < nb) ? A[ld * (k0 + c) + k0 + r] : 0.0;
    }
    __syncthreads();
    if (warp == 0) {
      double x = lane < nb ? b[k0 + lane] : 0.0;
      for (int c = nb - 1; c >= 0; --c) {
        if (lane == c) x /= T[c][c];
        const double xc = __shfl_sync(0xffffffffu, x, c);
        if (lane < c) x -= T[lane][c] * xc;
      }
      xs[lane] = x;
      if (lane < nb) b[k0 + lane] = x;
    }
    __syncthreads();
    for (int r = tid; r < k0; r += kLuThreads) {
      double acc = 0.0;
      for (int c = 0; c < nb; ++c) acc = fma(A[ld * (k0 + c) + r], xs[c], acc);
      b[r] -= acc;
    }
    __syncthreads();
  }
  if (scale != 1.0)
    for (int r = tid; r < m; r += kLuThreads) b[r] *= scale;
  }
}


// A (column-major m x m, leading dimension m) is overwritten by its factors; piv: m row interchanges (LAPACK getrf
// convention); *flag (device, zeroed by the caller) is set to 1 if a pivot is exactly zero.
inline void factor(double* A, int m, int32_t* piv, int32_t* flag, bool pivoting, cudaStream_t st) {
  for (int k0 = 0; k0 < m; k0 += kLuNb) {
    const int nb = std::min(kLuNb, m - k0), k1 = k0 + nb;
    lu_panel_kernel<<<1, kLuThreads, 0, st>>>(A, m, m, k0, nb, piv, flag, pivoting ? 1 : 0);
    if (pivoting && m > nb) lu_swap_kernel<<<(unsigned)((m - nb + 255) / 256), 256, 0, st>>>(A, m, m, k0, nb, piv);
    if (k1 < m) {
      lu_trsm_kernel<<<(unsigned)((m - k1 + 127) / 128), 128, 0, st>>>(A, m, m, k0, nb);
      launch_dgemm(m - k1, m - k1, nb, -1.0, colmajor(A + k1 + (size_t)m * k0, m), colmajor(A + k0 + (size_t)m * k1, m), 1.0,
                   A + k1 + (size_t)m * k1, 1, m, st);
    }
    count_launch(3);
  }
  CA_CUDA(cudaGetLastError());
}

// b <- scale * A^-1 b with the factors of factor()
inline void solve(const double* A, int m, const int32_t* piv, double* b, double scale, cudaStream_t st) {
  lu_solve_kernel<<<1, kLuThreads, 0, st>>>(A, m, m, piv, b, m, 1, scale);
  CA_CUDA(cudaGetLastError());
  count_launch();
}

// B (column-major m x nrhs, leading dimension m) <- A^-1 B: the single-CTA substitution, one CTA per right-hand side
// at a time (the columns are independent); nrhs = m with B = I gives the explicit inverse.
inline void solve_many(const double* A, int m, const int32_t* piv, double* B, int nrhs, cudaStream_t st) {
  if (nrhs <= 0) return;
  lu_solve_kernel<<<(unsigned)std::min(nrhs, 148 * 2), kLuThreads, 0, st>>>(A, m, m, piv, B, m, nrhs, 1.0);
  CA_CUDA(cudaGetLastError());
  count_launch();
}

}  // namespace lu
}  // namespace ca
