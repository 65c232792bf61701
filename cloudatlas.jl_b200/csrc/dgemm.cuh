// Small in-tree FP64 GEMM on strided views:  C = alpha * A * B + beta * C.
//
// For the dense steps AROUND the hot kernels -- basis mixing Phi = Psi C of the dense assembly (assemble.cu), Df'Df and
// the trailing updates of the blocked LU of the large-model Newton-hookstep (hookstep_large.cu).  Those are a few
// 10^9 .. 10^11 flops next to 10^13 of the contraction they prepare, so this is a plain shared-memory tiled DFMA
// kernel (64 x 64 x 16 tiles, 4 x 4 outputs per thread), not a DMMA pipeline: fixed summation order (k ascending),
// any strides (a transpose is a view), any sizes.
#pragma once
#include <cstdint>

#include "common.h"

namespace ca {

struct MatView {  // element (i, j) at p[i * rs + j * cs]
  const double* p;
  int64_t rs, cs;
};
inline MatView colmajor(const double* p, int64_t ld) { return MatView{p, 1, ld}; }
inline MatView rowmajor(const double* p, int64_t ld) { return MatView{p, ld, 1}; }
inline MatView transposed(MatView v) { return MatView{v.p, v.cs, v.rs}; }

constexpr int kGemmTile = 64, kGemmK = 16, kGemmThreads = 256;

static __global__ void __launch_bounds__(kGemmThreads)
dgemm_kernel(int M, int N, int K, double alpha, MatView A, MatView B, double beta, double* __restrict__ C,
                    int64_t rsC, int64_t csC) {
  __shared__ double As[kGemmK][kGemmTile + 4];
  __shared__ double Bs[kGemmK][kGemmTile + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.y * kGemmTile, j0 = blockIdx.x * kGemmTile;
  double acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
  const bool a_i_fast = A.rs <= A.cs, b_j_fast = B.cs <= B.rs;  // which index runs fastest in memory: coalesce along it
  for (int k0 = 0; k0 < K; k0 += kGemmK) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = threadIdx.x + kGemmThreads * q;
      const int i = a_i_fast ? e % kGemmTile : e / kGemmK, k = a_i_fast ? e / kGemmTile : e % kGemmK;
      As[k][i] = (i0 + i < M && k0 + k < K) ? A.p[(int64_t)(i0 + i) * A.rs + (int64_t)(k0 + k) * A.cs] : 0.0;
      const int j = b_j_fast ? e % kGemmTile : e / kGemmK, kb = b_j_fast ? e / kGemmTile : e % kGemmK;
      Bs[kb][j] = (j0 + j < N && k0 + kb < K) ? B.p[(int64_t)(k0 + kb) * B.rs + (int64_t)(j0 + j) * B.cs] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kGemmK; ++k) {
      double a[4], b[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) a[r] = As[k][ty * 4 + r];
#pragma unroll
      for (int c = 0; c < 4; ++c) b[c] = Bs[k][tx * 4 + c];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fma(a[r], b[c], acc[r][c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int i = i0 + ty * 4 + r, j = j0 + tx * 4 + c;
      if (i < M && j < N) {
        double* out = C + (int64_t)i * rsC + (int64_t)j * csC;
        *out = beta == 0.0 ? alpha * acc[r][c] : fma(alpha, acc[r][c], beta * *out);
      }
    }
}

// C (M x N, element strides rsC / csC) = alpha * A (M x K) * B (K x N) + beta * C
inline void launch_dgemm(int M, int N, int K, double alpha, MatView A, MatView B, double beta, double* C, int64_t rsC,
                         int64_t csC, cudaStream_t st) {
  if (M <= 0 || N <= 0) return;
  const dim3 grid((unsigned)((N + kGemmTile - 1) / kGemmTile), (unsigned)((M + kGemmTile - 1) / kGemmTile));
  dgemm_kernel<<<grid, kGemmThreads, 0, st>>>(M, N, K, alpha, A, B, beta, C, rsC, csC);
  CA_CUDA(cudaGetLastError());
  count_launch();
}

}  // namespace ca
