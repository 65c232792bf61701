// Batched sparse-bilinear evaluation on FP64 tensor instructions (DMMA.8x8x4), sm_100a.
//
// Replaces the COO scatter loop of the reference (src/SparseBilinear.jl:56-65) for ensembles of
// states.  One CTA owns 8*MT states; their coefficients sit in shared memory for the whole pass
// (layout [mode][state], state slot XOR-swizzled by mode parity so that the four modes a k-step
// touches fall into different bank halves).  The operator is streamed from L2 as row tiles
// (pack.h): per k-step a warp forms 4 pairs x 8*MT states of products u_a*w_b -> A fragments,
// loads NT B fragments of operator values (coalesced 256 B each, prefetched 4 k-steps ahead) and
// issues MT*NT DMMAs.  The 8 warps of the CTA split the k-steps of a tile and their partial
// accumulators are summed in a fixed order through shared memory: no atomics, bit-reproducible.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "pack.h"

namespace ca {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kNtMax = 2;
constexpr int kPrefetch = 4;

enum : int { MODE_QUAD = 0, MODE_ORDERED = 1, MODE_JVP = 2 };

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int MT>
__device__ __forceinline__ int swz(int mode) {
  return MT >= 2 ? ((mode & 1) << 3) : 0;
}

template <int MT, int NT, int MODE>
__device__ __forceinline__ void tile_accumulate(const TileDesc& T, const uint32_t* __restrict__ pairs,
                                                const double* __restrict__ vals, const double* xs,
                                                const double* ys, int warp, int lane,
                                                double (&acc)[MT][NT][2]) {
  constexpr int ROW = 8 * MT;
  const int r = lane >> 2, c = lane & 3;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
  const int nmine = T.nks > warp ? (T.nks - warp + kWarps - 1) / kWarps : 0;
  if (nmine == 0) return;
  const uint32_t* pp = pairs + (size_t)T.ks_begin * 4 + c;
  const double* vp = vals + T.vals_off + lane;

  uint32_t pr[kPrefetch];
  double bv[kPrefetch][NT];
#pragma unroll
  for (int u = 0; u < kPrefetch; ++u) {
    if (u < nmine) {
      const int ks = warp + u * kWarps;
      pr[u] = __ldg(pp + (size_t)ks * 4);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) bv[u][nt] = __ldg(vp + ((size_t)ks * NT + nt) * 32);
    }
  }
  for (int i = 0; i < nmine; i += kPrefetch) {
#pragma unroll
    for (int u = 0; u < kPrefetch; ++u) {
      if (i + u < nmine) {
        const uint32_t p = pr[u];
        double b[NT];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) b[nt] = bv[u][nt];
        if (i + u + kPrefetch < nmine) {
          const int ks = warp + (i + u + kPrefetch) * kWarps;
          pr[u] = __ldg(pp + (size_t)ks * 4);
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) bv[u][nt] = __ldg(vp + ((size_t)ks * NT + nt) * 32);
        }
        const int ja = p & 0xffff, kb = p >> 16;
        const double* xa = xs + ja * ROW;
        const double* xb = (MODE == MODE_ORDERED ? ys : xs) + kb * ROW;
        const int sa = swz<MT>(ja), sb = swz<MT>(kb);
        double a[MT];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const int s = mt * 8 + r;
          if (MODE == MODE_JVP) {
            const double* va = ys + ja * ROW;
            const double* vb = ys + kb * ROW;
            a[mt] = xa[s ^ sa] * vb[s ^ sb] + va[s ^ sa] * xb[s ^ sb];
          } else {
            a[mt] = xa[s ^ sa] * xb[s ^ sb];
          }
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
      }
    }
  }
}

template <int MT, int NT>
__device__ __forceinline__ void tile_reduce_store(const TileDesc& T, const int32_t* __restrict__ rows,
                                                  double* red, const double (&acc)[MT][NT][2],
                                                  int warp, int lane, int64_t s0, int64_t nstates,
                                                  double* __restrict__ OUT, int64_t ldo) {
  constexpr int ROW = 8 * MT;
  const int r = lane >> 2, c = lane & 3;
  __syncthreads();  // previous tile's readers are done with `red`
  double* mine = red + warp * (kNtMax * 8 * ROW);
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      mine[(nt * 8 + c * 2 + 0) * ROW + mt * 8 + r] = acc[mt][nt][0];
      mine[(nt * 8 + c * 2 + 1) * ROW + mt * 8 + r] = acc[mt][nt][1];
    }
  __syncthreads();
  for (int o = threadIdx.x; o < NT * 8 * ROW; o += kThreads) {
    const int n = o / ROW, s = o % ROW;
    const int row = __ldg(rows + T.rows_off + n);
    double sum = 0.0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) sum += red[w * (kNtMax * 8 * ROW) + o];
    if (row >= 0 && s0 + s < nstates) OUT[(s0 + s) + ldo * (int64_t)row] = sum;
  }
}

// X, Y: nstates x (nin-1) column-major with leading dimension ldx; OUT: nstates x m, ldo.
template <int MT, int MODE>
__global__ void __launch_bounds__(kThreads, MT == 4 ? 2 : (MT == 2 ? 3 : 4))
bilinear_batched_kernel(const TileDesc* __restrict__ tiles, int ntiles,
                        const uint32_t* __restrict__ pairs, const double* __restrict__ vals,
                        const int32_t* __restrict__ rows, int nin, const double* __restrict__ X,
                        const double* __restrict__ Y, int64_t nstates, int64_t ldx,
                        double* __restrict__ OUT, int64_t ldo) {
  constexpr int ROW = 8 * MT;
  extern __shared__ __align__(16) double smem[];
  double* xs = smem;
  double* ys = MODE == MODE_QUAD ? smem : smem + (size_t)nin * ROW;
  double* red = (MODE == MODE_QUAD ? smem + (size_t)nin * ROW : smem + 2 * (size_t)nin * ROW);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t nblocks = (nstates + ROW - 1) / ROW;

  for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    const int64_t s0 = blk * ROW;
    __syncthreads();  // everyone is done with the previous block's xs
    for (int o = threadIdx.x; o < nin * ROW; o += kThreads) {
      const int j = o / ROW, s = o % ROW;
      double vx = 0.0, vy = 0.0;
      if (j == nin - 1) {
        vx = 1.0;
        vy = MODE == MODE_JVP ? 0.0 : 1.0;  // d(1) = 0 for the Jacobian action of a linear term
      } else if (s0 + s < nstates) {
        vx = X[(s0 + s) + ldx * (int64_t)j];
        if (MODE != MODE_QUAD) vy = Y[(s0 + s) + ldx * (int64_t)j];
      }
      xs[j * ROW + (s ^ swz<MT>(j))] = vx;
      if (MODE != MODE_QUAD) ys[j * ROW + (s ^ swz<MT>(j))] = vy;
    }
    __syncthreads();
    for (int t = 0; t < ntiles; ++t) {
      const TileDesc T = tiles[t];
      if (T.nt == 1) {
        double acc[MT][1][2];
        tile_accumulate<MT, 1, MODE>(T, pairs, vals, xs, ys, warp, lane, acc);
        tile_reduce_store<MT, 1>(T, rows, red, acc, warp, lane, s0, nstates, OUT, ldo);
      } else {
        double acc[MT][2][2];
        tile_accumulate<MT, 2, MODE>(T, pairs, vals, xs, ys, warp, lane, acc);
        tile_reduce_store<MT, 2>(T, rows, red, acc, warp, lane, s0, nstates, OUT, ldo);
      }
    }
  }
}

template <int MT, int MODE>
inline size_t batched_smem_bytes(int nin) {
  const size_t row = 8 * MT;
  return ((MODE == MODE_QUAD ? 1 : 2) * (size_t)nin * row + (size_t)kWarps * kNtMax * 8 * row) *
         sizeof(double);
}

}  // namespace ca
