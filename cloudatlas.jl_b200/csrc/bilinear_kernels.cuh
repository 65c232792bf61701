// Batched sparse-bilinear evaluation on FP64 tensor instructions (DMMA.8x8x4), sm_100a.
//
// Replaces the COO scatter loop of the reference (src/SparseBilinear.jl:56-65) for ensembles of
// states.  One CTA owns 8*MT states; their coefficients sit in shared memory for the whole pass
// (layout [mode][state], state slot XOR-swizzled by mode parity so that the four modes a k-step
// touches fall into different bank halves).  The operator is streamed from L2 as row tiles
// (pack.h): per k-step a warp forms 4 pairs x 8*MT states of products u_a*w_b -> A fragments,
// loads NT B fragments of operator values (coalesced 256 B each) and issues MT*NT DMMAs.
// Three kernels share these k-steps:
//   bilinear_warptile_kernel  whole tiles per warp (LPT schedule from the packer): no reduction, no per-tile barrier;
//   bilinear_batched_kernel   the warps of a CTA split every tile's k-steps and sum their partial accumulators in a
//                             fixed order through shared memory (unbalanced tile schedules);
//   bilinear_windowed_kernel  large m: only the rows of the state tile that a window of k-steps touches are resident,
//                             everything streams through one cp.async pipeline.
// No atomics anywhere: results are bit-reproducible.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "pack.h"

namespace ca {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kNtMax = 2;
constexpr int kPrefetch = 4;       // operator values: k-steps in flight per warp
constexpr int kPrefetchPairs = 4;  // pair words are needed first in a k-step (address math): fetched further ahead

enum : int { MODE_QUAD = 0, MODE_ORDERED = 1, MODE_JVP = 2 };

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// State slot swizzle inside a mode's row of the shared tile.  An LDS.64 of a warp is served in two
// half-warp phases; in a phase the four pair columns read four modes x 32 B.  XOR-ing bits 2-3 of the
// state slot with the low bits of the mode index sends modes with distinct (mode & 3) -- e.g. the
// consecutive second factors of an a-major pair list -- to four distinct 8-bank groups: conflict-free.
template <int MT>
__host__ __device__ __forceinline__ int swz(int mode) {
  return MT >= 2 ? ((mode & 3) << 2) : ((mode & 2) << 1);  // MT == 1: rows are 64 B, bit 3 comes from the row
}

// Address word of a factor (built on the host per MT, see DevicePack::pairs_for): byte offset of the
// mode's row in the shared tile with the swizzle folded into bits 5-6, so that the lane's four loads are
//   [w ^ r8], [(w ^ r8) ^ 64], [.. + 128], [.. ^ 64 + 128]        (r8 = 8 * (lane >> 2))
// Bit 0 of the first factor's word is the "same first factors as the previous k-step" flag.
template <int MT>
__host__ __device__ __forceinline__ uint32_t factor_word(int mode) {
  return (uint32_t)mode * (64u * MT) | ((uint32_t)swz<MT>(mode) << 3);
}
constexpr uint32_t kSameA = 1u;
// windowed packs: rows are window-local (< 256), so both address words of a pair fit 16 bits and a pair is one word
__host__ __device__ __forceinline__ uint32_t pack_window_pair(int a_local, int b_local, bool same_a) {
  return (factor_word<4>(a_local) | (same_a ? kSameA : 0u)) | (factor_word<4>(b_local) << 16);
}

template <int NT>
struct Prefetch {
  uint2 pw[kPrefetchPairs];
  double bv[kPrefetch][NT];
};

__device__ __forceinline__ void prefetch_pair(uint2& pw, const uint2* pp, int step) { pw = __ldg(pp + (size_t)step * 4); }

template <int NT>
__device__ __forceinline__ void prefetch_vals(double (&bv)[NT], const double* vp, int step) {
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) bv[nt] = __ldg(vp + ((size_t)step * NT + nt) * 32);
}

// This warp's contiguous share [k0, k1) of a range of nks k-steps.
template <int NW = kWarps>
__device__ __forceinline__ void warp_share(int nks, int warp, int& k0, int& k1) {
  if (warp < 0) {  // the whole range belongs to the calling warp (whole-tile-per-warp schedule)
    k0 = 0;
    k1 = nks;
    return;
  }
  k0 = (int)(((int64_t)nks * warp) / NW);
  k1 = (int)(((int64_t)nks * (warp + 1)) / NW);
}

// Loads the first k-steps of this warp's share of a k-step range (a whole tile, or one window of it);
// issued before the previous range's barriers so that the L2 latency hides behind them.
// pairs0 / vals0 point at the range's first k-step.
template <int NT>
__device__ __forceinline__ void range_prologue(const uint2* __restrict__ pairs0, const double* __restrict__ vals0,
                                               int nks, int warp, int lane, Prefetch<NT>& pf) {
  int k0, k1;
  warp_share(nks, warp, k0, k1);
  const int n = k1 - k0;
  if (n <= 0) return;
  const uint2* pp = pairs0 + (size_t)k0 * 4 + (lane & 3);
  const double* vp = vals0 + (size_t)k0 * NT * 32 + lane;
#pragma unroll
  for (int u = 0; u < kPrefetchPairs; ++u) prefetch_pair(pf.pw[u], pp, min(u, n - 1));
#pragma unroll
  for (int u = 0; u < kPrefetch; ++u) prefetch_vals<NT>(pf.bv[u], vp, min(u, n - 1));
  pf.pw[0].x &= ~kSameA;  // the first k-step of a share always loads its first factors
}

template <int MT>
__device__ __forceinline__ double lds_state(const char* base, uint32_t w, int mt) {
  // mt is a compile-time constant after unrolling: (w ^ 64*(mt&1)) + 128*(mt>>1)
  return *reinterpret_cast<const double*>(base + ((mt & 1) ? (w ^ 64u) : w) + 128u * (mt >> 1));
}

template <int MT, int NT, int MODE>
__device__ __forceinline__ void kstep(const uint2 pw, const double (&bv)[NT], const char* xs, const char* ys,
                                      uint32_t r8, double (&xa)[MT], double (&va)[MT],
                                      double (&acc)[MT][NT][2]) {
  if (!(pw.x & kSameA)) {  // warp-uniform
    const uint32_t wa = (pw.x & ~7u) ^ r8;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      xa[mt] = lds_state<MT>(xs, wa, mt);
      if (MODE == MODE_JVP) va[mt] = lds_state<MT>(ys, wa, mt);
    }
  }
  const uint32_t wb = pw.y ^ r8;
  double a[MT];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    if (MODE == MODE_JVP) a[mt] = xa[mt] * lds_state<MT>(ys, wb, mt) + va[mt] * lds_state<MT>(xs, wb, mt);
    else a[mt] = xa[mt] * lds_state<MT>(MODE == MODE_ORDERED ? ys : xs, wb, mt);
  }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], bv[nt]);
}

// Factored quadratic evaluation on packs in kOrderFactored (pack.cpp: every k-step has ONE first factor a):
//   out_i = sum_a x_a * ( sum_b V_iab x_b )
// The DMMAs accumulate y = sum_b V x_b with the plain second factors as A fragments -- no pair products -- and at the
// end of a run of equal a the run's result enters the tile's accumulators with one DFMA per accumulator register:
// MT * NT * 2 DFMAs per run instead of MT DMULs per k-step (runs are ~9 k-steps at 307 modes).
template <int MT, int NT>
__device__ __forceinline__ void flush_run(const double (&xa)[MT], double (&y)[MT][NT][2], double (&acc)[MT][NT][2]) {
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        acc[mt][nt][e] = fma(xa[mt], y[mt][nt][e], acc[mt][nt][e]);  // accumulator row = state 8 mt + (lane >> 2), as xa
        y[mt][nt][e] = 0.0;
      }
}

template <int MT, int NT>
__device__ __forceinline__ void kstep_factored(const uint2 pw, const double (&bv)[NT], const char* xs, uint32_t r8,
                                               double (&xa)[MT], double (&y)[MT][NT][2], double (&acc)[MT][NT][2]) {
  if (!(pw.x & kSameA)) {  // warp-uniform: a new run
    flush_run<MT, NT>(xa, y, acc);
    const uint32_t wa = (pw.x & ~7u) ^ r8;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) xa[mt] = lds_state<MT>(xs, wa, mt);
  }
  const uint32_t wb = pw.y ^ r8;
  double a[MT];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) a[mt] = lds_state<MT>(xs, wb, mt);
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) dmma884(y[mt][nt][0], y[mt][nt][1], a[mt], bv[nt]);
}

template <int MT, int NT, int MODE>
__device__ __forceinline__ void range_accumulate(const uint2* __restrict__ pairs0, const double* __restrict__ vals0,
                                                 int nks, const double* xs, const double* ys, int warp, int lane,
                                                 Prefetch<NT>& pf, double (&acc)[MT][NT][2]) {
  const uint32_t r8 = 8u * (lane >> 2);
  int k0, k1;
  warp_share(nks, warp, k0, k1);
  const int n = k1 - k0;
  if (n <= 0) return;
  const uint2* pp = pairs0 + (size_t)k0 * 4 + (lane & 3);
  const double* vp = vals0 + (size_t)k0 * NT * 32 + lane;
  const char* xsb = reinterpret_cast<const char*>(xs);
  const char* ysb = reinterpret_cast<const char*>(ys);
  double xa[MT], va[MT];  // first factors, kept across k-steps while the pair list stays on the same a
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) xa[mt] = va[mt] = 0.0;
  int i = 0;
  for (; i + kPrefetchPairs <= n; i += kPrefetchPairs) {  // branch-free body: refills clamp to the last k-step
#pragma unroll
    for (int u = 0; u < kPrefetchPairs; ++u) {
      kstep<MT, NT, MODE>(pf.pw[u], pf.bv[u % kPrefetch], xsb, ysb, r8, xa, va, acc);
      prefetch_pair(pf.pw[u], pp, min(i + u + kPrefetchPairs, n - 1));
      prefetch_vals<NT>(pf.bv[u % kPrefetch], vp, min(i + u + kPrefetch, n - 1));
    }
  }
#pragma unroll
  for (int u = 0; u < kPrefetchPairs - 1; ++u) {
    if (i + u < n) {
      kstep<MT, NT, MODE>(pf.pw[u], pf.bv[u % kPrefetch], xsb, ysb, r8, xa, va, acc);
      if (u + kPrefetch < kPrefetchPairs - 1) prefetch_vals<NT>(pf.bv[u % kPrefetch], vp, min(i + u + kPrefetch, n - 1));
    }
  }
}

// n-tiles [NT0, NT0 + NTC) of an accumulator with NTA n-tiles (NTC <= kNtMax: the reduction buffer holds two)
template <int MT, int NTA, int NT0, int NTC, int NW>
__device__ __forceinline__ void tile_reduce_store_part(int rows_off, const int32_t* __restrict__ rows,
                                                       double* red, const double (&acc)[MT][NTA][2],
                                                       int warp, int lane, int64_t s0, int64_t nstates,
                                                       double* __restrict__ OUT, int64_t ldo) {
  static_assert(NTC <= kNtMax && NT0 + NTC <= NTA, "slice of the accumulator");
  constexpr int ROW = 8 * MT;
  const int r = lane >> 2, c = lane & 3;
  __syncthreads();  // previous tile's readers are done with `red`
  double* mine = red + warp * (kNtMax * 8 * ROW);
#pragma unroll
  for (int nt = 0; nt < NTC; ++nt)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      mine[(nt * 8 + c * 2 + 0) * ROW + mt * 8 + r] = acc[mt][NT0 + nt][0];
      mine[(nt * 8 + c * 2 + 1) * ROW + mt * 8 + r] = acc[mt][NT0 + nt][1];
    }
  __syncthreads();
  for (int o = threadIdx.x; o < NTC * 8 * ROW; o += NW * 32) {
    const int n = o / ROW, s = o % ROW;
    const int row = __ldg(rows + rows_off + NT0 * 8 + n);
    double sum = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) sum += red[w * (kNtMax * 8 * ROW) + o];
    if (row >= 0 && s0 + s < nstates) OUT[(s0 + s) + ldo * (int64_t)row] = sum;
  }
}

template <int MT, int NT, int NW = kWarps>
__device__ __forceinline__ void tile_reduce_store(int rows_off, const int32_t* __restrict__ rows,
                                                  double* red, const double (&acc)[MT][NT][2],
                                                  int warp, int lane, int64_t s0, int64_t nstates,
                                                  double* __restrict__ OUT, int64_t ldo) {
  if constexpr (NT <= kNtMax) {
    tile_reduce_store_part<MT, NT, 0, NT, NW>(rows_off, rows, red, acc, warp, lane, s0, nstates, OUT, ldo);
  } else {  // three n-tiles: two passes through the two-tile reduction buffer
    tile_reduce_store_part<MT, NT, 0, kNtMax, NW>(rows_off, rows, red, acc, warp, lane, s0, nstates, OUT, ldo);
    tile_reduce_store_part<MT, NT, kNtMax, NT - kNtMax, NW>(rows_off, rows, red, acc, warp, lane, s0, nstates, OUT, ldo);
  }
}

// All tiles [tb, te) have NT n-tiles (the packer orders tiles by nt).
template <int MT, int NT, int MODE>
__device__ __forceinline__ void run_tiles(const TileDesc* __restrict__ tiles, int tb, int te,
                                          const uint2* __restrict__ pairs, const double* __restrict__ vals,
                                          const int32_t* __restrict__ rows, const double* xs, const double* ys,
                                          double* red, int warp, int lane, int64_t s0, int64_t nstates,
                                          double* __restrict__ OUT, int64_t ldo) {
  if (tb >= te) return;
  Prefetch<NT> pf;
  TileDesc T = tiles[tb];
  range_prologue<NT>(pairs + (size_t)T.ks_begin * 4, vals + T.vals_off, T.nks, warp, lane, pf);
  for (int t = tb; t < te; ++t) {
    double acc[MT][NT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
    range_accumulate<MT, NT, MODE>(pairs + (size_t)T.ks_begin * 4, vals + T.vals_off, T.nks, xs, ys, warp, lane, pf, acc);
    const TileDesc done = T;
    if (t + 1 < te) {
      T = tiles[t + 1];
      range_prologue<NT>(pairs + (size_t)T.ks_begin * 4, vals + T.vals_off, T.nks, warp, lane, pf);
    }
    tile_reduce_store<MT, NT>(done.rows_off, rows, red, acc, warp, lane, s0, nstates, OUT, ldo);
  }
}

// ------------------------------------------------------------------------------------------------
// Whole-tile-per-warp variant: the packer assigns complete tiles to warps (LPT), so a warp owns every k-step of its
// tiles, keeps the accumulators to the end and writes its 32 x 8nt outputs straight from the C fragments.  No
// cross-warp reduction, no per-tile barrier, no reduction buffer; the only barriers are the two around the load
// of a state tile.
struct WarpSchedule {
  int32_t off[9];
  int32_t nt2[8];
  int32_t ks2[8], ks1[8];  // k-steps of the warp's nt == 2 and nt == 1 streams (contiguous in memory, pack.h)
};

// Pair words of a warp's stream go through a small per-warp ring in shared memory, filled with cp.async 24 k-steps
// ahead and across tile boundaries: they are the first thing a k-step needs (address arithmetic), and fetching them
// four k-steps ahead into registers left ~14 % of the issue slots waiting on L2.  Operator values are consumed last
// in a k-step (by the DMMAs) and stay register-prefetched.
constexpr int kPairBatch = 8;                    // k-steps per cp.async group (256 B)
constexpr int kPairBatches = 4;                  // ring = 4 batches = 32 k-steps = 1 KB per warp
constexpr int kPairRingBytes = kPairBatch * kPairBatches * 32;

struct PairRing {
  const uint2* stream;  // first pair word of the warp's stream
  char* ring;           // this warp's kPairRingBytes
  int total;            // k-steps in the stream
  int g;                // next k-step to consume

  __device__ __forceinline__ void issue_batch(int batch, int lane) const {
    if (batch * kPairBatch < total && lane < 16) {  // 16 lanes x 16 B = one batch; reading past the stream's end by
      const uint32_t d = (uint32_t)__cvta_generic_to_shared(ring + (batch % kPairBatches) * (kPairBatch * 32) + lane * 16);
      const char* src = reinterpret_cast<const char*>(stream + (size_t)batch * kPairBatch * 4) + lane * 16;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src) : "memory");  // < 256 B stays inside the pairs array + pad
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  }
  __device__ __forceinline__ void begin(int lane) {
    g = 0;
    __syncwarp();
#pragma unroll
    for (int b = 0; b < kPairBatches - 1; ++b) issue_batch(b, lane);
  }
  // pair word of k-step g for this lane's column; refills one batch ahead at batch boundaries
  __device__ __forceinline__ uint2 next(int lane) {
    if ((g & (kPairBatch - 1)) == 0) {  // warp-uniform
      __syncwarp();                     // the batch slot refilled now was consumed by every lane
      issue_batch(g / kPairBatch + kPairBatches - 1, lane);
      asm volatile("cp.async.wait_group %0;\n" ::"n"(kPairBatches - 1) : "memory");
      __syncwarp();                     // batch g / kPairBatch has landed for all lanes
    }
    const uint2 w = reinterpret_cast<const uint2*>(ring)[(g & (kPairBatch * kPairBatches - 1)) * 4 + (lane & 3)];
    ++g;
    return w;
  }
  __device__ __forceinline__ void end() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
};

// values-only register prefetch for the ring variant
template <int NT>
struct ValPrefetch {
  double bv[kPrefetch][NT];
};

template <int MT, int NT, int MODE, bool FACT = false>
__device__ __forceinline__ void tile_accumulate_ring(const double* __restrict__ vals0, int nks, const double* xs,
                                                     const double* ys, int lane, PairRing& pr, ValPrefetch<NT>& pf,
                                                     double (&acc)[MT][NT][2]) {
  const uint32_t r8 = 8u * (lane >> 2);
  const double* vp = vals0 + lane;
  const char* xsb = reinterpret_cast<const char*>(xs);
  const char* ysb = reinterpret_cast<const char*>(ys);
  double xa[MT], va[MT];
  double y[MT][FACT ? NT : 1][2];  // FACT: the open run's sum_b V x_b
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    xa[mt] = va[mt] = 0.0;
#pragma unroll
    for (int nt = 0; nt < (FACT ? NT : 1); ++nt) y[mt][nt][0] = y[mt][nt][1] = 0.0;
  }
  auto one = [&](uint2 pw, const double (&bv)[NT]) {
    if constexpr (FACT) kstep_factored<MT, NT>(pw, bv, xsb, r8, xa, y, acc);
    else kstep<MT, NT, MODE>(pw, bv, xsb, ysb, r8, xa, va, acc);
  };
  int i = 0;
  for (; i + kPrefetch <= nks; i += kPrefetch) {
#pragma unroll
    for (int u = 0; u < kPrefetch; ++u) {
      uint2 pw = pr.next(lane);
      if (i + u == 0) pw.x &= ~kSameA;  // first k-step of a tile loads its first factors
      one(pw, pf.bv[u]);
      prefetch_vals<NT>(pf.bv[u], vp, min(i + u + kPrefetch, nks - 1));
    }
  }
#pragma unroll
  for (int u = 0; u < kPrefetch - 1; ++u) {
    if (i + u < nks) {
      uint2 pw = pr.next(lane);
      if (i + u == 0) pw.x &= ~kSameA;
      one(pw, pf.bv[u]);
    }
  }
  if constexpr (FACT) flush_run<MT, NT>(xa, y, acc);  // the tile's last run
}


template <int MT, int NT>
__device__ __forceinline__ void tile_store_direct(const TileDesc& T, const int32_t* __restrict__ rows,
                                                  const double (&acc)[MT][NT][2], int lane, int64_t s0,
                                                  int64_t nstates, double* __restrict__ OUT, int64_t ldo) {
  const int r = lane >> 2, c = lane & 3;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int row = __ldg(rows + T.rows_off + nt * 8 + 2 * c + e);
      if (row < 0) continue;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int64_t s = s0 + mt * 8 + r;
        if (s < nstates) OUT[s + ldo * (int64_t)row] = acc[mt][nt][e];
      }
    }
}

template <int MT, int NT, int MODE, bool FACT = false>
__device__ __forceinline__ void run_my_tiles(const TileDesc* __restrict__ tiles, const int32_t* __restrict__ sched,
                                             int b, int e, int total, const uint2* __restrict__ pairs,
                                             const double* __restrict__ vals, const int32_t* __restrict__ rows,
                                             const double* xs, const double* ys, char* ring, int lane, int64_t s0,
                                             int64_t nstates, double* __restrict__ OUT, int64_t ldo) {
  if (b >= e) return;
  TileDesc T = tiles[__ldg(sched + b)];
  PairRing pr{pairs + (size_t)T.ks_begin * 4, ring, total, 0};
  pr.begin(lane);
  ValPrefetch<NT> pf;
  auto prologue = [&](const TileDesc& Tn) {
    const double* vp = vals + Tn.vals_off + lane;
#pragma unroll
    for (int u = 0; u < kPrefetch; ++u)
      if (Tn.nks > 0) prefetch_vals<NT>(pf.bv[u], vp, min(u, Tn.nks - 1));
  };
  prologue(T);
  for (int t = b; t < e; ++t) {
    double acc[MT][NT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
    tile_accumulate_ring<MT, NT, MODE, FACT>(vals + T.vals_off, T.nks, xs, ys, lane, pr, pf, acc);
    const TileDesc done = T;
    if (t + 1 < e) {
      T = tiles[__ldg(sched + t + 1)];
      prologue(T);
    }
    tile_store_direct<MT, NT>(done, rows, acc, lane, s0, nstates, OUT, ldo);
  }
  pr.end();
}

template <int MT, int MODE, bool FACT = false>
__global__ void __launch_bounds__(kThreads, 2)
bilinear_warptile_kernel(const TileDesc* __restrict__ tiles, const int32_t* __restrict__ sched, WarpSchedule ws,
                         const uint2* __restrict__ pairs, const double* __restrict__ vals,
                         const int32_t* __restrict__ rows, int nin, const double* __restrict__ X,
                         const double* __restrict__ Y, int64_t nstates, int64_t ldx, double* __restrict__ OUT,
                         int64_t ldo) {
  constexpr int ROW = 8 * MT;
  extern __shared__ __align__(16) double smem[];
  double* xs = smem;
  double* ys = MODE == MODE_QUAD ? smem : smem + (size_t)nin * ROW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t tile_doubles = ((MODE == MODE_QUAD ? 1 : 2) * (size_t)nin * ROW + 1) & ~(size_t)1;
  char* ring = reinterpret_cast<char*>(smem + tile_doubles) + warp * kPairRingBytes;
  const int64_t nblocks = (nstates + ROW - 1) / ROW;
  for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    const int64_t s0 = blk * ROW;
    __syncthreads();  // everyone is done with the previous block's state tile
    for (int o = threadIdx.x; o < nin * ROW; o += kThreads) {
      const int j = o / ROW, s = o % ROW;
      double vx = 0.0, vy = 0.0;
      if (j == nin - 1) {
        vx = 1.0;
        vy = MODE == MODE_JVP ? 0.0 : 1.0;
      } else if (s0 + s < nstates) {
        vx = X[(s0 + s) + ldx * (int64_t)j];
        if (MODE != MODE_QUAD) vy = Y[(s0 + s) + ldx * (int64_t)j];
      }
      xs[j * ROW + (s ^ swz<MT>(j))] = vx;
      if (MODE != MODE_QUAD) ys[j * ROW + (s ^ swz<MT>(j))] = vy;
    }
    __syncthreads();
    const int b = ws.off[warp], mid = b + ws.nt2[warp], e = ws.off[warp + 1];
    run_my_tiles<MT, 2, MODE, FACT>(tiles, sched, b, mid, ws.ks2[warp], pairs, vals, rows, xs, ys, ring, lane, s0, nstates, OUT, ldo);
    run_my_tiles<MT, 1, MODE, FACT>(tiles, sched, mid, e, ws.ks1[warp], pairs, vals, rows, xs, ys, ring, lane, s0, nstates, OUT, ldo);
  }
}

template <int MT, int MODE>
inline size_t warptile_smem_bytes(int nin) {
  const size_t tile_doubles = ((MODE == MODE_QUAD ? 1 : 2) * (size_t)nin * 8 * MT + 1) & ~(size_t)1;
  return tile_doubles * sizeof(double) + (size_t)kWarps * kPairRingBytes;
}

// X, Y: nstates x (nin-1) column-major with leading dimension ldx; OUT: nstates x m, ldo.
template <int MT, int MODE>
__global__ void __launch_bounds__(kThreads, 2)
bilinear_batched_kernel(const TileDesc* __restrict__ tiles, int ntiles, int nt2_tiles,
                        const uint2* __restrict__ pairs, const double* __restrict__ vals,
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
    run_tiles<MT, 2, MODE>(tiles, 0, nt2_tiles, pairs, vals, rows, xs, ys, red, warp, lane, s0, nstates, OUT, ldo);
    run_tiles<MT, 1, MODE>(tiles, nt2_tiles, ntiles, pairs, vals, rows, xs, ys, red, warp, lane, s0, nstates, OUT, ldo);
  }
}

// ------------------------------------------------------------------------------------------------
// Windowed variant for large m: instead of the whole state tile, shared memory holds the rows of the
// modes one WINDOW of k-steps touches (<= wmax, see pack.h), gathered with cp.async into a double
// buffer while the previous window is being multiplied.  One barrier per window; accumulators live
// across the windows of a tile.  Footprint: 2 x wmax x 256 B per vector, independent of m.
//
// Everything a window needs arrives through the same cp.async pipeline, so no warp ever waits on a
// dependent L2 load: at the top of window g (right after its barrier) the CTA issues
//   the state rows and the pair words of window g+1   (their descriptor and mode list are already in smem),
//   the mode list of window g+2                       (its descriptor is already in smem),
//   the descriptor of window g+3,
// and the next barrier's cp.async.wait_all makes all of it visible.  Operator values stay in registers,
// prefetched kPrefetch k-steps ahead with loads that are pinned in program order (volatile asm: the
// compiler otherwise sinks them to the end of the unrolled body and exposes the L2 latency).
constexpr int kWinMT = 4;
constexpr int kDescRing = 4;

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc, int src_bytes) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16z(double* smem_dst, const double* gsrc, int src_bytes) {  // zero-fills beyond src_bytes
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

constexpr int kValSlotBytes = 2048;               // one batch of operator values: 4 k-steps (nt = 2) or 8 (nt = 1)
constexpr int kValRingBytes = 2 * kValSlotBytes;  // per warp

struct WinSmem {
  double* bufs;      // [2][bufstride] state rows of the current / next window; a finished tile's reduction reuses the current one
  uint32_t* pairs;   // [2][kWindowMaxKs * 4] packed pair words (first | second << 16) of the current / next window
  WindowDesc* desc;  // ring of kDescRing descriptors: windows g .. g+3
  int32_t* modes;    // [2][wpad] mode lists of windows g+1 / g+2
  char* vring;       // [NW][kValRingBytes] operator values, two batches per warp
  int wmax, wpad;
  size_t bufstride;  // doubles
};

// warps per CTA: two CTAs of 8 warps per SM for one-vector evaluation; the two-vector modes need twice the window
// buffers, so one CTA of 16 warps shares them
// NT3 (tiles of three n-tiles, CA_NT3): 24 accumulator doubles per thread need more than 128 registers, so that
// variant runs one CTA of 8 warps per SM in every mode.
template <int MODE, bool NT3 = false>
struct WinCfg {
  static constexpr int NW = (MODE == MODE_QUAD || NT3) ? 8 : 16;
  static constexpr int CTAS = (MODE == MODE_QUAD && !NT3) ? 2 : 1;
};

template <int MODE, bool NT3 = false>
__host__ __device__ inline size_t window_buf_doubles(int wmax) {
  const size_t row = 8 * kWinMT;
  const size_t states = (size_t)(MODE == MODE_QUAD ? 1 : 2) * wmax * row;
  const size_t red = (size_t)WinCfg<MODE, NT3>::NW * kNtMax * 8 * row;
  return states > red ? states : red;
}

template <int MODE, bool NT3 = false>
__host__ __device__ inline size_t windowed_smem_bytes(int wmax) {
  const size_t wpad = ((size_t)wmax + 3) & ~(size_t)3;
  return 2 * window_buf_doubles<MODE, NT3>(wmax) * sizeof(double) + (size_t)2 * kWindowMaxKs * 4 * sizeof(uint32_t) +
         kDescRing * sizeof(WindowDesc) + 2 * wpad * sizeof(int32_t) + (size_t)WinCfg<MODE, NT3>::NW * kValRingBytes;
}

template <int MODE, bool NT3 = false>
__device__ __forceinline__ WinSmem carve_window_smem(double* smem, int wmax) {
  WinSmem S;
  S.wmax = wmax;
  S.wpad = (wmax + 3) & ~3;
  S.bufstride = window_buf_doubles<MODE, NT3>(wmax);
  S.bufs = smem;
  S.pairs = reinterpret_cast<uint32_t*>(smem + 2 * S.bufstride);
  S.desc = reinterpret_cast<WindowDesc*>(S.pairs + (size_t)2 * kWindowMaxKs * 4);
  S.modes = reinterpret_cast<int32_t*>(S.desc + kDescRing);
  S.vring = reinterpret_cast<char*>(S.modes + 2 * S.wpad);
  return S;
}

// Issued by the whole CTA right after the barrier of window g (g == -1: block prologue).
template <int MODE, bool NT3 = false>
__device__ __forceinline__ void stage_ahead(int g, int gtotal, const WindowDesc* __restrict__ windows,
                                            const int32_t* __restrict__ modes, const uint32_t* __restrict__ pairs,
                                            const WinSmem& S, int one_mode, const double* __restrict__ X,
                                            const double* __restrict__ Y, int64_t s0, int64_t nstates, int64_t ldx) {
  constexpr int ROW = 8 * kWinMT;
  constexpr int NW = WinCfg<MODE, NT3>::NW, NTHR = NW * 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (g + 1 < gtotal) {
    const WindowDesc& W = S.desc[(g + 1) & (kDescRing - 1)];
    const int nks = W.nks, nmodes = W.nmodes;
    const char* psrc = reinterpret_cast<const char*>(pairs + (size_t)W.ks_begin * 4);
    char* pdst = reinterpret_cast<char*>(S.pairs + (size_t)((g + 1) & 1) * kWindowMaxKs * 4);
    for (int c = tid; c < nks; c += NTHR) cp_async16(pdst + c * 16, psrc + c * 16);  // one k-step = 4 packed pairs = 16 B
    const int32_t* md = S.modes + ((g + 1) & 1) * S.wpad;
    double* xb = S.bufs + (size_t)((g + 1) & 1) * S.bufstride;
    double* yb = xb + (size_t)S.wmax * ROW;
    // 16-byte copies when the ensemble allows it (even leading dimension, 16-byte aligned base): a warp instruction
    // then moves TWO state rows (32 states = 256 B each; lane = row half, pair of states) -- half the staging
    // instructions of the 8-byte form, which remains for odd leading dimensions
    const bool wide = ((ldx & 1) == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0) &&
                      (MODE == MODE_QUAD || (reinterpret_cast<uintptr_t>(Y) & 15) == 0);
    if (wide) {
      const int h2 = 2 * (lane & 15), sub = lane >> 4;
      const int64_t s = s0 + h2;
      const int nbytes = s + 1 < nstates ? 16 : (s < nstates ? 8 : 0);
      const double* xg = X + (nbytes ? s : 0);
      const double* yg = MODE == MODE_QUAD ? nullptr : Y + (nbytes ? s : 0);
#pragma unroll 4
      for (int l = 2 * warp + sub; l < nmodes; l += 2 * NW) {
        const int gm = md[l];
        const int slot = l * ROW + (h2 ^ swz<kWinMT>(l));  // the swizzle flips bits 2-3 of the state: pairs stay adjacent
        if (gm == one_mode) {
          xb[slot] = xb[slot + 1] = 1.0;
          if (MODE != MODE_QUAD) yb[slot] = yb[slot + 1] = MODE == MODE_JVP ? 0.0 : 1.0;
        } else {
          const int64_t off = ldx * (int64_t)gm;
          cp_async16z(xb + slot, xg + off, nbytes);
          if (MODE != MODE_QUAD) cp_async16z(yb + slot, yg + off, nbytes);
        }
      }
    } else {
      // one state row per warp instruction; the lane is the state
      const bool ok = s0 + lane < nstates;
      const double* xg = X + (ok ? s0 + lane : 0);
      const double* yg = MODE == MODE_QUAD ? nullptr : Y + (ok ? s0 + lane : 0);
#pragma unroll 4
      for (int l = warp; l < nmodes; l += NW) {
        const int gm = md[l];
        const int slot = l * ROW + (lane ^ swz<kWinMT>(l));
        if (gm == one_mode) {
          xb[slot] = 1.0;
          if (MODE != MODE_QUAD) yb[slot] = MODE == MODE_JVP ? 0.0 : 1.0;
        } else {
          const int64_t off = ldx * (int64_t)gm;
          cp_async8(xb + slot, xg + off, ok ? 8 : 0);
          if (MODE != MODE_QUAD) cp_async8(yb + slot, yg + off, ok ? 8 : 0);
        }
      }
    }
  }
  if (g + 2 < gtotal) {
    const WindowDesc& W2 = S.desc[(g + 2) & (kDescRing - 1)];
    const int nmodes = W2.nmodes;
    const int32_t* src = modes + W2.modes_off;
    int32_t* md = S.modes + ((g + 2) & 1) * S.wpad;
    for (int l = tid; l < nmodes; l += NTHR) cp_async4(md + l, src + l);
  }
  if (g + 3 < gtotal && tid < (int)(sizeof(WindowDesc) / 16))
    cp_async16(reinterpret_cast<char*>(S.desc + ((g + 3) & (kDescRing - 1))) + tid * 16,
               reinterpret_cast<const char*>(windows + g + 3) + tid * 16);
  cp_async_commit();
}

// Operator values of a warp's share of a window go through a two-slot ring in shared memory, one cp.async group per
// batch: ptxas keeps LDGSTS / commit / wait in program order, whereas plain register prefetch loads (having no
// consumer inside the unrolled body) get sunk to the end of the body, right in front of their use.
template <int NT, int NW>
struct ValRing {
  static constexpr int KB = kValSlotBytes / (NT * 256);  // k-steps per batch
  char* ring;       // this warp's kValRingBytes
  const char* src;  // the share's first value
  int n;            // k-steps in the share

  static constexpr int kBatchBytes = KB * NT * 256;                 // 2048 (NT = 1, 2) or 1536 (NT = 3)
  static constexpr int kCopies = (kBatchBytes + 511) / 512;         // 16-byte copies per lane and full batch
  uint32_t dst_lane;     // shared-memory address of this lane's first 16 bytes in slot 0
  const char* src_lane;  // this lane's first 16 bytes of the share

  __device__ __forceinline__ void issue(int b, int lane) const {
    const int left = n - b * KB;  // k-steps from batch b on; <= 0 past the end: an empty group keeps the count uniform
    const char* s = src_lane + (size_t)b * kBatchBytes;
    const uint32_t d = dst_lane + (b & 1) * kValSlotBytes;
    if (left >= KB) {  // (warp-uniform) a full batch: constant offsets, no per-copy predicate
#pragma unroll
      for (int c = 0; c < kCopies; ++c)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d + c * 512), "l"(s + c * 512) : "memory");
    } else if (left > 0) {
      const int bytes = left * (NT * 256) - lane * 16;
#pragma unroll
      for (int c = 0; c < kCopies; ++c)
        if (c * 512 < bytes) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d + c * 512), "l"(s + c * 512) : "memory");
    }
    cp_async_commit();
  }
  // batches 0 and 1 of this warp's share of a window; issued before the previous window's barrier
  __device__ __forceinline__ void begin(const double* __restrict__ vals0, int nks, int warp, int lane) {
    int k0, k1;
    warp_share<NW>(nks, warp, k0, k1);
    n = k1 - k0;
    src = reinterpret_cast<const char*>(vals0 + (size_t)k0 * NT * 32);
    src_lane = src + lane * 16;
    dst_lane = (uint32_t)__cvta_generic_to_shared(ring) + lane * 16;
    __syncwarp();  // every lane is done reading the previous share's slots
    issue(0, lane);
    issue(1, lane);
  }
};

// ps: the window's pair words in shared memory
template <int MT, int NT, int MODE, bool NT3 = false>
__device__ __forceinline__ void window_accumulate(const uint32_t* ps, int nks, const double* xs, const double* ys,
                                                  int warp, int lane, const ValRing<NT, WinCfg<MODE, NT3>::NW>& vr,
                                                  double (&acc)[MT][NT][2]) {
  constexpr int KB = ValRing<NT, WinCfg<MODE, NT3>::NW>::KB;
  const uint32_t r8 = 8u * (lane >> 2);
  int k0, k1;
  warp_share<WinCfg<MODE, NT3>::NW>(nks, warp, k0, k1);
  const int n = k1 - k0;
  if (n <= 0) return;
  const uint32_t* pp = ps + (size_t)k0 * 4 + (lane & 3);
  const char* vb = vr.ring + lane * 8;
  const char* xsb = reinterpret_cast<const char*>(xs);
  const char* ysb = reinterpret_cast<const char*>(ys);
  double xa[MT], va[MT];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) xa[mt] = va[mt] = 0.0;
  auto step = [&](int k, const char* slot, int kk) {
    const uint32_t packed = pp[(size_t)k * 4];  // window-local rows: both address words fit 16 bits (pack_window_pair)
    uint2 pw = make_uint2(packed & 0xffffu, packed >> 16);
    if (k == 0) pw.x &= ~kSameA;  // the first k-step of a share always loads its first factors
    double bv[NT];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) bv[nt] = *reinterpret_cast<const double*>(slot + (kk * NT + nt) * 256);
    kstep<MT, NT, MODE>(pw, bv, xsb, ysb, r8, xa, va, acc);
  };
  const int nb = (n + KB - 1) / KB;
  for (int b = 0; b < nb; ++b) {
    if (b >= 2) {  // batches 0 and 1 landed before the window's barrier
      asm volatile("cp.async.wait_group 1;\n" ::: "memory");
      __syncwarp();
    }
    const char* slot = vb + (b & 1) * kValSlotBytes;
    const int cnt = min(KB, n - b * KB);
    if (cnt == KB) {
#pragma unroll
      for (int kk = 0; kk < KB; ++kk) step(b * KB + kk, slot, kk);
    } else {
      for (int kk = 0; kk < cnt; ++kk) step(b * KB + kk, slot, kk);
    }
    __syncwarp();  // the slot is refilled now
    vr.issue(b + 2, lane);
  }
}

// windows [gb, ge) all belong to tiles with NT n-tiles; gtotal = number of windows of the operator
template <int NT, int MODE, bool NT3 = false>
__device__ __forceinline__ void run_windows(const WindowDesc* __restrict__ windows, int gb, int ge, int gtotal,
                                            const int32_t* __restrict__ modes, const uint32_t* __restrict__ pairs,
                                            const double* __restrict__ vals, const int32_t* __restrict__ rows,
                                            const WinSmem& S, int one_mode, int warp, int lane,
                                            const double* __restrict__ X, const double* __restrict__ Y, int64_t s0,
                                            int64_t nstates, int64_t ldx, double* __restrict__ OUT, int64_t ldo) {
  constexpr int MT = kWinMT, ROW = 8 * MT;
  if (gb >= ge) return;
  ValRing<NT, WinCfg<MODE, NT3>::NW> vr;
  vr.ring = S.vring + warp * kValRingBytes;
  double acc[MT][NT][2];
  {
    const WindowDesc& W0 = S.desc[gb & (kDescRing - 1)];
    vr.begin(vals + W0.vals_off, W0.nks, warp, lane);
  }
  bool fresh = true;
  for (int g = gb; g < ge; ++g) {
    cp_async_wait_all();
    __syncthreads();  // window g (+ mode list g+1, descriptor g+2) has landed; window g-1's buffers are free
    stage_ahead<MODE, NT3>(g, gtotal, windows, modes, pairs, S, one_mode, X, Y, s0, nstates, ldx);
    if (fresh) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
    }
    const WindowDesc& W = S.desc[g & (kDescRing - 1)];
    const int nks = W.nks, rows_off = W.rows_off;
    const bool last = W.last != 0;
    double* cur = S.bufs + (size_t)(g & 1) * S.bufstride;
    const double* xs = cur;
    const double* ys = MODE == MODE_QUAD ? xs : xs + (size_t)S.wmax * ROW;
    window_accumulate<MT, NT, MODE, NT3>(S.pairs + (size_t)(g & 1) * kWindowMaxKs * 4, nks, xs, ys, warp, lane, vr, acc);
    if (g + 1 < ge) {
      const WindowDesc& Wn = S.desc[(g + 1) & (kDescRing - 1)];
      vr.begin(vals + Wn.vals_off, Wn.nks, warp, lane);
    }
    // the reduction buffer is the finished window's state buffer: its first barrier frees it, and the next
    // window's barrier comes before anything is staged into it again
    if (last) tile_reduce_store<MT, NT, WinCfg<MODE, NT3>::NW>(rows_off, rows, cur, acc, warp, lane, s0, nstates, OUT, ldo);
    fresh = last;
  }
}

// NT3: the pack has tiles of three n-tiles (CA_NT3); a separate instantiation, so that the register allocation of the
// two-n-tile kernel does not change with it
template <int MODE, bool NT3>
__global__ void __launch_bounds__(WinCfg<MODE, NT3>::NW * 32, WinCfg<MODE, NT3>::CTAS)
bilinear_windowed_kernel(const WindowDesc* __restrict__ windows, int nwin, int nwin_nt3, int nwin_nt2,
                         const int32_t* __restrict__ modes, const uint32_t* __restrict__ pairs,
                         const double* __restrict__ vals, const int32_t* __restrict__ rows, int wmax, int one_mode,
                         const double* __restrict__ X, const double* __restrict__ Y, int64_t nstates, int64_t ldx,
                         double* __restrict__ OUT, int64_t ldo) {
  constexpr int ROW = 8 * kWinMT;
  extern __shared__ __align__(16) double smem[];
  const WinSmem S = carve_window_smem<MODE, NT3>(smem, wmax);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t nblocks = (nstates + ROW - 1) / ROW;
  if (nwin <= 0) return;
  for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    const int64_t s0 = blk * ROW;
    __syncthreads();  // the previous block's last window and reduction are finished
    constexpr int kDescWords = sizeof(WindowDesc) / 4;
    if (tid < 2 * kDescWords && tid / kDescWords < nwin)  // descriptors 0 and 1 directly
      reinterpret_cast<uint32_t*>(S.desc)[tid] = __ldg(reinterpret_cast<const uint32_t*>(windows) + tid);
    __syncthreads();
    for (int l = tid; l < S.desc[0].nmodes; l += WinCfg<MODE, NT3>::NW * 32) S.modes[l] = __ldg(modes + S.desc[0].modes_off + l);
    __syncthreads();
    stage_ahead<MODE, NT3>(-1, nwin, windows, modes, pairs, S, one_mode, X, Y, s0, nstates, ldx);
    if constexpr (NT3)
      run_windows<3, MODE, NT3>(windows, 0, nwin_nt3, nwin, modes, pairs, vals, rows, S, one_mode, warp, lane, X, Y, s0,
                           nstates, ldx, OUT, ldo);
    run_windows<2, MODE, NT3>(windows, nwin_nt3, nwin_nt3 + nwin_nt2, nwin, modes, pairs, vals, rows, S, one_mode, warp, lane,
                         X, Y, s0, nstates, ldx, OUT, ldo);
    run_windows<1, MODE, NT3>(windows, nwin_nt3 + nwin_nt2, nwin, nwin, modes, pairs, vals, rows, S, one_mode, warp, lane, X, Y,
                         s0, nstates, ldx, OUT, ldo);
  }
}

template <int MT, int MODE>
inline size_t batched_smem_bytes(int nin) {
  const size_t row = 8 * MT;
  return ((MODE == MODE_QUAD ? 1 : 2) * (size_t)nin * row + (size_t)kWarps * kNtMax * 8 * row) *
         sizeof(double);
}

}  // namespace ca
