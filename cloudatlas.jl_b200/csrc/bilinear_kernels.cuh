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

// barrier `id` (1 ..) among NTHREADS threads of the CTA (the consumer warps of one team of the windowed kernel)
template <int NTHREADS>
__device__ __forceinline__ void named_barrier(int id) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "n"(NTHREADS) : "memory");
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

// Ranges [first[p], first[p+1]) of windows (windowed kernel) or tiles (k-split kernel) for the CTAs that share a block
// of states when the ensemble has fewer blocks than the machine has CTA slots; n == 1: one CTA walks everything
struct WindowParts {
  static constexpr int kMax = 32;
  int n;
  int first[kMax + 1];
};
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
template <int MT, int NTA, int NT0, int NTC, int NW, bool WS = false>
__device__ __forceinline__ void tile_reduce_store_part(int rows_off, const int32_t* __restrict__ rows,
                                                       double* red, const double (&acc)[MT][NTA][2],
                                                       int warp, int lane, int64_t s0, int64_t nstates,
                                                       double* __restrict__ OUT, int64_t ldo, int bar_id = 1) {
  static_assert(NTC <= kNtMax && NT0 + NTC <= NTA, "slice of the accumulator");
  constexpr int ROW = 8 * MT;
  const int r = lane >> 2, c = lane & 3;
  auto barrier = [bar_id] {  // WS: `warp` counts within the NW warps that share the reduction
    if constexpr (WS) named_barrier<NW * 32>(bar_id);
    else __syncthreads();
  };
  barrier();  // previous tile's readers are done with `red`
  double* mine = red + warp * (kNtMax * 8 * ROW);
#pragma unroll
  for (int nt = 0; nt < NTC; ++nt)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      mine[(nt * 8 + c * 2 + 0) * ROW + mt * 8 + r] = acc[mt][NT0 + nt][0];
      mine[(nt * 8 + c * 2 + 1) * ROW + mt * 8 + r] = acc[mt][NT0 + nt][1];
    }
  barrier();
  for (int o = warp * 32 + lane; o < NTC * 8 * ROW; o += NW * 32) {
    const int n = o / ROW, s = o % ROW;
    const int row = __ldg(rows + rows_off + NT0 * 8 + n);
    double sum = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) sum += red[w * (kNtMax * 8 * ROW) + o];
    if (row >= 0 && s0 + s < nstates) OUT[(s0 + s) + ldo * (int64_t)row] = sum;
  }
}

template <int MT, int NT, int NW = kWarps, bool WS = false>
__device__ __forceinline__ void tile_reduce_store(int rows_off, const int32_t* __restrict__ rows,
                                                  double* red, const double (&acc)[MT][NT][2],
                                                  int warp, int lane, int64_t s0, int64_t nstates,
                                                  double* __restrict__ OUT, int64_t ldo, int bar_id = 1) {
  if constexpr (NT <= kNtMax) {
    tile_reduce_store_part<MT, NT, 0, NT, NW, WS>(rows_off, rows, red, acc, warp, lane, s0, nstates, OUT, ldo, bar_id);
  } else {  // three n-tiles: two passes through the two-tile reduction buffer
    tile_reduce_store_part<MT, NT, 0, kNtMax, NW, WS>(rows_off, rows, red, acc, warp, lane, s0, nstates, OUT, ldo, bar_id);
    tile_reduce_store_part<MT, NT, kNtMax, NT - kNtMax, NW, WS>(rows_off, rows, red, acc, warp, lane, s0, nstates, OUT, ldo, bar_id);
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
static_assert(kPairBatch == kStreamAlign, "tiles start on batch boundaries of the pair stream (pack.cpp)");
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
  // g is on a batch boundary (tiles start on one, pack.cpp): refills one batch ahead and makes the batch of k-steps
  // g .. g + kPairBatch - 1 readable
  __device__ __forceinline__ void batch(int lane) const {
    __syncwarp();                       // the batch slot refilled now was consumed by every lane
    issue_batch(g / kPairBatch + kPairBatches - 1, lane);
    asm volatile("cp.async.wait_group %0;\n" ::"n"(kPairBatches - 1) : "memory");
    __syncwarp();                       // batch g / kPairBatch has landed for all lanes
  }
  // pair word of k-step g + u (u < kPairBatch) for this lane's column
  __device__ __forceinline__ uint2 word(int u, int lane) const {
    return reinterpret_cast<const uint2*>(ring)[((g & (kPairBatch * kPairBatches - 1)) + u) * 4 + (lane & 3)];
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
  // whole batches of the pair ring: one refill, then kPairBatch k-steps without a branch -- ptxas moves the shared-memory
  // loads of the following k-steps up between the DMMAs of the current one (with a refill check in front of every
  // k-step each of them was a basic block of its own and exposed its LDS -> DMUL -> DMMA chain)
  static_assert(kPairBatch % kPrefetch == 0, "the value registers rotate with period kPrefetch");
  int i = 0;
  for (; i + kPairBatch <= nks; i += kPairBatch) {
    pr.batch(lane);
#pragma unroll
    for (int u = 0; u < kPairBatch; ++u) {
      uint2 pw = pr.word(u, lane);
      if (i + u == 0) pw.x &= ~kSameA;  // first k-step of a tile loads its first factors
      one(pw, pf.bv[u % kPrefetch]);
      prefetch_vals<NT>(pf.bv[u % kPrefetch], vp, min(i + u + kPrefetch, nks - 1));
    }
    pr.g += kPairBatch;
  }
  if (i < nks) {  // the tile's last, partial batch; the stream continues on the next batch boundary
    pr.batch(lane);
#pragma unroll
    for (int u = 0; u < kPairBatch - 1; ++u) {
      if (i + u < nks) {
        uint2 pw = pr.word(u, lane);
        if (i + u == 0) pw.x &= ~kSameA;
        one(pw, pf.bv[u % kPrefetch]);
        if (u + kPrefetch < kPairBatch - 1) prefetch_vals<NT>(pf.bv[u % kPrefetch], vp, min(i + u + kPrefetch, nks - 1));
      }
    }
    pr.g += kPairBatch;
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
    // Two-n-tile k-steps saturate the FP64 pipe (8 DMMA + 4 DMUL), one-n-tile k-steps (4 + 4) keep it ~70 % busy: half
    // the warps of every scheduler (warps 4-7) start with their one-n-tile tiles, so the two kinds overlap (-1.1 %)
    if (warp & 4) {
      run_my_tiles<MT, 1, MODE, FACT>(tiles, sched, mid, e, ws.ks1[warp], pairs, vals, rows, xs, ys, ring, lane, s0, nstates, OUT, ldo);
      run_my_tiles<MT, 2, MODE, FACT>(tiles, sched, b, mid, ws.ks2[warp], pairs, vals, rows, xs, ys, ring, lane, s0, nstates, OUT, ldo);
      continue;
    }
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
bilinear_batched_kernel(const TileDesc* __restrict__ tiles, const WindowParts parts, int nt2_tiles,
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
  const int part = blockIdx.x % parts.n;  // this CTA's range of tiles
  const int t0 = parts.first[part], t1 = parts.first[part + 1];
  if (t1 <= t0) return;

  for (int64_t blk = blockIdx.x / parts.n; blk < nblocks; blk += gridDim.x / parts.n) {
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
    run_tiles<MT, 2, MODE>(tiles, t0, min(t1, nt2_tiles), pairs, vals, rows, xs, ys, red, warp, lane, s0, nstates, OUT, ldo);
    run_tiles<MT, 1, MODE>(tiles, max(t0, nt2_tiles), t1, pairs, vals, rows, xs, ys, red, warp, lane, s0, nstates, OUT, ldo);
  }
}

// ------------------------------------------------------------------------------------------------
// Windowed variant for large m: instead of the whole state tile, shared memory holds the rows of the
// modes one WINDOW of k-steps touches (<= wmax, see pack.h), gathered with cp.async into a double
// buffer while the previous window is being multiplied.  Accumulators live across the windows of a
// tile.  Footprint: 2 x wmax x 256 B per vector, independent of m.
//
// The CTA is split by role:
//   producer warps  gather the state rows and the pair words of window g+1 into the free slot -- every byte through
//                   cp.async, completion signalled by cp.async.mbarrier.arrive on the slot's `full` barrier;
//   consumer warps  wait on `full`, multiply their share of the window's k-steps (operator values through a per-warp
//                   cp.async ring straight from L2), and release the slot on its `empty` barrier.
// (WinCfg: warps per role, CTAs per SM and slots per evaluation mode.)
// No CTA-wide barrier per window: consumer warps drift apart by up to a window, so one warp's window turn-over
// (barrier wait, value-ring restart, ragged last batch) hides behind the other warps' DMMAs instead of idling the
// FP64 pipe for the whole CTA -- in the one-role version of round 1 the k-step bodies ran at pipe speed but made up
// only 58 % of the time.  Consumers meet (named barrier) only at the end of a tile, to sum their partial accumulators
// in a fixed order.
constexpr int kWinMT = 4;

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

// ---- mbarriers (shared memory, CTA scope)
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
// the arrival happens when all cp.async operations this thread has issued so far have landed
__device__ __forceinline__ void mbar_arrive_on_copies(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!done);
}

constexpr int kValSlotBytes = 2048;               // one batch of operator values: 4 k-steps (nt = 2) or 8 (nt = 1)
constexpr int kValRingBytes = 2 * kValSlotBytes;  // per warp

// the constant rows of the pseudo mode (index nin - 1: 1 for x, 1 or 0 for the second vector), staged like any other
// row so that every byte of a window arrives through cp.async and one arrive-on-copies covers the slot
__device__ __align__(16) const double kRowOfOnes[32] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                          1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
__device__ __align__(16) const double kRowOfZeros[32] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                           0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

struct WinSmem {
  double* bufs;      // [SLOTS][bufstride] state rows of the window slots
  uint32_t* pairs;   // [SLOTS][kWindowMaxKs * 4] packed pair words (first | second << 16)
  double* red;       // reduction buffer of a finished tile (SEPRED), else nullptr: the tile's last slot is reused
  char* vring;       // [NWC][kValRingBytes] operator values, two batches per consumer warp
  uint64_t* full;    // [SLOTS] slot filled   (producer threads arrive when their copies have landed)
  uint64_t* empty;   // [SLOTS] slot released (one arrival per consumer warp)
  int wmax;
  size_t bufstride;  // doubles
};

// One CTA per SM holds TEAMS independent teams of NWC consumer + NWP producer warps; a team works on its own block of
// 32 states with its own window slots and barriers.  What fits into 227 KB of shared memory and 64 K registers:
//   one vector  (f):                       two teams of 6 + 2 warps, two window slots of <= 32 KB each -- the teams drift
//                                          apart, so one's tile reductions and window turn-overs overlap the other's DMMAs
//   two vectors (Jacobian action, N(x,y)): one team of 12 + 4 warps, two slots of <= 64 KB
//   packs with three-n-tile tiles (CA_NT3 / the packer's choice at m ~ 3000) carry 24 accumulator doubles per thread:
//                                          8 + 4 warps at <= 168 registers, as two teams of 4 + 2 for one vector
// Warp order: all consumers (team by team), then all producers, so that every scheduler (warp % 4) gets the same
// number of consumer warps -- two CTAs of 6 + 2 warps per SM instead leave two schedulers with half the DMMA work of
// the other two (measured: 14.1 against 11.4 ms at 999 modes).
// SLOTS: window slots in flight.  SEPRED: the tile reduction has a buffer of its own and the slot is released before
// the reduction; otherwise the reduction reuses the tile's last slot.
template <int MODE, bool NT3 = false>
struct WinCfg {
#ifdef CA_WIN_ONE_TEAM  // tuning hook (tools/build_variant.sh)
  static constexpr int TEAMS = 1;
#else
  static constexpr int TEAMS = MODE == MODE_QUAD ? 2 : 1;
#endif
  static constexpr int NWC = (NT3 ? 8 : 12) / TEAMS;
  static constexpr int NWP = 4 / TEAMS;
  static constexpr int THREADS = TEAMS * (NWC + NWP) * 32;
  static constexpr int SLOTS = (MODE == MODE_QUAD && TEAMS == 1) ? 3 : 2;
  static constexpr bool SEPRED = TEAMS == 1 && (MODE == MODE_QUAD || NT3);
};

struct RingPos {  // position in the ring of window slots: slot index and the parity of its current use
  int slot = 0;
  uint32_t phase = 0;
  template <int SLOTS>
  __device__ __forceinline__ void next() {
    if (++slot == SLOTS) {
      slot = 0;
      phase ^= 1u;
    }
  }
};

template <int MODE, bool NT3 = false>
__host__ __device__ inline size_t window_buf_doubles(int wmax) {
  const size_t row = 8 * kWinMT;
  const size_t states = (size_t)(MODE == MODE_QUAD ? 1 : 2) * wmax * row;
  const size_t red = (size_t)WinCfg<MODE, NT3>::NWC * kNtMax * 8 * row;
  return WinCfg<MODE, NT3>::SEPRED || states > red ? states : red;
}

template <int MODE, bool NT3 = false>
__host__ __device__ inline size_t windowed_team_bytes(int wmax) {  // a multiple of 16
  using Cfg = WinCfg<MODE, NT3>;
  return Cfg::SLOTS * (window_buf_doubles<MODE, NT3>(wmax) * sizeof(double) + (size_t)kWindowMaxKs * 4 * sizeof(uint32_t)) +
         (Cfg::SEPRED ? (size_t)Cfg::NWC * kNtMax * 8 * 8 * kWinMT * sizeof(double) : 0) + (size_t)Cfg::NWC * kValRingBytes +
         2 * Cfg::SLOTS * sizeof(uint64_t);
}

template <int MODE, bool NT3 = false>
__host__ __device__ inline size_t windowed_smem_bytes(int wmax) {
  return WinCfg<MODE, NT3>::TEAMS * windowed_team_bytes<MODE, NT3>(wmax);
}

template <int MODE, bool NT3 = false>
__device__ __forceinline__ WinSmem carve_window_smem(double* smem, int wmax) {
  WinSmem S;
  S.wmax = wmax;
  S.bufstride = window_buf_doubles<MODE, NT3>(wmax);
  S.bufs = smem;
  using Cfg = WinCfg<MODE, NT3>;
  double* after = smem + Cfg::SLOTS * S.bufstride;
  S.red = Cfg::SEPRED ? after : nullptr;
  if (Cfg::SEPRED) after += (size_t)Cfg::NWC * kNtMax * 8 * 8 * kWinMT;
  S.pairs = reinterpret_cast<uint32_t*>(after);
  S.vring = reinterpret_cast<char*>(S.pairs + (size_t)Cfg::SLOTS * kWindowMaxKs * 4);
  S.full = reinterpret_cast<uint64_t*>(S.vring + (size_t)Cfg::NWC * kValRingBytes);
  S.empty = S.full + Cfg::SLOTS;
  return S;
}

// ---- producer warps: stage every window of every state block this CTA owns, in consumption order
template <int MODE, bool NT3>
__device__ __forceinline__ void window_producer(const WindowDesc* __restrict__ windows, int w0, int w1, int nparts,
                                                const int32_t* __restrict__ modes, const uint32_t* __restrict__ pairs,
                                                const WinSmem& S, int one_mode, const double* __restrict__ X,
                                                const double* __restrict__ Y, int64_t nstates, int64_t ldx, int team,
                                                int ptid) {  // ptid: index among the team's producer threads
  constexpr int ROW = 8 * kWinMT;
  constexpr int NWP = WinCfg<MODE, NT3>::NWP, NPT = NWP * 32, TEAMS = WinCfg<MODE, NT3>::TEAMS;
  const int pwarp = ptid >> 5, lane = ptid & 31;
  const int64_t nblocks = (nstates + ROW - 1) / ROW;
  // 16-byte copies when the ensemble allows it (even leading dimension, 16-byte aligned base): a warp instruction
  // then moves TWO state rows (32 states = 256 B each; lane = row half, pair of states) -- half the staging
  // instructions of the 8-byte form, which remains for odd leading dimensions
  const bool wide = ((ldx & 1) == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0) &&
                    (MODE == MODE_QUAD || (reinterpret_cast<uintptr_t>(Y) & 15) == 0);
  const int h2 = 2 * (lane & 15), sub = lane >> 4;
  const double* const one_y = MODE == MODE_JVP ? kRowOfZeros : kRowOfOnes;
  RingPos pos;
  for (int64_t blk = (int64_t)(blockIdx.x / nparts) * TEAMS + team; blk < nblocks; blk += (int64_t)(gridDim.x / nparts) * TEAMS) {
    const int64_t s0 = blk * ROW;
    // this lane's piece of a state row: byte count inside the ensemble (0: beyond the last state -> zero fill)
    const int64_t sw = s0 + h2;
    const int nbytes_w = sw + 1 < nstates ? 16 : (sw < nstates ? 8 : 0);
    const int nbytes_n = s0 + lane < nstates ? 8 : 0;
    const int64_t soff = wide ? (nbytes_w ? sw : 0) : (nbytes_n ? s0 + lane : 0);
    const int coff = wide ? h2 : lane;  // the same piece of a constant row
    for (int g = w0; g < w1; ++g, pos.next<WinCfg<MODE, NT3>::SLOTS>()) {
      const WindowDesc* W = windows + g;
      const int nks = __ldg(&W->nks), nmodes = __ldg(&W->nmodes);
      const int32_t* mp = modes + __ldg(&W->modes_off);
      const char* psrc = reinterpret_cast<const char*>(pairs + (size_t)__ldg(&W->ks_begin) * 4);
      int mch[8];  // the window's mode list, 32 per register (wmax <= 256: pair words hold 8-bit local rows)
#pragma unroll
      for (int j = 0; j < 8; ++j) mch[j] = j * 32 + lane < nmodes ? __ldg(mp + j * 32 + lane) : 0;
      const int slot = pos.slot;
      mbar_wait(&S.empty[slot], pos.phase ^ 1u);  // the consumers are done with the slot's previous window
      char* pdst = reinterpret_cast<char*>(S.pairs + (size_t)slot * kWindowMaxKs * 4);
      for (int c = ptid; c < nks; c += NPT) cp_async16(pdst + c * 16, psrc + c * 16);  // one k-step = 4 packed pairs = 16 B
      double* xb = S.bufs + (size_t)slot * S.bufstride;
      double* yb = xb + (size_t)S.wmax * ROW;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j * 32 >= nmodes) break;
        if (wide) {
#pragma unroll
          for (int r = 2 * pwarp + sub; r < 32; r += 2 * NWP) {
            const int gm = __shfl_sync(0xffffffffu, mch[j], r);
            const int l = j * 32 + r;
            if (l < nmodes) {
              const int slotw = l * ROW + (h2 ^ swz<kWinMT>(l));  // the swizzle flips bits 2-3 of the state: pairs stay adjacent
              const bool one = gm == one_mode;
              const int64_t off = ldx * (int64_t)gm + soff;
              cp_async16z(xb + slotw, one ? kRowOfOnes + coff : X + off, one ? 16 : nbytes_w);
              if (MODE != MODE_QUAD) cp_async16z(yb + slotw, one ? one_y + coff : Y + off, one ? 16 : nbytes_w);
            }
          }
        } else {
#pragma unroll
          for (int r = pwarp; r < 32; r += NWP) {
            const int gm = __shfl_sync(0xffffffffu, mch[j], r);
            const int l = j * 32 + r;
            if (l < nmodes) {
              const int slotn = l * ROW + (lane ^ swz<kWinMT>(l));
              const bool one = gm == one_mode;
              const int64_t off = ldx * (int64_t)gm + soff;
              cp_async8(xb + slotn, one ? kRowOfOnes + coff : X + off, one ? 8 : nbytes_n);
              if (MODE != MODE_QUAD) cp_async8(yb + slotn, one ? one_y + coff : Y + off, one ? 8 : nbytes_n);
            }
          }
        }
      }
      mbar_arrive_on_copies(&S.full[slot]);
    }
  }
  cp_async_wait_all();
}

// Operator values of a warp's shares go through a two-slot ring in shared memory, one cp.async group per batch of KB
// k-steps: ptxas keeps LDGSTS / commit / wait in program order, whereas plain register prefetch loads (having no
// consumer inside the unrolled body) get sunk to the end of the body, right in front of their use.
// The ring runs ACROSS windows: while the last two batches of a share are multiplied, the first two of the warp's share
// of the next window are already on their way, so a window turn-over does not expose the L2 / HBM latency.
// Group accounting: exactly one group is committed per multiplied batch (real or empty), and `advance` tops the new
// share up to two; hence at the top of batch b the two newest groups are (batch b, batch b + 1) and
// cp.async.wait_group 1 means "batch b has landed".
template <int NT>
struct ValRing {
  static constexpr int KB = kValSlotBytes / (NT * 256);      // k-steps per batch
  static constexpr int kBatchBytes = KB * NT * 256;          // 2048 (NT = 1, 2) or 1536 (NT = 3)
  static constexpr int kCopies = (kBatchBytes + 511) / 512;  // 16-byte copies per lane and full batch
  const char* ring;     // this warp's kValRingBytes
  uint32_t dst_lane;    // shared-memory address of this lane's first 16 bytes in slot 0
  uint32_t par;         // slot of batch 0 of the current share
  const char* cur_src;  // this lane's first 16 bytes of the current share
  int cur_n;            // its k-steps
  const char* nxt_src;  // the warp's share of the next window (nxt_n < 0: not known)
  int nxt_n, nxt_issued;

  __device__ __forceinline__ void init(char* ring_, int lane) {
    ring = ring_;
    dst_lane = (uint32_t)__cvta_generic_to_shared(ring_) + lane * 16;
    par = 0;
    cur_src = nxt_src = nullptr;
    cur_n = 0;
    nxt_n = -1;
    nxt_issued = 0;
  }
  __device__ __forceinline__ void copy_batch(const char* src_lane, int n, int b, uint32_t slot, int lane) const {
    const int left = n - b * KB;  // k-steps from batch b on (> 0)
    const char* s = src_lane + (size_t)b * kBatchBytes;
    const uint32_t d = dst_lane + (slot & 1) * kValSlotBytes;
    if (left >= KB) {  // (warp-uniform) a full batch: constant offsets, no per-copy predicate
#pragma unroll
      for (int c = 0; c < kCopies; ++c)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d + c * 512), "l"(s + c * 512) : "memory");
    } else {
      const int bytes = left * (NT * 256) - lane * 16;
#pragma unroll
      for (int c = 0; c < kCopies; ++c)
        if (c * 512 < bytes) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d + c * 512), "l"(s + c * 512) : "memory");
    }
  }
  // share [k0, k1) of the window whose values start at vals0 comes after the current one
  __device__ __forceinline__ void set_next(const double* __restrict__ vals0, int k0, int k1, int lane) {
    nxt_src = reinterpret_cast<const char*>(vals0 + (size_t)k0 * NT * 32) + lane * 16;
    nxt_n = k1 - k0;
    nxt_issued = 0;
  }
  // the next share becomes the current one; afterwards its batches 0 and 1 are the two newest groups
  __device__ __forceinline__ void advance(int lane) {
    par = (par + (uint32_t)((cur_n + KB - 1) / KB)) & 1u;
    cur_src = nxt_src;
    cur_n = nxt_n > 0 ? nxt_n : 0;
    const int have = nxt_issued;
    nxt_n = -1;
    nxt_issued = 0;
    __syncwarp();  // every lane is done reading the finished share's slots
    for (int j = have; j < 2; ++j) {
      if (j * KB < cur_n) copy_batch(cur_src, cur_n, j, par + j, lane);
      cp_async_commit();
    }
  }
  // after batch b of the current share has been multiplied: batch b + 2, or else the next share's next batch
  __device__ __forceinline__ void issue_after(int b, int lane) {
    const int q = b + 2;
    if (q * KB < cur_n) {
      copy_batch(cur_src, cur_n, q, par + q, lane);
#ifdef CA_WIN_RING_CONTINUOUS  // tuning hook (measured 2 % slower than restarting the ring at every window)
    } else if (nxt_issued < 2 && nxt_issued * KB < nxt_n) {
      copy_batch(nxt_src, nxt_n, nxt_issued, par + (uint32_t)((cur_n + KB - 1) / KB) + nxt_issued, lane);
      ++nxt_issued;
#endif
    }
    cp_async_commit();
  }
};

// The share of consumer warp `warp` (of NW) of window g (nks k-steps): whole batches of KB k-steps, so that only the
// last share can end on a ragged batch (the slow, not unrolled path); which warp takes which part rotates with g.
// (Tried and dropped: uneven shares in the first / last window of a tile, to spread the warps' window turn-overs in
// time -- no gain once the one-vector mode runs two teams, and a loss combined with the rotation; one team of 14 + 2
// warps for the two-vector modes with shares weighted 3 : 4 to balance the schedulers -- 6 % slower than 12 + 4; the
// second team walking the one-n-tile windows first, as the whole-tile kernel's warps 4-7 do -- 4-7 % slower: the two
// teams then stream different parts of the operator and no longer share it in L2.)
template <int KB, int NW>
__device__ __forceinline__ void batch_share(int nks, int warp, int g, int& k0, int& k1) {
  const int nb = (nks + KB - 1) / KB;
  const int share = (warp + g) % NW;
  k0 = min(nks, (nb * share / NW) * KB);
  k1 = min(nks, (nb * (share + 1) / NW) * KB);
}

// ps: the window's pair words in shared memory; [k0, k1): this warp's k-steps (the current share of vr)
template <int MT, int NT, int MODE>
__device__ __forceinline__ void window_accumulate(const uint32_t* ps, int k0, int k1, const double* xs, const double* ys,
                                                  int lane, ValRing<NT>& vr, double (&acc)[MT][NT][2]) {
  constexpr int KB = ValRing<NT>::KB;
  const uint32_t r8 = 8u * (lane >> 2);
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
    asm volatile("cp.async.wait_group 1;\n" ::: "memory");  // batch b has landed (see ValRing)
    __syncwarp();
    const char* slot = vb + ((vr.par + b) & 1) * kValSlotBytes;
    const int cnt = min(KB, n - b * KB);
    if (cnt == KB) {
#pragma unroll
      for (int kk = 0; kk < KB; ++kk) step(b * KB + kk, slot, kk);
    } else {
      for (int kk = 0; kk < cnt; ++kk) step(b * KB + kk, slot, kk);
    }
    __syncwarp();  // the slot is refilled now
    vr.issue_after(b, lane);
  }
}

// Consumer side of windows [gb, ge), all of tiles with NT n-tiles.  // (`pos`: the ring position, advanced in step with the producers').  Window descriptors are read from global
// memory two windows ahead, one 4-byte word per lane, and broadcast by shuffles when their turn comes.
template <int NT, int MODE, bool NT3>
__device__ __forceinline__ void run_windows(const WindowDesc* __restrict__ windows, int gb, int ge,
                                            const double* __restrict__ vals, const int32_t* __restrict__ rows,
                                            const WinSmem& S, RingPos& pos, int warp, int lane, int64_t s0,
                                            int64_t nstates, double* __restrict__ OUT, int64_t ldo, int bar_id) {
  constexpr int MT = kWinMT, ROW = 8 * MT, NWC = WinCfg<MODE, NT3>::NWC, KB = ValRing<NT>::KB;
  constexpr int kWords = sizeof(WindowDesc) / 4;
  if (gb >= ge) return;
  auto desc_word = [&](int g) -> int {
    return (g < ge && lane < kWords) ? __ldg(reinterpret_cast<const int*>(windows + g) + lane) : 0;
  };
  auto field = [&](int dw, int word) -> int { return __shfl_sync(0xffffffffu, dw, word); };
  ValRing<NT> vr;
  vr.init(S.vring + warp * kValRingBytes, lane);
  double acc[MT][NT][2];
  int dw_n = desc_word(gb);  // descriptor of the window after the current one (here: of the first)
  int k0 = 0, k1 = 0, last = 0, rows_off = 0;
  // window g, described by dw_n: this warp's share of it becomes vr's next share
  auto take_next = [&](int g) {
    const int nks = field(dw_n, 1);
    const int64_t vals_off = (int64_t)(uint32_t)field(dw_n, 8) | ((int64_t)field(dw_n, 9) << 32);
    last = field(dw_n, 5);
    rows_off = field(dw_n, 6);
    batch_share<KB, NWC>(nks, warp, g, k0, k1);
    vr.set_next(vals + vals_off, k0, k1, lane);
  };
  take_next(gb);
  vr.advance(lane);
  dw_n = desc_word(gb + 1);
  bool fresh = true;
  for (int g = gb; g < ge; ++g, pos.next<WinCfg<MODE, NT3>::SLOTS>()) {
    const int dw_nn = desc_word(g + 2);  // in flight during this window
    const int k0c = k0, k1c = k1, rows_off_c = rows_off;
    const bool last_c = last != 0;
    const int slot = pos.slot;
    mbar_wait(&S.full[slot], pos.phase);
    if (fresh) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
    }
    double* cur = S.bufs + (size_t)slot * S.bufstride;
    const double* xs = cur;
    const double* ys = MODE == MODE_QUAD ? xs : xs + (size_t)S.wmax * ROW;
    window_accumulate<MT, NT, MODE>(S.pairs + (size_t)slot * kWindowMaxKs * 4, k0c, k1c, xs, ys, lane, vr, acc);
    // only now: the next window's share in registers across the k-steps costs the Jacobian action (at its 128
    // registers) 4 % at 999 modes and 9 % at 2962
    if (g + 1 < ge) {
      take_next(g + 1);  // k0, k1, last, rows_off now describe window g + 1
      vr.advance(lane);
    }
    if constexpr (WinCfg<MODE, NT3>::SEPRED) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.empty[slot]);
      if (last_c) tile_reduce_store<MT, NT, NWC, true>(rows_off_c, rows, S.red, acc, warp, lane, s0, nstates, OUT, ldo, bar_id);
    } else {
      // the reduction buffer is the finished window's slot: the consumers' barrier inside frees it of readers, and it
      // is released to the producers only after the sums
      if (last_c) tile_reduce_store<MT, NT, NWC, true>(rows_off_c, rows, cur, acc, warp, lane, s0, nstates, OUT, ldo, bar_id);
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.empty[slot]);
    }
    fresh = last_c;
    dw_n = dw_nn;
  }
}

// NT3: the pack has tiles of three n-tiles (CA_NT3); a separate instantiation, so that the register allocation of the
// two-n-tile kernel does not change with it
// SPLIT: several CTAs share a block of states (parts.n > 1).  A separate instantiation, so that the one-CTA-per-block
// code of the large ensembles stays exactly what it was (with run-time ranges ptxas arranged the one-n-tile k-steps
// of the three-n-tile variant in blocks of one again: +4 % at m = 2962).
template <int MODE, bool NT3, bool SPLIT>
__global__ void __launch_bounds__(WinCfg<MODE, NT3>::THREADS, 1)
bilinear_windowed_kernel(const WindowDesc* __restrict__ windows, const WindowParts parts, int nwin_nt3, int nwin_nt2,
                         const int32_t* __restrict__ modes, const uint32_t* __restrict__ pairs,
                         const double* __restrict__ vals, const int32_t* __restrict__ rows, int wmax, int one_mode,
                         const double* __restrict__ X, const double* __restrict__ Y, int64_t nstates, int64_t ldx,
                         double* __restrict__ OUT, int64_t ldo) {
  constexpr int ROW = 8 * kWinMT;
  using Cfg = WinCfg<MODE, NT3>;
  constexpr int NWC = Cfg::NWC, NWP = Cfg::NWP, TEAMS = Cfg::TEAMS;
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // this CTA's range of windows (whole tiles): CTAs blockIdx.x % parts.n of the same block of states share the tiles
  const int nparts = SPLIT ? parts.n : 1;
  const int part = SPLIT ? blockIdx.x % nparts : 0;
  const int w0 = SPLIT ? parts.first[part] : 0, w1 = SPLIT ? parts.first[part + 1] : parts.first[1];
  if (w1 <= w0) return;
  const size_t team_doubles = windowed_team_bytes<MODE, NT3>(wmax) / sizeof(double);
  if (tid == 0) {
    for (int t = 0; t < TEAMS; ++t) {
      const WinSmem St = carve_window_smem<MODE, NT3>(smem + t * team_doubles, wmax);
      for (int q = 0; q < Cfg::SLOTS; ++q) {
        mbar_init(&St.full[q], NWP * 32);
        mbar_init(&St.empty[q], NWC);
      }
    }
  }
  __syncthreads();  // the only CTA-wide barrier
  const bool producer = warp >= TEAMS * NWC;
  const int team = producer ? (warp - TEAMS * NWC) / NWP : warp / NWC;
  const WinSmem S = carve_window_smem<MODE, NT3>(smem + team * team_doubles, wmax);
  if (producer) {
    window_producer<MODE, NT3>(windows, w0, w1, nparts, modes, pairs, S, one_mode, X, Y, nstates, ldx, team,
                               tid - (TEAMS * NWC + team * NWP) * 32);
    return;
  }
  const int cwarp = warp - team * NWC;  // consumer warp within the team
  const int64_t nblocks = (nstates + ROW - 1) / ROW;
  RingPos pos;
  const int c3 = nwin_nt3, c2 = nwin_nt3 + nwin_nt2;  // class boundaries in the window list: [0, c3) three n-tiles, [c3, c2) two, rest one
  for (int64_t blk = (int64_t)(blockIdx.x / nparts) * TEAMS + team; blk < nblocks; blk += (int64_t)(gridDim.x / nparts) * TEAMS) {
    const int64_t s0 = blk * ROW;
    if constexpr (SPLIT) {
      if constexpr (NT3)
        run_windows<3, MODE, NT3>(windows, w0, min(w1, c3), vals, rows, S, pos, cwarp, lane, s0, nstates, OUT, ldo, 1 + team);
      run_windows<2, MODE, NT3>(windows, max(w0, c3), min(w1, c2), vals, rows, S, pos, cwarp, lane, s0, nstates, OUT, ldo,
                                1 + team);
      run_windows<1, MODE, NT3>(windows, max(w0, c2), w1, vals, rows, S, pos, cwarp, lane, s0, nstates, OUT, ldo, 1 + team);
    } else {
      if constexpr (NT3)
        run_windows<3, MODE, NT3>(windows, 0, c3, vals, rows, S, pos, cwarp, lane, s0, nstates, OUT, ldo, 1 + team);
      run_windows<2, MODE, NT3>(windows, c3, c2, vals, rows, S, pos, cwarp, lane, s0, nstates, OUT, ldo, 1 + team);
      run_windows<1, MODE, NT3>(windows, c2, w1, vals, rows, S, pos, cwarp, lane, s0, nstates, OUT, ldo, 1 + team);
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
