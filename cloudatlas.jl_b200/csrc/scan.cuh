// In-tree device-wide exclusive prefix sum (replaces cub::DeviceScan: the library's device code is all hand-written).
//
// Three-phase scan: every CTA scans a tile of kScanTile items (thread-sequential runs of kScanItems, warp shuffles for
// the thread totals), the tile totals are scanned recursively, a last pass adds the tile offsets.  Memory-bound:
// 2 reads + 2 writes per item; the inputs it serves (per-(i,j) entry counts of the assembly, bitmap word popcounts and
// non-zero flags of the device model build) are a few 10^6 .. 10^8 items, i.e. well under a millisecond.
#pragma once
#include <cstdint>

#include "common.h"

namespace ca {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

// Scratch for the tile totals of every level; grows on demand, reusable across calls on the same stream.
struct ScanScratch {
  DeviceBuffer<int64_t> sums;
};

template <typename Tin>
__global__ void __launch_bounds__(kScanThreads)
scan_tile_kernel(const Tin* in, int64_t* out, int64_t n, int64_t* tile_sums) {
  __shared__ int64_t warp_tot[kScanThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int64_t v[kScanItems];
  int64_t tsum = 0;
#pragma unroll
  for (int q = 0; q < kScanItems; ++q) {
    v[q] = base + q < n ? (int64_t)in[base + q] : 0;  // read before any write: in may alias out
    tsum += v[q];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int64_t incl = tsum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int64_t up = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += up;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  int64_t woff = 0;
  for (int w = 0; w < warp; ++w) woff += warp_tot[w];
  int64_t run = woff + incl - tsum;
#pragma unroll
  for (int q = 0; q < kScanItems; ++q) {
    if (base + q < n) out[base + q] = run;
    run += v[q];
  }
  if (threadIdx.x == kScanThreads - 1 && tile_sums) tile_sums[blockIdx.x] = run;
}

static __global__ void __launch_bounds__(kScanThreads)
scan_add_kernel(int64_t* __restrict__ out, int64_t n, const int64_t* __restrict__ tile_offs) {
  const int64_t off = tile_offs[blockIdx.x];
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  for (int q = threadIdx.x; q < kScanTile; q += kScanThreads)
    if (base + q < n) out[base + q] += off;
}

// out[i] = in[0] + ... + in[i-1] for i in [0, n); `in` may alias `out` when Tin is int64_t.
// To obtain the grand total, scan n + 1 items with in[n] readable (any value): out[n] is the total.
template <typename Tin>
void exclusive_scan(const Tin* in, int64_t* out, int64_t n, ScanScratch& scratch, cudaStream_t st) {
  if (n <= 0) return;
  // levels: n -> ceil(n / tile) -> ... -> 1
  int64_t need = 0;
  for (int64_t k = (n + kScanTile - 1) / kScanTile; k > 1; k = (k + kScanTile - 1) / kScanTile) need += k;
  need += 1;
  if ((int64_t)scratch.sums.count < need) scratch.sums.alloc((size_t)need);
  const int64_t tiles = (n + kScanTile - 1) / kScanTile;
  if (tiles == 1) {
    scan_tile_kernel<Tin><<<1, kScanThreads, 0, st>>>(in, out, n, nullptr);
    CA_CUDA(cudaGetLastError());
    count_launch();
    return;
  }
  int64_t* level = scratch.sums.ptr;
  scan_tile_kernel<Tin><<<(unsigned)tiles, kScanThreads, 0, st>>>(in, out, n, level);
  CA_CUDA(cudaGetLastError());
  count_launch();
  // scan the tile totals in place (deeper levels use the scratch space behind this level)
  struct Level { int64_t* ptr; int64_t n; };
  Level stack[8];
  int depth = 0;
  stack[depth++] = Level{level, tiles};
  while (stack[depth - 1].n > kScanTile) {
    const Level cur = stack[depth - 1];
    const int64_t t = (cur.n + kScanTile - 1) / kScanTile;
    int64_t* next = cur.ptr + cur.n;
    scan_tile_kernel<int64_t><<<(unsigned)t, kScanThreads, 0, st>>>(cur.ptr, cur.ptr, cur.n, next);
    CA_CUDA(cudaGetLastError());
    count_launch();
    stack[depth++] = Level{next, t};
  }
  {
    const Level top = stack[depth - 1];
    scan_tile_kernel<int64_t><<<1, kScanThreads, 0, st>>>(top.ptr, top.ptr, top.n, nullptr);
    CA_CUDA(cudaGetLastError());
    count_launch();
  }
  for (int d = depth - 1; d >= 1; --d) {  // push the offsets of level d down into level d-1
    const Level lo = stack[d - 1];
    scan_add_kernel<<<(unsigned)stack[d].n, kScanThreads, 0, st>>>(lo.ptr, lo.n, stack[d].ptr);
    CA_CUDA(cudaGetLastError());
    count_launch();
  }
  scan_add_kernel<<<(unsigned)tiles, kScanThreads, 0, st>>>(out, n, level);
  CA_CUDA(cudaGetLastError());
  count_launch();
}

}  // namespace ca
