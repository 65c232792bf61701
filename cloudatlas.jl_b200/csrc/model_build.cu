// ODEModel construction on the device (reference: src/ODEModels.jl:51-58 -- COO of N, Bfact = lu(B), closures).
//
// ca_model_create_dev takes the assembled operators as DEVICE arrays (straight from ca_assembly_get_*_dev or from the
// all-gathered slabs) and builds the evaluation side of the model without moving the operator through the host:
//   fold   Q = B^-1 N, L1 = B^-1 A1, L2 = B^-1 A2   per block of the block-diagonal B (blocks found on the host from
//          B's pattern, inverted there -- they are tiny -- and applied here)
//   sym    (j,k) ~ (k,j) merged, linear placeholders (j, ONE) added, exact zeros dropped -> per-row pair lists (CSR)
//   pack   row sketches, pair-set unions for the row clustering, per-tile pair unions, value placement into the
//          DMMA fragment layout, linear slots
// Sets of pair ids are BITMAPS over the key space (m^2 resp. (m+1)^2 bits: 1.1 MB at m = 2962), one per block / tile:
// union = atomicOr, |A and B| = bit tests, sorted list = popcount prefix (scan.cuh) + expansion, rank = prefix lookup.
// Every DECISION of the packer (greedy row clustering, tile split, segment / window formation, warp schedule) is made
// by the same host code as in the host build (pack.cpp, staged interface) on descriptors that are a few MB; the host
// never sees the operator's values.  The result is the same pack, bit for bit, as ca_model_create builds on the host
// (tests/test_model_build_gpu.py compares the digests of every array).
//
// Arithmetic matches the host folding exactly: sums over b ascending, zeros of B^-1 skipped, separate multiply and add
// (__dmul_rn / __dadd_rn: no FMA contraction, as in the host code compiled for baseline x86-64).
#include <algorithm>
#include <cmath>
#include <functional>
#include <map>
#include <memory>
#include <numeric>

#include "bilinear.h"
#include "model.h"
#include "scan.cuh"

using namespace ca;

namespace {

// ------------------------------------------------------------------------------------------------ scan adaptors
struct PopcWord {  // exclusive_scan over popcounts of bitmap words
  uint32_t w;
  __host__ __device__ operator int64_t() const {
#ifdef __CUDA_ARCH__
    return (int64_t)__popc(w);
#else
    return (int64_t)__builtin_popcount(w);
#endif
  }
};
struct NonZero {  // exclusive_scan over "value != 0" flags of a dense array
  double v;
  __host__ __device__ operator int64_t() const { return v != 0.0 ? 1 : 0; }
};

// global index (into the concatenated sorted lists of a bitmap family) of `key` in set `set`, -1 if absent
__device__ __forceinline__ int64_t bitmap_rank(const uint32_t* __restrict__ bits, const int64_t* __restrict__ prefix,
                                               int64_t set, int64_t words, uint32_t key) {
  const int64_t w = set * words + (key >> 5);
  const uint32_t word = bits[w];
  const uint32_t bit = 1u << (key & 31);
  if (!(word & bit)) return -1;
  return prefix[w] + __popc(word & (bit - 1));
}

// keys of a bitmap family as concatenated ascending lists: out[prefix[w] + t] = 32 * (w mod words) + bit
__global__ void expand_bits_kernel(const uint32_t* __restrict__ bits, const int64_t* __restrict__ prefix,
                                   int64_t nwords, int64_t words, uint32_t* __restrict__ out) {
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += (int64_t)gridDim.x * blockDim.x) {
    uint32_t word = bits[w];
    if (!word) continue;
    int64_t at = prefix[w];
    const uint32_t base = (uint32_t)(w % words) * 32u;
    while (word) {
      const int b = __ffs(word) - 1;
      out[at++] = base + (uint32_t)b;
      word &= word - 1;
    }
  }
}

// ------------------------------------------------------------------------------------------------ input check
// flags[0]: index outside 1..m; flags[1]: not ascending in (i,j,k); flags[2]: duplicate (i,j,k)
__global__ void check_coo_kernel(const int64_t* __restrict__ di, const int64_t* __restrict__ dj,
                                 const int64_t* __restrict__ dk, int64_t nnz, int64_t m, int32_t* flags) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nnz; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = di[r], j = dj[r], k = dk[r];
    if (i < 1 || i > m || j < 1 || j > m || k < 1 || k > m) flags[0] = 1;
    if (r > 0) {
      const int64_t pi = di[r - 1], pj = dj[r - 1], pk = dk[r - 1];
      if (pi > i || (pi == i && (pj > j || (pj == j && pk > k)))) flags[1] = 1;
      if (pi == i && pj == j && pk == k) flags[2] = 1;
    }
  }
}

// compact device copy of N as given (0-based int32 indices), kept by the model for the time stepper's raw operator
__global__ void copy_coo_kernel(const int64_t* __restrict__ di, const int64_t* __restrict__ dj,
                                const int64_t* __restrict__ dk, const double* __restrict__ dv, int64_t nnz,
                                Term* __restrict__ out) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nnz; r += (int64_t)gridDim.x * blockDim.x)
  {
    Term t;
    t.i = (int32_t)(di[r] - 1);
    t.a = (int32_t)(dj[r] - 1);
    t.b = (int32_t)(dk[r] - 1);
    t.v = dv[r];
    out[r] = t;
  }
}

// ------------------------------------------------------------------------------------------------ fold
struct BlockMap {               // B's blocks on the device
  const int32_t* blk_of;        // [m] block of a row
  const int32_t* pos_in;        // [m] position of a row inside its block
  const int32_t* row_off;       // [nb+1] offsets into rows
  const int32_t* rows;          // concatenated member lists (ascending inside a block)
  const double* binv;           // concatenated row-major g x g inverses
  const int64_t* binv_off;      // [nb+1]
};

// ordered pair ids j * m + k of every N entry, marked in the bitmap of the entry's block
__global__ void mark_ordered_kernel(const int64_t* __restrict__ di, const int64_t* __restrict__ dj,
                                    const int64_t* __restrict__ dk, int64_t nnz, int64_t m, BlockMap bm, int64_t words,
                                    uint32_t* __restrict__ bits) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nnz; r += (int64_t)gridDim.x * blockDim.x) {
    const int32_t b = bm.blk_of[di[r] - 1];
    const uint32_t key = (uint32_t)((dj[r] - 1) * m + (dk[r] - 1));
    uint32_t* w = bits + (int64_t)b * words + (key >> 5);
    const uint32_t bit = 1u << (key & 31);
    if (!(*(volatile uint32_t*)w & bit)) atomicOr(w, bit);  // the g rows of a block set the same bits: most are set already
  }
}

// D[block][row in block][rank of pair] = N value (the input is strictly ascending: no duplicates to sum)
__global__ void scatter_dense_kernel(const int64_t* __restrict__ di, const int64_t* __restrict__ dj,
                                     const int64_t* __restrict__ dk, const double* __restrict__ dv, int64_t nnz,
                                     int64_t m, BlockMap bm, int64_t words, const uint32_t* __restrict__ bits,
                                     const int64_t* __restrict__ prefix, const int64_t* __restrict__ dense_off,
                                     double* __restrict__ D) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nnz; r += (int64_t)gridDim.x * blockDim.x) {
    const int32_t i = (int32_t)(di[r] - 1);
    const int32_t b = bm.blk_of[i];
    const uint32_t key = (uint32_t)((dj[r] - 1) * m + (dk[r] - 1));
    const int64_t u0 = prefix[(int64_t)b * words], nu = prefix[(int64_t)(b + 1) * words] - u0;
    const int64_t p = bitmap_rank(bits, prefix, b, words, key) - u0;
    D[dense_off[b] + (int64_t)bm.pos_in[i] * nu + p] = dv[r];
  }
}

// R = Binv * D per block.  grid (chunks of 128 pairs, blocks); blockDim (128, 4): a thread owns one pair column and the
// rows a = ty, ty + 4, ...  D rows are re-read from L1 for every a.
__global__ void __launch_bounds__(512)
fold_quadratic_kernel(const double* __restrict__ D, double* __restrict__ R, BlockMap bm, int64_t words,
                      const int64_t* __restrict__ prefix, const int64_t* __restrict__ dense_off) {
  const int b = blockIdx.y;
  const int64_t u0 = prefix[(int64_t)b * words], nu = prefix[(int64_t)(b + 1) * words] - u0;
  const int64_t p = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if ((int64_t)blockIdx.x * 128 >= nu) return;
  const int g = bm.row_off[b + 1] - bm.row_off[b];
  extern __shared__ double binv_s[];  // g x g
  const double* binv = bm.binv + bm.binv_off[b];
  for (int e = threadIdx.y * 128 + threadIdx.x; e < g * g; e += 512) binv_s[e] = binv[e];
  __syncthreads();
  if (p >= nu) return;
  const double* Db = D + dense_off[b] + p;
  double* Rb = R + dense_off[b] + p;
  for (int a = threadIdx.y; a < g; a += 4) {
    double s = 0.0;
    for (int bb = 0; bb < g; ++bb) {
      const double w = binv_s[a * g + bb];
      if (w == 0.0) continue;
      s = __dadd_rn(s, __dmul_rn(w, Db[(int64_t)bb * nu]));
    }
    Rb[(int64_t)a * nu] = s;
  }
}

// L1 = Binv A1, L2 = Binv A2 (column-major m x m) and their row-major copies for the streaming kernel
__global__ void fold_linear_kernel(const double* __restrict__ A1, const double* __restrict__ A2, int m, BlockMap bm,
                                   double* __restrict__ L1, double* __restrict__ L2, double* __restrict__ L1r,
                                   double* __restrict__ L2r) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)m * m) return;
  const int i = (int)(t % m), j = (int)(t / m);
  const int b = bm.blk_of[i], a = bm.pos_in[i];
  const int g = bm.row_off[b + 1] - bm.row_off[b];
  const int32_t* G = bm.rows + bm.row_off[b];
  const double* binv = bm.binv + bm.binv_off[b] + (int64_t)a * g;
  double s1 = 0.0, s2 = 0.0;
  for (int bb = 0; bb < g; ++bb) {
    const double w = binv[bb];
    s1 = __dadd_rn(s1, __dmul_rn(w, A1[G[bb] + (int64_t)m * j]));
    s2 = __dadd_rn(s2, __dmul_rn(w, A2[G[bb] + (int64_t)m * j]));
  }
  L1[t] = s1;
  L2[t] = s2;
  L1r[(int64_t)i * m + j] = s1;
  L2r[(int64_t)i * m + j] = s2;
}

// ------------------------------------------------------------------------------------------------ symmetrise
// symmetric pair ids min * nin + max of the ordered pairs of every block
__global__ void mark_sym_kernel(const uint32_t* __restrict__ ukeys, const int64_t* __restrict__ uprefix, int64_t uwords,
                                int nb, int64_t total_u, int m, int nin, int64_t swords, uint32_t* __restrict__ sbits) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total_u; e += (int64_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = nb;  // block of element e: last b with uprefix[b * uwords] <= e
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (uprefix[(int64_t)mid * uwords] <= e) lo = mid; else hi = mid;
    }
    const uint32_t key = ukeys[e];
    const uint32_t j = key / (uint32_t)m, k = key % (uint32_t)m;
    const uint32_t s = min(j, k) * (uint32_t)nin + max(j, k);
    atomicOr(sbits + (int64_t)lo * swords + (s >> 5), 1u << (s & 31));
  }
}

// linear placeholders: pair (j, ONE) wherever a row of the block has L1 or L2 non-zero in column j
__global__ void mark_linear_kernel(const double* __restrict__ L1, const double* __restrict__ L2, int m, int nin,
                                   BlockMap bm, int64_t swords, uint32_t* __restrict__ sbits) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)m * m) return;
  if (L1[t] == 0.0 && L2[t] == 0.0) return;
  const int i = (int)(t % m), j = (int)(t / m);
  const uint32_t s = (uint32_t)j * (uint32_t)nin + (uint32_t)m;
  atomicOr(sbits + (int64_t)bm.blk_of[i] * swords + (s >> 5), 1u << (s & 31));
}

// Vs[block][row in block][rank of symmetric pair] = folded value of (j,k) plus that of (k,j); 1 for a linear placeholder.
// grid (chunks of 256 symmetric pairs, blocks)
__global__ void __launch_bounds__(256)
sym_values_kernel(const double* __restrict__ R, const uint32_t* __restrict__ ubits, const int64_t* __restrict__ uprefix,
                  int64_t uwords, const int64_t* __restrict__ udense_off, const uint32_t* __restrict__ skeys,
                  const int64_t* __restrict__ sprefix, int64_t swords, const int64_t* __restrict__ sdense_off,
                  const double* __restrict__ L1, const double* __restrict__ L2, int m, int nin, BlockMap bm,
                  double* __restrict__ Vs) {
  const int b = blockIdx.y;
  const int64_t s0 = sprefix[(int64_t)b * swords], ns = sprefix[(int64_t)(b + 1) * swords] - s0;
  const int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (q >= ns) return;
  const int g = bm.row_off[b + 1] - bm.row_off[b];
  const int32_t* G = bm.rows + bm.row_off[b];
  const uint32_t s = skeys[s0 + q];
  const uint32_t x = s / (uint32_t)nin, y = s % (uint32_t)nin;
  double* out = Vs + sdense_off[b] + q;
  if (y == (uint32_t)m) {
    for (int a = 0; a < g; ++a) {
      const int64_t at = G[a] + (int64_t)m * x;
      out[(int64_t)a * ns] = (L1[at] != 0.0 || L2[at] != 0.0) ? 1.0 : 0.0;
    }
    return;
  }
  const int64_t u0 = uprefix[(int64_t)b * uwords], nu = uprefix[(int64_t)(b + 1) * uwords] - u0;
  const int64_t r1 = bitmap_rank(ubits, uprefix, b, uwords, x * (uint32_t)m + y);
  const int64_t r2 = x != y ? bitmap_rank(ubits, uprefix, b, uwords, y * (uint32_t)m + x) : -1;
  const double* Rb = R + udense_off[b];
  for (int a = 0; a < g; ++a) {
    const double v1 = r1 >= 0 ? Rb[(int64_t)a * nu + (r1 - u0)] : 0.0;
    const double v2 = r2 >= 0 ? Rb[(int64_t)a * nu + (r2 - u0)] : 0.0;
    out[(int64_t)a * ns] = __dadd_rn(v1, v2);
  }
}

// non-zeros of Vs -> CSR (keys, vals) in (block, row in block, pair) order; pos = exclusive scan of the non-zero flags
__global__ void __launch_bounds__(256)
compact_rows_kernel(const double* __restrict__ Vs, const int64_t* __restrict__ pos, const uint32_t* __restrict__ skeys,
                    const int64_t* __restrict__ sprefix, int64_t swords, const int64_t* __restrict__ sdense_off,
                    BlockMap bm, uint32_t* __restrict__ keys, double* __restrict__ vals, int64_t* __restrict__ row_begin,
                    int64_t* __restrict__ row_end) {
  const int b = blockIdx.y;
  const int64_t s0 = sprefix[(int64_t)b * swords], ns = sprefix[(int64_t)(b + 1) * swords] - s0;
  const int g = bm.row_off[b + 1] - bm.row_off[b];
  const int32_t* G = bm.rows + bm.row_off[b];
  const int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (q == 0)
    for (int a = 0; a < g; ++a) {  // (blocks without pairs: empty rows)
      row_begin[G[a]] = pos[sdense_off[b] + (int64_t)a * ns];
      row_end[G[a]] = pos[sdense_off[b] + (int64_t)(a + 1) * ns];
    }
  if (q >= ns) return;
  const uint32_t s = skeys[s0 + q];
  for (int a = 0; a < g; ++a) {
    const int64_t at = sdense_off[b] + (int64_t)a * ns + q;
    const double v = Vs[at];
    if (v != 0.0) {
      keys[pos[at]] = s;
      vals[pos[at]] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ row sets
__device__ __forceinline__ uint64_t mix64_dev(uint64_t x) {
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}

// min-hash sketch of every row (pack.h: pack_sketch_hash_host), one CTA per row
__global__ void __launch_bounds__(256)
sketch_kernel(const uint32_t* __restrict__ keys, const int64_t* __restrict__ row_begin,
              const int64_t* __restrict__ row_end, uint64_t* __restrict__ sketch) {
  const int i = blockIdx.x;
  uint64_t best[kPackSketch];
#pragma unroll
  for (int h = 0; h < kPackSketch; ++h) best[h] = ~0ull;
  for (int64_t e = row_begin[i] + threadIdx.x; e < row_end[i]; e += 256) {
    const uint64_t k = (uint64_t)keys[e] * 0x100000001b3ull;
#pragma unroll
    for (int h = 0; h < kPackSketch; ++h) best[h] = min(best[h], mix64_dev(k + (uint64_t)h * 0x51ed27ull));
  }
  __shared__ uint64_t red[8][kPackSketch];
#pragma unroll
  for (int h = 0; h < kPackSketch; ++h) {
    uint64_t v = best[h];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, d));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][h] = v;
  }
  __syncthreads();
  if (threadIdx.x < kPackSketch) {
    uint64_t v = red[0][threadIdx.x];
    for (int w = 1; w < 8; ++w) v = min(v, red[w][threadIdx.x]);
    sketch[(int64_t)i * kPackSketch + threadIdx.x] = v;
  }
}

// bits |= keys of one row
__global__ void mark_row_kernel(const uint32_t* __restrict__ keys, int64_t begin, int64_t end, uint32_t* __restrict__ bits) {
  for (int64_t e = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < end; e += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t s = keys[e];
    atomicOr(bits + (s >> 5), 1u << (s & 31));
  }
}

// counts[c] = |keys(cand[c]) and bits|; grid (parts, candidates)
__global__ void __launch_bounds__(256)
intersect_kernel(const uint32_t* __restrict__ keys, const int64_t* __restrict__ row_begin,
                 const int64_t* __restrict__ row_end, const int32_t* __restrict__ cand,
                 const uint32_t* __restrict__ bits, unsigned long long* __restrict__ counts) {
  const int r = cand[blockIdx.y];
  int n = 0;
  for (int64_t e = row_begin[r] + (int64_t)blockIdx.x * 256 + threadIdx.x; e < row_end[r]; e += (int64_t)gridDim.x * 256) {
    const uint32_t s = keys[e];
    n += (bits[s >> 5] >> (s & 31)) & 1u;
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) n += __shfl_xor_sync(0xffffffffu, n, d);
  if ((threadIdx.x & 31) == 0 && n) atomicAdd(counts + blockIdx.y, (unsigned long long)n);
}

// tile bitmaps: bits[tile of row] |= keys(row); grid (parts, rows)
__global__ void __launch_bounds__(256)
mark_tiles_kernel(const uint32_t* __restrict__ keys, const int64_t* __restrict__ row_begin,
                  const int64_t* __restrict__ row_end, const int32_t* __restrict__ tile_of_row, int64_t words,
                  uint32_t* __restrict__ bits) {
  const int i = blockIdx.y;
  const int t = tile_of_row[i];
  if (t < 0) return;
  uint32_t* tb = bits + (int64_t)t * words;
  for (int64_t e = row_begin[i] + (int64_t)blockIdx.x * 256 + threadIdx.x; e < row_end[i]; e += (int64_t)gridDim.x * 256) {
    const uint32_t s = keys[e];
    uint32_t* w = tb + (s >> 5);
    const uint32_t bit = 1u << (s & 31);
    if (!(*(volatile uint32_t*)w & bit)) atomicOr(w, bit);
  }
}

// number of distinct pairs over all tiles: popcount of the OR of the tile bitmaps
__global__ void distinct_pairs_kernel(const uint32_t* __restrict__ bits, int ntiles, int64_t words,
                                      unsigned long long* __restrict__ count) {
  int n = 0;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += (int64_t)gridDim.x * blockDim.x) {
    uint32_t acc = 0;
    for (int t = 0; t < ntiles; ++t) acc |= bits[(int64_t)t * words + w];
    n += __popc(acc);
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) n += __shfl_xor_sync(0xffffffffu, n, d);
  if ((threadIdx.x & 31) == 0 && n) atomicAdd(count, (unsigned long long)n);
}

// values into the fragment layout [k-step][n-tile][lane = 4 * (row % 8) + (pair % 4)] of the row's tile; the pairs
// (j, ONE) are the linear slots that set_R re-targets: their positions and L1 / L2 values are appended to the slot lists
__global__ void __launch_bounds__(256)
place_values_kernel(const uint32_t* __restrict__ keys, const double* __restrict__ vals,
                    const int64_t* __restrict__ row_begin, const int64_t* __restrict__ row_end,
                    const int32_t* __restrict__ tile_of_row, const int32_t* __restrict__ slot_of_row,
                    const int64_t* __restrict__ tile_vals_off, const int32_t* __restrict__ tile_nt, int64_t words,
                    const uint32_t* __restrict__ bits, const int64_t* __restrict__ prefix,
                    const int32_t* __restrict__ where, int m, int nin, const double* __restrict__ L1,
                    const double* __restrict__ L2, double* __restrict__ V, unsigned long long* __restrict__ nlin,
                    int64_t* __restrict__ lin_pos, double* __restrict__ lin_l1, double* __restrict__ lin_l2) {
  const int i = blockIdx.y;
  const int t = tile_of_row[i];
  if (t < 0) return;
  const int n = slot_of_row[i], nt = tile_nt[t];
  const int64_t voff = tile_vals_off[t];
  for (int64_t e = row_begin[i] + (int64_t)blockIdx.x * 256 + threadIdx.x; e < row_end[i]; e += (int64_t)gridDim.x * 256) {
    const uint32_t s = keys[e];
    const int64_t rk = bitmap_rank(bits, prefix, t, words, s);
    const int at = where[rk];
    const int64_t slot = voff + ((int64_t)(at >> 2) * nt + (n >> 3)) * 32 + (n & 7) * 4 + (at & 3);
    V[slot] = vals[e];
    const uint32_t a = s / (uint32_t)nin, b = s % (uint32_t)nin;
    if (b == (uint32_t)m && a != (uint32_t)m) {
      const unsigned long long q = atomicAdd(nlin, 1ull);
      lin_pos[q] = slot;
      lin_l1[q] = L1[i + (int64_t)m * a];
      lin_l2[q] = L2[i + (int64_t)m * a];
    }
  }
}

// ------------------------------------------------------------------------------------------------ streaming form
// flag of "quadratic entry" (second factor is a mode, not ONE) per CSR entry, for the compaction scan
__global__ void quad_flags_kernel(const uint32_t* __restrict__ keys, int64_t n, int nin, int m, int32_t* __restrict__ flag) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    flag[e] = (int)(keys[e] % (uint32_t)nin) != m;
}
__global__ void stream_fill_kernel(const uint32_t* __restrict__ keys, const double* __restrict__ vals, int64_t n, int nin,
                                   int m, const int64_t* __restrict__ pos, uint32_t* __restrict__ ab,
                                   double* __restrict__ val) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t a = keys[e] / (uint32_t)nin, b = keys[e] % (uint32_t)nin;
    if ((int)b == m) continue;
    ab[pos[e]] = a | (b << 16);
    val[pos[e]] = vals[e];
  }
}

// out[i] = pos[row_begin[i]], out[m + i] = pos[row_end[i]]
__global__ void gather_ranges_kernel(const int64_t* __restrict__ pos, const int64_t* __restrict__ row_begin,
                                     const int64_t* __restrict__ row_end, int m, int64_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  out[i] = pos[row_begin[i]];
  out[m + i] = pos[row_end[i]];
}

unsigned grid_for(int64_t n, int threads = 256, int64_t cap = 148 * 32) {
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + threads - 1) / threads, cap));
}

// pinned host staging (small D2H / H2D round trips of the clustering loop)
template <typename T>
struct PinnedBuffer {
  T* ptr = nullptr;
  size_t count = 0;
  ~PinnedBuffer() {
    if (ptr) cudaFreeHost(ptr);
  }
  void need(size_t n) {
    if (n <= count) return;
    if (ptr) cudaFreeHost(ptr);
    CA_CUDA(cudaMallocHost(&ptr, n * sizeof(T)));
    count = n;
  }
};

// The rows' pair sets on the device, as the clustering stage of the packer sees them (pack.h: RowSetOps).
struct DeviceRowSets final : RowSetOps {
  int32_t m;
  int64_t words;
  const uint32_t* keys;
  const int64_t *row_begin, *row_end;  // device
  std::vector<int64_t> hbegin, hend;   // host copies
  std::vector<uint64_t> hsketch;
  DeviceBuffer<uint32_t> bits;         // the open set
  DeviceBuffer<int32_t> dcand;
  DeviceBuffer<unsigned long long> dcount;
  PinnedBuffer<int32_t> pcand;
  PinnedBuffer<unsigned long long> pcount;
  std::vector<int32_t> last_cand;
  std::vector<int64_t> last_union;
  int64_t usize = 0;
  cudaStream_t st;
  int32_t nrows() const override { return m; }
  int64_t nkeys(int32_t r) const override { return hend[r] - hbegin[r]; }
  const uint64_t* sketch(int32_t r) const override { return hsketch.data() + (size_t)r * kPackSketch; }
  int batch_hint() const override { return 1 << 30; }
  void mark(int32_t r) {
    const int64_t n = nkeys(r);
    if (n <= 0) return;
    mark_row_kernel<<<grid_for(n), 256, 0, st>>>(keys, hbegin[r], hend[r], bits.ptr);
    CA_CUDA(cudaGetLastError());
    count_launch();
  }
  void open(int32_t seed) override {
    CA_CUDA(cudaMemsetAsync(bits.ptr, 0, (size_t)words * sizeof(uint32_t), st));
    mark(seed);
    usize = nkeys(seed);
    last_cand.clear();
  }
  int64_t open_size() const override { return usize; }
  void union_sizes(const int32_t* cand, int n, int64_t* out) override {
    if (n <= 0) return;
    if (dcand.count < (size_t)n) {
      dcand.alloc((size_t)n + 1024);
      dcount.alloc((size_t)n + 1024);
    }
    pcand.need((size_t)n);
    pcount.need((size_t)n);
    std::copy(cand, cand + n, pcand.ptr);
    CA_CUDA(cudaMemcpyAsync(dcand.ptr, pcand.ptr, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CA_CUDA(cudaMemsetAsync(dcount.ptr, 0, (size_t)n * sizeof(unsigned long long), st));
    intersect_kernel<<<dim3(8, (unsigned)n), 256, 0, st>>>(keys, row_begin, row_end, dcand.ptr, bits.ptr, dcount.ptr);
    CA_CUDA(cudaGetLastError());
    count_launch();
    CA_CUDA(cudaMemcpyAsync(pcount.ptr, dcount.ptr, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CA_CUDA(cudaStreamSynchronize(st));
    last_cand.assign(cand, cand + n);
    last_union.resize((size_t)n);
    for (int q = 0; q < n; ++q) last_union[q] = out[q] = usize + nkeys(cand[q]) - (int64_t)pcount.ptr[q];
  }
  void absorb(int32_t r) override {
    mark(r);
    for (size_t q = 0; q < last_cand.size(); ++q)
      if (last_cand[q] == r) {
        usize = last_union[q];
        return;
      }
    fail(CA_ERR_INVALID, "device row sets: absorb() of a row whose union size was not asked for");
  }
};

double now_seconds() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

// In-place LU with partial pivoting of a dense g x g block (row-major), then inverse by solving for the identity --
// the per-block equivalent of  Bfact = lu(B); Bfact \ I  (ODEModels.jl:54).  Same operation order as the host build.
bool invert_block_rm(std::vector<double>& A, int g, std::vector<double>& inv) {
  std::vector<int> piv(g);
  for (int c = 0; c < g; ++c) {
    int p = c;
    for (int r = c + 1; r < g; ++r)
      if (std::fabs(A[r * g + c]) > std::fabs(A[p * g + c])) p = r;
    if (A[p * g + c] == 0.0) return false;
    piv[c] = p;
    if (p != c)
      for (int k = 0; k < g; ++k) std::swap(A[c * g + k], A[p * g + k]);
    for (int r = c + 1; r < g; ++r) {
      const double f = A[r * g + c] / A[c * g + c];
      A[r * g + c] = f;
      for (int k = c + 1; k < g; ++k) A[r * g + k] -= f * A[c * g + k];
    }
  }
  inv.assign((size_t)g * g, 0.0);
  std::vector<double> b(g);
  for (int col = 0; col < g; ++col) {
    std::fill(b.begin(), b.end(), 0.0);
    b[col] = 1.0;
    for (int c = 0; c < g; ++c)
      if (piv[c] != c) std::swap(b[c], b[piv[c]]);
    for (int r = 1; r < g; ++r) {
      double s = b[r];
      for (int k = 0; k < r; ++k) s -= A[r * g + k] * b[k];
      b[r] = s;
    }
    for (int r = g - 1; r >= 0; --r) {
      double s = b[r];
      for (int k = r + 1; k < g; ++k) s -= A[r * g + k] * b[k];
      b[r] = s / A[r * g + r];
    }
    for (int r = 0; r < g; ++r) inv[(size_t)r * g + col] = b[r];
  }
  return true;
}

constexpr int kMaxBlockRows = 64;                       // larger blocks of B: host build
constexpr int64_t kMaxBitmapBytes = (int64_t)12 << 30;  // per bitmap family

}  // namespace

namespace ca {

// Expands the device-resident folded operator into host terms (lazy: Df structure, generated small kernel, solver CSR).
void FoldedDevice::to_host_terms(std::vector<Term>& out, int32_t m) const {
  out.clear();
  if (total_u == 0) return;
  std::vector<uint32_t> hk((size_t)total_u);
  CA_CUDA(cudaMemcpy(hk.data(), ukeys.ptr, hk.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  std::vector<double> hr((size_t)dense_off.back());
  CA_CUDA(cudaMemcpy(hr.data(), R.ptr, hr.size() * sizeof(double), cudaMemcpyDeviceToHost));
  const int nb = (int)u_off.size() - 1;
  std::vector<size_t> at((size_t)nb + 1, 0);
  parallel_chunks((size_t)nb, (size_t)nb, [&](size_t b0, size_t b1, size_t) {
    for (size_t b = b0; b < b1; ++b) {
      size_t n = 0;
      const int64_t nu = u_off[b + 1] - u_off[b];
      const int g = row_off[b + 1] - row_off[b];
      const double* Rb = hr.data() + dense_off[b];
      for (int64_t e = 0; e < (int64_t)g * nu; ++e) n += Rb[e] != 0.0;
      at[b + 1] = n;
    }
  });
  for (int b = 0; b < nb; ++b) at[(size_t)b + 1] += at[b];
  out.resize(at[(size_t)nb]);
  parallel_chunks((size_t)nb, (size_t)nb, [&](size_t b0, size_t b1, size_t) {
    for (size_t b = b0; b < b1; ++b) {
      const int64_t nu = u_off[b + 1] - u_off[b];
      const int g = row_off[b + 1] - row_off[b];
      const double* Rb = hr.data() + dense_off[b];
      const uint32_t* Ub = hk.data() + u_off[b];
      size_t w = at[b];
      for (int a = 0; a < g; ++a)
        for (int64_t p = 0; p < nu; ++p) {
          const double v = Rb[(int64_t)a * nu + p];
          if (v != 0.0) out[w++] = Term{rows[(size_t)row_off[b] + a], (int32_t)(Ub[p] / (uint32_t)m), (int32_t)(Ub[p] % (uint32_t)m), v};
        }
    }
  });
}

// off-diagonal non-zeros of B (column-major m x m) as a list of (row, column); count may exceed cap (the caller checks)
__global__ void __launch_bounds__(256)
b_offdiag_kernel(const double* __restrict__ B, int m, int2* __restrict__ list, int32_t* __restrict__ count, int32_t cap) {
  const int64_t n = (int64_t)m * m;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {  // (grid_for caps the grid)
    const int i = (int)(e % m), j = (int)(e / m);
    if (i != j && B[e] != 0.0) {
      const int32_t pos = atomicAdd(count, 1);
      if (pos < cap) list[pos] = make_int2(i, j);
    }
  }
}

// the diagonal blocks of B, row-major one after the other: out[off[b] + a g + c] = B[rows[a], rows[c]]
__global__ void __launch_bounds__(256)
b_gather_blocks_kernel(const double* __restrict__ B, int m, const int32_t* __restrict__ rows,
                       const int32_t* __restrict__ row_off, const int64_t* __restrict__ off, int nb,
                       double* __restrict__ out) {
  const int b = blockIdx.x;
  if (b >= nb) return;
  const int r0 = row_off[b], g = row_off[b + 1] - r0;
  double* o = out + off[b];
  for (int e = threadIdx.x; e < g * g; e += blockDim.x) {
    const int a = e / g, c = e % g;
    o[e] = B[rows[r0 + a] + (int64_t)m * rows[r0 + c]];
  }
}

}  // namespace ca

// Builds *h from device-resident operators.  Returns false (nothing built) when the device build does not apply and
// the caller should take the host build: unsorted or duplicated COO, a block of B larger than kMaxBlockRows, bitmaps
// over budget.  stats (may be null): seconds per phase.
bool build_model_on_device(ca_model* h, const double* dB, const double* dA1, const double* dA2, int64_t m64,
                           const int64_t* di, const int64_t* dj, const int64_t* dk, const double* dval, int64_t nnz,
                           cudaStream_t st) {
  const int m = (int)m64, nin = m + 1;
  if (nin > 32768) fail(CA_ERR_UNSUPPORTED, "ca_model_create_dev: m = %d exceeds the 15-bit mode index", m);
  PhaseTimer tm("ca_model_create_dev");
  const double t_start = now_seconds();
  ModelBuildStats& S = h->build_stats;
  S = ModelBuildStats();
  auto lap = [&](double& slot, double& t0) {
    CA_CUDA(cudaStreamSynchronize(st));
    const double t = now_seconds();
    slot += t - t0;
    t0 = t;
  };
  double t0 = t_start;

  // ---- input check
  DeviceBuffer<int32_t> dflags;
  dflags.alloc(4);
  CA_CUDA(cudaMemsetAsync(dflags.ptr, 0, 4 * sizeof(int32_t), st));
  if (nnz > 0) {
    check_coo_kernel<<<grid_for(nnz), 256, 0, st>>>(di, dj, dk, nnz, m64, dflags.ptr);
    CA_CUDA(cudaGetLastError());
    count_launch();
  }
  int32_t flags[4];
  CA_CUDA(cudaMemcpyAsync(flags, dflags.ptr, sizeof flags, cudaMemcpyDeviceToHost, st));
  // ---- blocks of B = connected components of its pattern.  B stays on the device (a copy for cnab2, which needs it
  // unfolded); only the list of its off-diagonal non-zeros and, afterwards, its diagonal blocks travel to the host,
  // where the components are found (union-find) and the blocks inverted (round 2, first version: the whole of B was
  // downloaded and scanned, 41 ms of the 0.18 s build at m = 2962)
  h->dB.alloc((size_t)m * m);
  CA_CUDA(cudaMemcpyAsync(h->dB.ptr, dB, (size_t)m * m * sizeof(double), cudaMemcpyDeviceToDevice, st));
  h->hB.clear();
  h->B_on_device = true;
  const int32_t cap = (int32_t)std::min<int64_t>((int64_t)m * (kMaxBlockRows - 1), (int64_t)1 << 30);
  DeviceBuffer<int2> d_list;
  DeviceBuffer<int32_t> d_count;
  d_list.alloc((size_t)std::max(cap, 1));
  d_count.alloc(1);
  CA_CUDA(cudaMemsetAsync(d_count.ptr, 0, sizeof(int32_t), st));
  b_offdiag_kernel<<<grid_for((int64_t)m * m), 256, 0, st>>>(dB, m, d_list.ptr, d_count.ptr, cap);
  CA_CUDA(cudaGetLastError());
  count_launch();
  int32_t noff = 0;
  CA_CUDA(cudaMemcpyAsync(&noff, d_count.ptr, sizeof noff, cudaMemcpyDeviceToHost, st));
  CA_CUDA(cudaStreamSynchronize(st));
  if (flags[0]) fail(CA_ERR_INVALID, "ca_model_create: an N entry has an index outside 1:%lld", (long long)m64);
  if (flags[1] || flags[2]) return false;
  if (noff > cap) return false;  // some block has more than kMaxBlockRows rows: host build
  std::vector<int2> edges((size_t)noff);
  if (noff) CA_CUDA(cudaMemcpyAsync(edges.data(), d_list.ptr, (size_t)noff * sizeof(int2), cudaMemcpyDeviceToHost, st));
  CA_CUDA(cudaStreamSynchronize(st));
  std::vector<int> comp(m);
  std::iota(comp.begin(), comp.end(), 0);
  auto find = [&](int a) {
    while (comp[a] != a) a = comp[a] = comp[comp[a]];
    return a;
  };
  for (const int2& e : edges) {  // either direction links the two modes; the root of a component is its smallest mode
    const int a = find(e.x), b = find(e.y);
    if (a != b) comp[std::max(a, b)] = std::min(a, b);
  }
  std::map<int, std::vector<int>> blocks;
  for (int i = 0; i < m; ++i) blocks[find(i)].push_back(i);
  const int nb = (int)blocks.size();
  std::vector<int32_t> h_rows, h_row_off{0}, h_blk_of(m), h_pos_in(m);
  std::vector<int64_t> h_binv_off{0};
  int gmax = 0;
  {
    int b = 0;
    for (auto& kv : blocks) {
      const std::vector<int>& G = kv.second;
      const int g = (int)G.size();
      gmax = std::max(gmax, g);
      if (g > kMaxBlockRows) return false;
      for (int a = 0; a < g; ++a) {
        h_rows.push_back(G[a]);
        h_blk_of[G[a]] = b;
        h_pos_in[G[a]] = a;
      }
      h_row_off.push_back((int32_t)h_rows.size());
      h_binv_off.push_back(h_binv_off.back() + (int64_t)g * g);
      ++b;
    }
  }
  DeviceBuffer<int32_t> d_rows, d_row_off, d_blk_of, d_pos_in;
  DeviceBuffer<double> d_binv;
  DeviceBuffer<int64_t> d_binv_off;
  d_rows.upload(h_rows.data(), h_rows.size(), st);
  d_row_off.upload(h_row_off.data(), h_row_off.size(), st);
  d_binv_off.upload(h_binv_off.data(), h_binv_off.size(), st);
  d_binv.alloc((size_t)std::max<int64_t>(h_binv_off.back(), 1));
  b_gather_blocks_kernel<<<(unsigned)nb, 256, 0, st>>>(dB, m, d_rows.ptr, d_row_off.ptr, d_binv_off.ptr, nb, d_binv.ptr);
  CA_CUDA(cudaGetLastError());
  count_launch();
  std::vector<double> h_binv((size_t)h_binv_off.back());
  if (!h_binv.empty())
    CA_CUDA(cudaMemcpyAsync(h_binv.data(), d_binv.ptr, h_binv.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
  CA_CUDA(cudaStreamSynchronize(st));
  for (int b = 0; b < nb; ++b) {  // invert in place, block by block (row-major g x g)
    const int g = h_row_off[(size_t)b + 1] - h_row_off[(size_t)b];
    std::vector<double> Bg(h_binv.begin() + h_binv_off[(size_t)b], h_binv.begin() + h_binv_off[(size_t)b + 1]), Binv;
    if (!invert_block_rm(Bg, g, Binv)) fail(CA_ERR_INVALID, "ca_model_create: B is singular");
    std::copy(Binv.begin(), Binv.end(), h_binv.begin() + h_binv_off[(size_t)b]);
  }
  const int64_t uwords = ((int64_t)m * m + 31) / 32, swords = ((int64_t)nin * nin + 31) / 32;
  if ((int64_t)nb * swords * 4 > kMaxBitmapBytes) return false;
  d_blk_of.upload(h_blk_of.data(), h_blk_of.size(), st);
  d_pos_in.upload(h_pos_in.data(), h_pos_in.size(), st);
  if (!h_binv.empty())
    CA_CUDA(cudaMemcpyAsync(d_binv.ptr, h_binv.data(), h_binv.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  CA_CUDA(cudaStreamSynchronize(st));  // h_binv is a pageable local
  const BlockMap bm{d_blk_of.ptr, d_pos_in.ptr, d_row_off.ptr, d_rows.ptr, d_binv.ptr, d_binv_off.ptr};
  h->nblocks = nb;
  lap(S.blocks, t0);
  tm.lap("blocks of B");

  // ---- linear operators
  const int64_t mm = (int64_t)m * m;
  h->L1.alloc((size_t)mm);
  h->L2.alloc((size_t)mm);
  h->L1r.alloc((size_t)mm);
  h->L2r.alloc((size_t)mm);
  h->dA1.alloc((size_t)mm);
  h->dA2.alloc((size_t)mm);
  CA_CUDA(cudaMemcpyAsync(h->dA1.ptr, dA1, (size_t)mm * sizeof(double), cudaMemcpyDeviceToDevice, st));
  CA_CUDA(cudaMemcpyAsync(h->dA2.ptr, dA2, (size_t)mm * sizeof(double), cudaMemcpyDeviceToDevice, st));
  fold_linear_kernel<<<(unsigned)((mm + 255) / 256), 256, 0, st>>>(dA1, dA2, m, bm, h->L1.ptr, h->L2.ptr, h->L1r.ptr, h->L2r.ptr);
  CA_CUDA(cudaGetLastError());
  count_launch();
  // compact copy of N as given (the time stepper packs it on first use)
  h->dncoo.alloc((size_t)nnz);
  if (nnz > 0) {
    copy_coo_kernel<<<grid_for(nnz), 256, 0, st>>>(di, dj, dk, dval, nnz, h->dncoo.ptr);
    CA_CUDA(cudaGetLastError());
    count_launch();
  }
  h->ncoo_on_device = true;
  h->linear_on_device = true;

  // ---- ordered pair sets of the blocks, D, R = Binv D
  ScanScratch scan_tmp;
  FoldedDevice& F = h->folded;
  DeviceBuffer<uint32_t> ubits;
  DeviceBuffer<int64_t> uprefix;
  const int64_t nuw = (int64_t)nb * uwords;
  ubits.alloc((size_t)nuw + 1);
  uprefix.alloc((size_t)nuw + 1);
  CA_CUDA(cudaMemsetAsync(ubits.ptr, 0, ((size_t)nuw + 1) * sizeof(uint32_t), st));
  if (nnz > 0) {
    mark_ordered_kernel<<<grid_for(nnz), 256, 0, st>>>(di, dj, dk, nnz, m64, bm, uwords, ubits.ptr);
    CA_CUDA(cudaGetLastError());
    count_launch();
  }
  exclusive_scan(reinterpret_cast<const PopcWord*>(ubits.ptr), uprefix.ptr, nuw + 1, scan_tmp, st);
  // per-block list offsets to the host
  std::vector<int64_t> u_off((size_t)nb + 1);
  CA_CUDA(cudaMemcpy2DAsync(u_off.data(), sizeof(int64_t), uprefix.ptr, (size_t)uwords * sizeof(int64_t), sizeof(int64_t),
                            (size_t)nb + 1, cudaMemcpyDeviceToHost, st));
  CA_CUDA(cudaStreamSynchronize(st));
  const int64_t total_u = u_off[(size_t)nb];
  std::vector<int64_t> udense_off((size_t)nb + 1, 0);
  int64_t max_nu = 0;
  for (int b = 0; b < nb; ++b) {
    const int64_t nu = u_off[(size_t)b + 1] - u_off[b];
    max_nu = std::max(max_nu, nu);
    udense_off[(size_t)b + 1] = udense_off[b] + nu * (h_row_off[(size_t)b + 1] - h_row_off[b]);
  }
  F.total_u = total_u;
  F.u_off = u_off;
  F.dense_off = udense_off;
  F.row_off = h_row_off;
  F.rows = h_rows;
  F.ukeys.alloc((size_t)std::max<int64_t>(total_u, 1));
  expand_bits_kernel<<<grid_for(nuw), 256, 0, st>>>(ubits.ptr, uprefix.ptr, nuw, uwords, F.ukeys.ptr);
  CA_CUDA(cudaGetLastError());
  count_launch();
  DeviceBuffer<int64_t> d_udense_off;
  d_udense_off.upload(udense_off.data(), udense_off.size(), st);
  DeviceBuffer<double> D;
  const int64_t ndense = udense_off[(size_t)nb];
  D.alloc((size_t)std::max<int64_t>(ndense, 1));
  F.R.alloc((size_t)std::max<int64_t>(ndense, 1));
  CA_CUDA(cudaMemsetAsync(D.ptr, 0, (size_t)std::max<int64_t>(ndense, 1) * sizeof(double), st));
  if (nnz > 0) {
    scatter_dense_kernel<<<grid_for(nnz), 256, 0, st>>>(di, dj, dk, dval, nnz, m64, bm, uwords, ubits.ptr, uprefix.ptr,
                                                         d_udense_off.ptr, D.ptr);
    CA_CUDA(cudaGetLastError());
    count_launch();
  }
  if (max_nu > 0) {
    const dim3 grid((unsigned)((max_nu + 127) / 128), (unsigned)nb);
    fold_quadratic_kernel<<<grid, dim3(128, 4), (size_t)gmax * gmax * sizeof(double), st>>>(D.ptr, F.R.ptr, bm, uwords,
                                                                                         uprefix.ptr, d_udense_off.ptr);
    CA_CUDA(cudaGetLastError());
    count_launch();
  }
  h->qterms_on_device = true;
  lap(S.fold, t0);
  D.release();
  tm.lap("fold B^-1");

  // ---- symmetrise: pair sets min * nin + max of the blocks plus the linear placeholders, values, CSR
  DeviceBuffer<uint32_t> sbits, skeys;
  DeviceBuffer<int64_t> sprefix;
  const int64_t nsw = (int64_t)nb * swords;
  sbits.alloc((size_t)nsw + 1);
  sprefix.alloc((size_t)nsw + 1);
  CA_CUDA(cudaMemsetAsync(sbits.ptr, 0, ((size_t)nsw + 1) * sizeof(uint32_t), st));
  if (total_u > 0) {
    mark_sym_kernel<<<grid_for(total_u), 256, 0, st>>>(F.ukeys.ptr, uprefix.ptr, uwords, nb, total_u, m, nin, swords, sbits.ptr);
    CA_CUDA(cudaGetLastError());
    count_launch();
  }
  mark_linear_kernel<<<(unsigned)((mm + 255) / 256), 256, 0, st>>>(h->L1.ptr, h->L2.ptr, m, nin, bm, swords, sbits.ptr);
  CA_CUDA(cudaGetLastError());
  count_launch();
  exclusive_scan(reinterpret_cast<const PopcWord*>(sbits.ptr), sprefix.ptr, nsw + 1, scan_tmp, st);
  std::vector<int64_t> s_off((size_t)nb + 1);
  CA_CUDA(cudaMemcpy2DAsync(s_off.data(), sizeof(int64_t), sprefix.ptr, (size_t)swords * sizeof(int64_t), sizeof(int64_t),
                            (size_t)nb + 1, cudaMemcpyDeviceToHost, st));
  CA_CUDA(cudaStreamSynchronize(st));
  const int64_t total_s = s_off[(size_t)nb];
  std::vector<int64_t> sdense_off((size_t)nb + 1, 0);
  int64_t max_ns = 0;
  for (int b = 0; b < nb; ++b) {
    const int64_t ns = s_off[(size_t)b + 1] - s_off[b];
    max_ns = std::max(max_ns, ns);
    sdense_off[(size_t)b + 1] = sdense_off[b] + ns * (h_row_off[(size_t)b + 1] - h_row_off[b]);
  }
  skeys.alloc((size_t)std::max<int64_t>(total_s, 1));
  expand_bits_kernel<<<grid_for(nsw), 256, 0, st>>>(sbits.ptr, sprefix.ptr, nsw, swords, skeys.ptr);
  CA_CUDA(cudaGetLastError());
  count_launch();
  DeviceBuffer<int64_t> d_sdense_off, vpos;
  d_sdense_off.upload(sdense_off.data(), sdense_off.size(), st);
  const int64_t nsd = sdense_off[(size_t)nb];
  DeviceBuffer<double> Vs;
  Vs.alloc((size_t)nsd + 1);
  vpos.alloc((size_t)nsd + 1);
  CA_CUDA(cudaMemsetAsync(Vs.ptr + nsd, 0, sizeof(double), st));
  DeviceBuffer<int64_t> row_begin, row_end;
  row_begin.alloc((size_t)m);
  row_end.alloc((size_t)m);
  CA_CUDA(cudaMemsetAsync(row_begin.ptr, 0, (size_t)m * sizeof(int64_t), st));
  CA_CUDA(cudaMemsetAsync(row_end.ptr, 0, (size_t)m * sizeof(int64_t), st));
  if (max_ns > 0) {
    const dim3 grid((unsigned)((max_ns + 255) / 256), (unsigned)nb);
    sym_values_kernel<<<grid, 256, 0, st>>>(F.R.ptr, ubits.ptr, uprefix.ptr, uwords, d_udense_off.ptr, skeys.ptr, sprefix.ptr,
                                            swords, d_sdense_off.ptr, h->L1.ptr, h->L2.ptr, m, nin, bm, Vs.ptr);
    CA_CUDA(cudaGetLastError());
    count_launch();
  }
  exclusive_scan(reinterpret_cast<const NonZero*>(Vs.ptr), vpos.ptr, nsd + 1, scan_tmp, st);
  int64_t nterms = 0;
  CA_CUDA(cudaMemcpyAsync(&nterms, vpos.ptr + nsd, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CA_CUDA(cudaStreamSynchronize(st));
  DeviceBuffer<uint32_t> keys;
  DeviceBuffer<double> vals;
  keys.alloc((size_t)std::max<int64_t>(nterms, 1));
  vals.alloc((size_t)std::max<int64_t>(nterms, 1));
  if (max_ns > 0) {
    const dim3 grid((unsigned)((max_ns + 255) / 256), (unsigned)nb);
    compact_rows_kernel<<<grid, 256, 0, st>>>(Vs.ptr, vpos.ptr, skeys.ptr, sprefix.ptr, swords, d_sdense_off.ptr, bm, keys.ptr,
                                              vals.ptr, row_begin.ptr, row_end.ptr);
    CA_CUDA(cudaGetLastError());
    count_launch();
  }
  lap(S.symmetrise, t0);
  Vs.release();
  vpos.release();
  ubits.release();
  uprefix.release();
  sbits.release();
  sprefix.release();
  skeys.release();
  tm.lap("symmetrise + CSR");

  // ---- packer: sketches, clustering (host decisions, device set operations)
  const int window_modes = window_modes_for(m);
  int nt_max = nt_max_for(window_modes);
  DeviceRowSets sets;
  sets.m = m;
  sets.words = swords;
  sets.keys = keys.ptr;
  sets.row_begin = row_begin.ptr;
  sets.row_end = row_end.ptr;
  sets.st = st;
  sets.hbegin.resize((size_t)m);
  sets.hend.resize((size_t)m);
  sets.hsketch.resize((size_t)m * kPackSketch);
  {
    DeviceBuffer<uint64_t> dsketch;
    dsketch.alloc((size_t)m * kPackSketch);
    sketch_kernel<<<(unsigned)m, 256, 0, st>>>(keys.ptr, row_begin.ptr, row_end.ptr, dsketch.ptr);
    CA_CUDA(cudaGetLastError());
    count_launch();
    CA_CUDA(cudaMemcpyAsync(sets.hsketch.data(), dsketch.ptr, sets.hsketch.size() * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CA_CUDA(cudaMemcpyAsync(sets.hbegin.data(), row_begin.ptr, (size_t)m * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CA_CUDA(cudaMemcpyAsync(sets.hend.data(), row_end.ptr, (size_t)m * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CA_CUDA(cudaStreamSynchronize(st));
  }
  sets.bits.alloc((size_t)swords);
  std::vector<std::vector<int32_t>> clusters;
  std::vector<int32_t> empties;
  cluster_rows(sets, clusters, empties);
  nt_max = decide_nt_max(clusters, m, nt_max);
  const std::vector<std::vector<int32_t>> tile_rows = split_clusters(clusters, nt_max);
  const int ntiles = (int)tile_rows.size();
  lap(S.cluster, t0);
  tm.lap("sketches + clustering");

  // ---- per-tile pair unions -> host layout (pair sequence, windows) -> value placement
  if ((int64_t)ntiles * swords * 4 > kMaxBitmapBytes) fail(CA_ERR_NOMEM, "ca_model_create_dev: tile bitmaps exceed the budget");
  std::vector<int32_t> h_tile_of((size_t)m, -1), h_slot_of((size_t)m, 0);
  for (int t = 0; t < ntiles; ++t)
    for (size_t n = 0; n < tile_rows[t].size(); ++n) {
      h_tile_of[(size_t)tile_rows[t][n]] = t;
      h_slot_of[(size_t)tile_rows[t][n]] = (int32_t)n;
    }
  DeviceBuffer<int32_t> d_tile_of, d_slot_of;
  d_tile_of.upload(h_tile_of.data(), h_tile_of.size(), st);
  d_slot_of.upload(h_slot_of.data(), h_slot_of.size(), st);
  DeviceBuffer<uint32_t> tbits, tkeys;
  DeviceBuffer<int64_t> tprefix;
  const int64_t ntw = (int64_t)ntiles * swords;
  tbits.alloc((size_t)ntw + 1);
  tprefix.alloc((size_t)ntw + 1);
  CA_CUDA(cudaMemsetAsync(tbits.ptr, 0, ((size_t)ntw + 1) * sizeof(uint32_t), st));
  mark_tiles_kernel<<<dim3(8, (unsigned)m), 256, 0, st>>>(keys.ptr, row_begin.ptr, row_end.ptr, d_tile_of.ptr, swords, tbits.ptr);
  CA_CUDA(cudaGetLastError());
  count_launch();
  exclusive_scan(reinterpret_cast<const PopcWord*>(tbits.ptr), tprefix.ptr, ntw + 1, scan_tmp, st);
  std::vector<int64_t> t_off((size_t)ntiles + 1, 0);
  if (ntiles > 0)
    CA_CUDA(cudaMemcpy2DAsync(t_off.data(), sizeof(int64_t), tprefix.ptr, (size_t)swords * sizeof(int64_t), sizeof(int64_t),
                              (size_t)ntiles + 1, cudaMemcpyDeviceToHost, st));
  CA_CUDA(cudaStreamSynchronize(st));
  const int64_t total_t = t_off[(size_t)ntiles];
  tkeys.alloc((size_t)std::max<int64_t>(total_t, 1));
  if (ntw > 0) {
    expand_bits_kernel<<<grid_for(ntw), 256, 0, st>>>(tbits.ptr, tprefix.ptr, ntw, swords, tkeys.ptr);
    CA_CUDA(cudaGetLastError());
    count_launch();
  }
  DeviceBuffer<unsigned long long> dcounters;
  dcounters.alloc(2);
  CA_CUDA(cudaMemsetAsync(dcounters.ptr, 0, 2 * sizeof(unsigned long long), st));
  if (ntiles > 0) {
    distinct_pairs_kernel<<<grid_for(swords), 256, 0, st>>>(tbits.ptr, ntiles, swords, dcounters.ptr);
    CA_CUDA(cudaGetLastError());
    count_launch();
  }
  std::vector<uint32_t> h_tkeys((size_t)total_t);
  if (total_t > 0)
    CA_CUDA(cudaMemcpyAsync(h_tkeys.data(), tkeys.ptr, (size_t)total_t * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  unsigned long long h_distinct = 0;
  CA_CUDA(cudaMemcpyAsync(&h_distinct, dcounters.ptr, sizeof h_distinct, cudaMemcpyDeviceToHost, st));
  CA_CUDA(cudaStreamSynchronize(st));
  tkeys.release();
  lap(S.tile_unions, t0);
  tm.lap("tile unions");

  PackedHost P;
  P.m = m;
  P.nin = nin;
  P.symmetric = true;
  P.window_modes = window_modes;
  P.nterms = nterms;
  P.npairs_distinct = (int64_t)h_distinct;
  std::vector<PackedHost> parts((size_t)ntiles);
  std::vector<int32_t> h_where((size_t)total_t);
  {
    const int order = pack_pair_order(window_modes);
    P.pair_order = order;
    parallel_chunks((size_t)ntiles, (size_t)ntiles, [&](size_t b0, size_t b1, size_t) {
      std::vector<uint64_t> U;
      std::vector<int32_t> where;
      for (size_t t = b0; t < b1; ++t) {
        const int64_t n = t_off[t + 1] - t_off[t];
        U.resize((size_t)n);
        for (int64_t q = 0; q < n; ++q) U[(size_t)q] = h_tkeys[(size_t)(t_off[t] + q)];
        layout_tile(tile_rows[t], U.data(), U.size(), nin, window_modes, order, parts[t], where);
        std::copy(where.begin(), where.end(), h_where.begin() + (ptrdiff_t)t_off[t]);
      }
    });
  }
  // emission-order rows offset of every tile identifies it after finalize_pack has sorted the descriptors
  std::vector<int32_t> rows_base((size_t)ntiles, 0);
  {
    int32_t at = 0;
    for (int t = 0; t < ntiles; ++t) {
      rows_base[(size_t)t] = at;
      at += (int32_t)parts[(size_t)t].rows.size();
    }
  }
  P = finalize_pack(std::move(P), parts, empties, nt_max);
  std::vector<int64_t> h_tile_voff((size_t)ntiles, 0);
  std::vector<int32_t> h_tile_nt((size_t)ntiles, 1);
  {
    std::map<int32_t, const TileDesc*> by_rows;
    for (const TileDesc& T : P.tiles) by_rows[T.rows_off] = &T;
    for (int t = 0; t < ntiles; ++t) {
      const TileDesc* T = by_rows.at(rows_base[(size_t)t]);
      h_tile_voff[(size_t)t] = T->vals_off;
      h_tile_nt[(size_t)t] = T->nt;
    }
  }
  lap(S.layout, t0);
  tm.lap("layout (host)");

  DeviceBuffer<int32_t> d_where, d_tile_nt;
  DeviceBuffer<int64_t> d_tile_voff;
  d_where.upload(h_where.data(), h_where.size(), st);
  d_tile_nt.upload(h_tile_nt.data(), h_tile_nt.size(), st);
  d_tile_voff.upload(h_tile_voff.data(), h_tile_voff.size(), st);
  h->op.upload_meta(P);
  h->op.vals.alloc((size_t)std::max<int64_t>(P.vals_count, 1));
  CA_CUDA(cudaMemsetAsync(h->op.vals.ptr, 0, (size_t)std::max<int64_t>(P.vals_count, 1) * sizeof(double), st));
  h->lin_pos.alloc((size_t)mm);
  h->lin_l1.alloc((size_t)mm);
  h->lin_l2.alloc((size_t)mm);
  if (ntiles > 0) {
    place_values_kernel<<<dim3(8, (unsigned)m), 256, 0, st>>>(keys.ptr, vals.ptr, row_begin.ptr, row_end.ptr, d_tile_of.ptr,
                                                             d_slot_of.ptr, d_tile_voff.ptr, d_tile_nt.ptr, swords, tbits.ptr,
                                                             tprefix.ptr, d_where.ptr, m, nin, h->L1.ptr, h->L2.ptr,
                                                             h->op.vals.ptr, dcounters.ptr + 1, h->lin_pos.ptr,
                                                             h->lin_l1.ptr, h->lin_l2.ptr);
    CA_CUDA(cudaGetLastError());
    count_launch();
  }
  unsigned long long h_nlin = 0;
  CA_CUDA(cudaMemcpyAsync(&h_nlin, dcounters.ptr + 1, sizeof h_nlin, cudaMemcpyDeviceToHost, st));
  CA_CUDA(cudaStreamSynchronize(st));
  h->nnz_lin = (int64_t)h_nlin;
  lap(S.place, t0);
  tm.lap("value placement");

  // ---- few-state streaming form of Q straight from the CSR (row-chunked; stream.cu)
  {
    DeviceBuffer<int32_t> qflag;
    DeviceBuffer<int64_t> qpos;
    qflag.alloc((size_t)nterms + 1);
    qpos.alloc((size_t)nterms + 1);
    CA_CUDA(cudaMemsetAsync(qflag.ptr + nterms, 0, sizeof(int32_t), st));
    if (nterms > 0) {
      quad_flags_kernel<<<grid_for(nterms), 256, 0, st>>>(keys.ptr, nterms, nin, m, qflag.ptr);
      CA_CUDA(cudaGetLastError());
      count_launch();
    }
    exclusive_scan(qflag.ptr, qpos.ptr, nterms + 1, scan_tmp, st);
    int64_t nq = 0;
    CA_CUDA(cudaMemcpyAsync(&nq, qpos.ptr + nterms, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    std::vector<int64_t> qb((size_t)m), qe((size_t)m);
    DeviceBuffer<int64_t> dq;
    dq.alloc((size_t)2 * m);
    gather_ranges_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(qpos.ptr, row_begin.ptr, row_end.ptr, m, dq.ptr);
    CA_CUDA(cudaGetLastError());
    count_launch();
    CA_CUDA(cudaMemcpyAsync(qb.data(), dq.ptr, (size_t)m * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CA_CUDA(cudaMemcpyAsync(qe.data(), dq.ptr + m, (size_t)m * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CA_CUDA(cudaStreamSynchronize(st));
    DeviceBuffer<uint32_t> ab;
    DeviceBuffer<double> qv;
    ab.alloc((size_t)std::max<int64_t>(nq, 1));
    qv.alloc((size_t)std::max<int64_t>(nq, 1));
    if (nterms > 0) {
      stream_fill_kernel<<<grid_for(nterms), 256, 0, st>>>(keys.ptr, vals.ptr, nterms, nin, m, qpos.ptr, ab.ptr, qv.ptr);
      CA_CUDA(cudaGetLastError());
      count_launch();
    }
    CA_CUDA(cudaStreamSynchronize(st));
    h->stream.adopt(std::move(ab), std::move(qv), qb, qe, m, /*symmetric=*/true);
    h->stream_built = true;
  }
  lap(S.stream, t0);
  tm.lap("streaming form");

  // small models: the host copies that the generated kernel, the Df structure and the in-kernel solver need
  if (m <= 96) {
    h->host_qterms();
    h->host_linear();
  }
  CA_CUDA(cudaStreamSynchronize(st));
  S.total = now_seconds() - t_start;
  S.on_device = 1;
  S.nclusters = (int64_t)clusters.size();
  S.ntiles = (int64_t)P.tiles.size();
  S.nterms = nterms;
  return true;
}
