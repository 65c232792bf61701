// Host-side packing of a sparse bilinear (+ linear) operator into DMMA row tiles.
//
// The operator  out_i = sum_{(a,b)} V[i,(a,b)] * u_a * w_b  is re-expressed as a block-sparse matrix
// product over PAIR PRODUCTS: out[s, i] = sum_p P[s, p] * V[p, i] with P[s,(a,b)] = u_a[s] * w_b[s].
// Output rows with (nearly) identical pair sets -- in a Galerkin model: modes of one wavenumber group
// and wall-normal parity -- are clustered into tiles of 8*nt rows that share one pair list, so that
// inside a tile V is dense and feeds mma.sync.m8n8k4.f64 directly:
//   A fragment = pair products (8 states x 4 pairs), generated on the fly from the state tile in
//   shared memory;  B fragment = V (4 pairs x 8 rows), streamed from L2 in fragment order.
// A linear term L[i,j] x_j is the pair (j, ONE) with the pseudo mode ONE == m holding 1.0.
#pragma once
#include <cstdint>
#include <vector>

namespace ca {

struct Term {
  int32_t i, a, b;  // output row, first and second factor (0-based; a or b == m means the constant 1)
  double v;
  __host__ __device__ Term() {}  // deliberately empty: std::vector<Term>(n) must not zero-fill gigabytes that are overwritten right away
  Term(int32_t i_, int32_t a_, int32_t b_, double v_) : i(i_), a(a_), b(b_), v(v_) {}
};

struct TileDesc {
  int32_t ks_begin;  // first k-step of this tile (one k-step = 4 pairs) in `pairs`
  int32_t nks;       // number of k-steps
  int64_t vals_off;  // offset (doubles) of this tile's values: [nks][nt][32 lanes]
  int32_t rows_off;  // offset into `rows` (8*nt entries, -1 = padding)
  int32_t nt;        // n-tiles of 8 output rows
  int32_t nrows;     // real rows in the tile
  int32_t npairs;    // real pairs in the tile (<= 4*nks)
};

// A window is a contiguous range of a tile's k-steps whose pairs touch at most `window_modes` distinct
// modes; the windowed kernel gathers exactly those rows of the state tile into shared memory (double
// buffered), so its footprint no longer grows with m.  Pair words of a windowed pack hold LOCAL row
// indices into the window's mode list.
struct alignas(16) WindowDesc {
  int32_t ks_begin;   // first k-step (global index, as TileDesc::ks_begin)
  int32_t nks;        // <= kWindowMaxKs
  int32_t modes_off;  // offset into PackedHost::modes
  int32_t nmodes;
  int32_t tile;       // owning tile
  int32_t last;       // 1 if this is the last window of its tile
  int32_t rows_off;   // the owning tile's TileDesc::rows_off and nt, repeated here so that the kernel
  int32_t nt;         // needs no dependent second descriptor load
  int64_t vals_off;   // offset (doubles) of this window's first k-step in `vals`
  int64_t pad_;
};
static_assert(sizeof(WindowDesc) == 48, "the windowed kernel copies descriptors as three 16-byte words");
constexpr int kStreamAlign = 8;    // = kPairBatch of the whole-tile kernel (bilinear_kernels.cuh)
constexpr int kWindowMaxKs = 448;  // k-steps per window (the kernel stages a window's pair words, 4 B per pair, in shared memory)

struct PackedHost {
  int32_t m = 0;    // outputs
  int32_t nin = 0;  // inputs including the constant-one pseudo mode (index nin-1)
  bool symmetric = false;
  std::vector<TileDesc> tiles;
  std::vector<uint32_t> pairs;  // a | (b << 16), 4 per k-step
  std::vector<double> vals;
  std::vector<int32_t> rows;
  int32_t window_modes = 0;     // 0: not windowed (pair words hold global mode indices)
  std::vector<WindowDesc> windows;  // in tile order
  std::vector<int32_t> modes;       // global mode index of every local row of every window
  int32_t max_window_modes = 0;
  // Whole-tile-per-warp schedule (longest-processing-time assignment of tiles to the 8 warps of a CTA): warp w
  // runs tiles sched[sched_off[w] .. sched_off[w+1]), the first sched_nt2[w] of them have nt == 2.  With it a
  // warp owns all k-steps of its tiles: no cross-warp reduction, no per-tile barrier.  sched_imbalance =
  // heaviest warp / mean; the kernel falls back to splitting k-steps across warps when it is poor.
  std::vector<int32_t> sched;
  int32_t sched_off[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  int32_t sched_nt2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  double sched_imbalance = 0.0;
  // When the pack is not windowed the tiles of a warp are laid out contiguously in schedule order, so that a warp's
  // pairs and values form one stream per nt class that can be prefetched across tiles; in the pair stream every tile
  // starts on a multiple of kStreamAlign k-steps (sched_ks2 / sched_ks1: stream lengths in k-steps, padding included).
  int32_t sched_ks2[8] = {0, 0, 0, 0, 0, 0, 0, 0}, sched_ks1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  bool stream_layout = false;
  int64_t nterms = 0;           // merged non-zero terms
  int64_t npairs_distinct = 0;  // distinct (a,b) over the whole operator
  int64_t padded_fma = 0;       // FMAs the tiles execute per state (incl. padding)
  int64_t pair_products = 0;    // pair products formed per state (sum over tiles)
  int64_t blocked_pairs = 0;    // of these, in (4 first factors) x (one second factor) k-steps
  int32_t pair_order = 0;       // kOrderAMajor / kOrderBlocked / kOrderFactored (pack_pair_order)
  int64_t vals_count = 0;       // value slots ([k-step][n-tile][lane]); == vals.size() when the values are held here
};

// ---- staged interface (the device model build, model_build.cu, runs the bulk data movement of these stages on
// the GPU and reuses the decision logic below verbatim, so that both builds produce the same pack bit for bit)

constexpr int kPackSketch = 8;  // min-hash values per row
// hash h (0..kPackSketch-1) of pair id key = a * nin + b; a row's sketch is the minimum over its keys
inline uint64_t pack_sketch_hash_host(uint64_t key, int h) {
  uint64_t x = key * 0x100000001b3ull + (uint64_t)h * 0x51ed27ull;
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}

// The packer's view of the rows' pair sets during clustering: sizes, sketches and unions against one open set.
struct RowSetOps {
  virtual ~RowSetOps() {}
  virtual int32_t nrows() const = 0;
  virtual int64_t nkeys(int32_t r) const = 0;
  virtual const uint64_t* sketch(int32_t r) const = 0;  // kPackSketch values
  virtual int batch_hint() const = 0;                    // how many union sizes to ask for at once
  virtual void open(int32_t seed) = 0;                   // open set U := keys(seed)
  virtual int64_t open_size() const = 0;
  virtual void union_sizes(const int32_t* cand, int n, int64_t* out) = 0;  // |U + keys(cand[q])|
  virtual void absorb(int32_t r) = 0;                    // U := U + keys(r)
};
// Greedy clustering of rows with near-identical pair sets (seeds by descending size; candidates agreeing in at least half
// of the sketch; accepted while the union stays within 1.2 x the larger set).  Clusters hold ascending row indices.
void cluster_rows(RowSetOps& sets, std::vector<std::vector<int32_t>>& clusters, std::vector<int32_t>& empties);
// The rule for tiles of three n-tiles (windowed packs): only where clusters of 17-24 rows hold >= 70 % of the rows.
int decide_nt_max(const std::vector<std::vector<int32_t>>& clusters, int32_t m, int nt_max);
// Cuts clusters into tiles of <= 8 * nt_max rows (even split).
std::vector<std::vector<int32_t>> split_clusters(const std::vector<std::vector<int32_t>>& clusters, int nt_max);
// Pair sequence of one tile: given the sorted union U of the tile's pair ids (a * nin + b), emits into P (which must be
// empty) the tile descriptor, its rows, pair words, windows and window mode lists -- everything but the values -- and
// where[q] = position of U[q] in the emitted (padded) pair sequence.  P.vals_count is the number of value slots.
void layout_tile(const std::vector<int32_t>& trows, const uint64_t* U, size_t nU, int32_t nin, int window_modes,
                 int order, PackedHost& P, std::vector<int32_t>& where);
enum : int { kOrderAMajor = 0, kOrderBlocked = 1, kOrderFactored = 2 };
int pack_pair_order(int window_modes);
// Appends the per-tile parts (in order), adds the empty-row tiles, sorts tiles and windows, builds the warp schedule and
// (whole-tile packs) the schedule-order layout.  When the parts carry values they are moved along; otherwise only the
// offsets are computed (vals_count / TileDesc::vals_off) and the caller places the values itself.
PackedHost finalize_pack(PackedHost P, std::vector<PackedHost>& parts, const std::vector<int32_t>& empties, int nt_max);

// FNV-1a digests of the arrays of a pack, in the order: tiles, pairs, vals, rows, windows, modes, sched (+ schedule
// scalars), counters.  Two packs are the same pack iff all eight agree (diagnostics: host packer vs device build).
void pack_digest(const PackedHost& P, uint64_t out[8]);
uint64_t fnv1a(const void* data, size_t bytes, uint64_t h = 0xcbf29ce484222325ull);

// In-place sort by (i, a, b): counting sort by output row, then per-row sorts on worker threads.
void sort_terms(std::vector<Term>& terms, int32_t m);

// Sorted by (i,a,b), duplicates summed, exact zeros dropped; symmetric: (a,b) canonicalised to a <= b first.
std::vector<Term> merge_terms(std::vector<Term> terms, bool symmetric, int32_t m = -1);

// Merges duplicate (i,a,b), drops exact zeros, clusters rows, lays out fragments.
// symmetric: (a,b) and (b,a) are the same pair (u == w); ordered otherwise.
// n_inputs: number of input modes (default: m, a square operator); the constant-one pseudo mode is n_inputs.
PackedHost pack_terms(std::vector<Term> terms, int32_t m, bool symmetric, int nt_max, int window_modes = 0,
                      int32_t n_inputs = -1);

}  // namespace ca
