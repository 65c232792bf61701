// Device-resident packed operators and their launchers.
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <vector>

#include "common.h"
#include "pack.h"

namespace ca {

// A PackedHost uploaded to the current device.
struct DevicePack {
  int32_t m = 0, nin = 0, ntiles = 0, nt2_tiles = 0;  // tiles are ordered nt == 2 first
  bool symmetric = false;
  int64_t nterms = 0, npairs_distinct = 0, padded_fma = 0, pair_products = 0;
  DeviceBuffer<TileDesc> tiles;
  std::vector<uint32_t> host_pairs;      // a | flag << 15 | b << 16 (pack.h), kept to derive the per-MT form
  mutable DeviceBuffer<uint2> pairs_mt[3];  // address-word form for MT = 1, 2, 4 (built on first use)
  const uint2* pairs_for(int mt) const;
  mutable DeviceBuffer<uint32_t> pairs_win;  // packed one-word form of windowed packs (built on first use)
  const uint32_t* pairs_windowed() const;
  DeviceBuffer<double> vals;
  DeviceBuffer<int32_t> rows;
  // whole-tile-per-warp schedule (pack.h)
  DeviceBuffer<int32_t> sched;
  int32_t sched_off[9] = {0}, sched_nt2[8] = {0}, sched_ks2[8] = {0}, sched_ks1[8] = {0};
  bool stream_layout = false;
  int32_t pair_order = 0;  // pack.h: kOrderFactored lets the quadratic evaluation run without pair products
  double sched_imbalance = 1e9;
  // windowed packs (large m): pair words hold window-local rows
  int32_t window_modes = 0, nwin = 0, nwin_nt3 = 0, nwin_nt2 = 0, max_window_modes = 0;
  DeviceBuffer<WindowDesc> windows;
  DeviceBuffer<int32_t> modes;
  // tile boundaries in the window list, with the DMMAs before them: small ensembles split the windows among CTAs there
  std::vector<int32_t> tile_first_window;  // ntiles_with_windows + 1 entries, the last one = nwin
  std::vector<int64_t> dmma_before_tile;   // same length: cumulative nks * nt
  std::vector<int64_t> dmma_before_tile_index;  // whole-tile packs: ntiles + 1 entries in the order of `tiles`
  void upload(const PackedHost& P);
  // everything but the values (device build: the values are placed on the device, DevicePack::vals is filled there)
  void upload_meta(const PackedHost& P);
};

// mode: MODE_QUAD (Y ignored), MODE_ORDERED (u = X, w = Y), MODE_JVP (X = state, Y = direction).
void launch_batched(const DevicePack& op, int mode, const double* dX, const double* dY,
                    int64_t nstates, int64_t ldx, double* dOUT, int64_t ldo, cudaStream_t stream);

// Host ensembles are processed in chunks on kStage streams with their own staging buffers: with three stages the
// H2D of chunk c+1 and the D2H of chunk c-1 both overlap the kernel of chunk c with no bubble (with two, every
// other kernel waits for a D2H + H2D pair on its own stream).
constexpr int kStage = 3;
// launches one chunk of a host ensemble: (dX, nstates (= leading dimension), dOUT, stream)
using LaunchFn = std::function<void(const double*, int64_t, double*, cudaStream_t)>;
void eval_batched_host(const LaunchFn& launch, cudaStream_t streams[kStage], DeviceBuffer<double> in[kStage],
                       DeviceBuffer<double> out[kStage], const double* X, int64_t nstates, double* OUT,
                       int64_t m_in, int64_t m_out, int64_t ld_host = -1 /* leading dimension of X and OUT; default nstates */);

// Packing policy: operators whose two-vector state tile (x and y/v, 32 states) fits in shared memory use
// whole-tile packing; larger ones are packed in windows of kWindowModes modes (128: two window slots per team fit for two teams
// (one vector) or one team (two vectors) per SM, bilinear_kernels.cuh WinCfg).
constexpr int kWindowModes = 128;
int window_modes_for(int32_t m);
// n-tiles (of 8 output rows) a tile may have: 2; windowed packs may take 3 (row clusters of 17-24 rows -- most of a
// 3000-mode model -- then form one tile instead of two half-empty ones); the packer uses that only where such clusters
// hold most of the rows, because the three-n-tile kernel runs one CTA of 8 warps per SM
int nt_max_for(int window_modes);

// true if the DMMA ensemble kernel can hold a state tile of this operator in shared memory
bool batched_fits(const DevicePack& op, int mode);
// which DMMA ensemble kernel launch_batched picks: 0 whole tiles per warp, 3 k-steps split across warps, 4 windowed
int batched_variant(const DevicePack& op, int64_t nstates = 0);
int choose_parts(int64_t nblocks, int64_t slots, int max_parts);
constexpr int64_t kStreamMaxStates = 16;  // at or below this many states the streaming kernel is used

// Row-chunked CSR for the few-state ("streaming") evaluation: the operator is read once from HBM per pass
// of up to 8 states; one warp per row chunk, warp-shuffle reduction, per-row combination of the chunk
// partials in a fixed order (no atomics).  This is the HBM-bound regime of SURVEY 8(d): single-state
// f / N / JVP at m ~ 3000 streams nnz_sym * 12 B.
struct StreamChunk {
  int32_t row;
  int32_t count;
  int64_t begin;
};
struct StreamPack {
  int32_t m = 0, nin = 0;
  bool symmetric = false;
  int64_t nentries = 0, nchunks = 0;
  DeviceBuffer<uint32_t> ab;        // a | b << 16 per entry (16-bit mode indices)
  DeviceBuffer<double> val;
  DeviceBuffer<StreamChunk> chunks;
  DeviceBuffer<int32_t> row_chunk_ptr;  // [m+1]
  std::vector<int32_t> h_row_chunk_ptr;  // host copy: a row range of the operator is a contiguous range of chunks
  // [nchunks][8] chunk partials, one buffer per CUDA stream that evaluates this operator
  mutable std::map<cudaStream_t, std::unique_ptr<DeviceBuffer<double>>> partials;
  double* partial_for(cudaStream_t st) const;
  void build(std::vector<Term> terms, int32_t m, bool symmetric);
  // rectangular operator: m outputs, n_inputs input modes (the constant-one pseudo mode is n_inputs)
  void build_rect(std::vector<Term> terms, int32_t m, int32_t n_inputs);
  // takes device arrays of entries (a | b << 16, value) whose rows are the ranges [row_begin[i], row_end[i]) (host)
  void adopt(DeviceBuffer<uint32_t>&& ab_, DeviceBuffer<double>&& val_, const std::vector<int64_t>& row_begin,
             const std::vector<int64_t>& row_end, int32_t m, bool symmetric);
};
// Optional dense linear part (row-major L1r, L2r, value L1 + L2*invR), may be null.
// row_begin / row_end: only the output rows [row_begin, row_end) are evaluated and written (row_end < 0: all rows) --
// the row-slab sharding of a single-state Jacobian action over several GPUs streams 1/N of the operator per device.
void launch_stream(const StreamPack& sp, int mode, const double* dX, const double* dY, int64_t nstates,
                   int64_t ldx, double* dOUT, int64_t ldo, const double* L1r, const double* L2r, double invR,
                   cudaStream_t stream, int row_begin = 0, int row_end = -1);

// Sliced-ELL structure of the Jacobian  DN[i,j] = sum_k M[(i,j),k] x_k  (SparseBilinear.jl:75-91).
struct JacobianPack {
  int32_t m = 0;
  int64_t nout = 0;      // structurally non-zero Jacobian entries
  int64_t nentries = 0;  // merged (out, src) contributions
  int64_t nslices = 0;
  DeviceBuffer<int64_t> slice_off;  // [nslices+1] offsets (in entries) of each 32-wide slice
  DeviceBuffer<int32_t> out_pos;    // [nslices*32] linear index i + m*j of each lane's output (-1 pad)
  DeviceBuffer<uint16_t> src;       // [entries] column-major inside a slice
  DeviceBuffer<double> val;
  // Output-major copy for ensembles (launch_jacobian_batched; built on its first use from the host mirror below):
  // the contributions of output o = i + m*j are the contiguous range [optr[o], optr[o+1]) of (osrc, oval), so that a
  // warp walking one output reads its list as a sequential, line-reusing, warp-uniform stream.
  mutable DeviceBuffer<int64_t> optr;
  mutable DeviceBuffer<uint16_t> osrc;
  mutable DeviceBuffer<double> oval;
  mutable std::vector<int64_t> h_out;     // host mirror: output index of every merged contribution (sorted)
  mutable std::vector<uint16_t> h_src;
  mutable std::vector<double> h_val;
  void ensure_output_major() const;
  void build(const std::vector<Term>& coo, int32_t m);
};
// Dense Jacobians of an ensemble: OUT[s + ldo * (i + m*j)] = DN(x_s)[i,j] (+ L1[i,j] + L2[i,j]*invR when L1 != null),
// X nstates x m column-major with leading dimension ldx.  State index fastest, like every ensemble of the library.
void launch_jacobian_batched(const JacobianPack& J, const double* dX, int64_t nstates, int64_t ldx, const double* L1,
                             const double* L2, double invR, double* dOUT, int64_t ldo, cudaStream_t stream);

// DN = A_lin + derivative: dDN must hold m*m doubles; it is overwritten.
void launch_jacobian(const JacobianPack& J, const double* dx, double* dDN, cudaStream_t stream);

}  // namespace ca
