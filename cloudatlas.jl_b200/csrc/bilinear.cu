#include "bilinear.h"

#include <algorithm>
#include <numeric>

#include "bilinear_kernels.cuh"

namespace ca {

void DevicePack::upload(const PackedHost& P) {
  upload_meta(P);
  vals.upload(P.vals.data(), P.vals.size());
  CA_CUDA(cudaStreamSynchronize(nullptr));  // the host vectors may die after we return
}

void DevicePack::upload_meta(const PackedHost& P) {
  m = P.m;
  nin = P.nin;
  ntiles = (int32_t)P.tiles.size();
  nt2_tiles = 0;
  for (const TileDesc& T : P.tiles) nt2_tiles += T.nt == 2;
  symmetric = P.symmetric;
  nterms = P.nterms;
  npairs_distinct = P.npairs_distinct;
  padded_fma = P.padded_fma;
  pair_products = P.pair_products;
  tiles.upload(P.tiles.data(), P.tiles.size());
  window_modes = P.window_modes;
  max_window_modes = P.max_window_modes;
  nwin = (int32_t)P.windows.size();
  nwin_nt3 = nwin_nt2 = 0;
  for (const WindowDesc& W : P.windows) {
    nwin_nt3 += P.tiles[W.tile].nt == 3;
    nwin_nt2 += P.tiles[W.tile].nt == 2;
  }
  windows.upload(P.windows.data(), P.windows.size());
  tile_first_window.clear();
  dmma_before_tile.clear();
  int64_t dm = 0;
  for (size_t g = 0; g < P.windows.size(); ++g) {
    if (g == 0 || P.windows[g - 1].last) {
      tile_first_window.push_back((int32_t)g);
      dmma_before_tile.push_back(dm);
    }
    dm += (int64_t)P.windows[g].nks * P.windows[g].nt;
  }
  tile_first_window.push_back((int32_t)P.windows.size());
  dmma_before_tile.push_back(dm);
  dmma_before_tile_index.assign(P.tiles.size() + 1, 0);
  for (size_t t = 0; t < P.tiles.size(); ++t)
    dmma_before_tile_index[t + 1] = dmma_before_tile_index[t] + (int64_t)P.tiles[t].nks * P.tiles[t].nt;
  modes.upload(P.modes.data(), P.modes.size());
  sched.upload(P.sched.data(), P.sched.size());
  for (int w = 0; w < 9; ++w) sched_off[w] = P.sched_off[w];
  for (int w = 0; w < 8; ++w) {
    sched_nt2[w] = P.sched_nt2[w];
    sched_ks2[w] = P.sched_ks2[w];
    sched_ks1[w] = P.sched_ks1[w];
  }
  stream_layout = P.stream_layout;
  pair_order = P.pair_order;
  sched_imbalance = P.sched_imbalance;
  host_pairs = P.pairs;
  for (auto& b : pairs_mt) b.release();
  pairs_win.release();
  rows.upload(P.rows.data(), P.rows.size());
  CA_CUDA(cudaStreamSynchronize(nullptr));  // the host vectors may die after we return
}

const uint2* DevicePack::pairs_for(int mt) const {
  const int slot = mt == 4 ? 2 : (mt == 2 ? 1 : 0);
  DeviceBuffer<uint2>& buf = pairs_mt[slot];
  if (!buf.ptr && !host_pairs.empty()) {
    std::vector<uint2> w(host_pairs.size() + 64, make_uint2(0, 0));  // + one cp.async batch of padding (PairRing)
    for (size_t e = 0; e < host_pairs.size(); ++e) {
      const int a = host_pairs[e] & 0x7fff, b = host_pairs[e] >> 16;
      const uint32_t flag = (host_pairs[e] & 0x8000u) ? kSameA : 0u;
      if (mt == 4) w[e] = make_uint2(factor_word<4>(a) | flag, factor_word<4>(b));
      else if (mt == 2) w[e] = make_uint2(factor_word<2>(a) | flag, factor_word<2>(b));
      else w[e] = make_uint2(factor_word<1>(a) | flag, factor_word<1>(b));
    }
    buf.upload(w.data(), w.size());
    CA_CUDA(cudaStreamSynchronize(nullptr));
  }
  return buf.ptr;
}

const uint32_t* DevicePack::pairs_windowed() const {
  if (!pairs_win.ptr && !host_pairs.empty()) {
    std::vector<uint32_t> w(host_pairs.size() + 64, 0u);
    for (size_t e = 0; e < host_pairs.size(); ++e) {
      const int a = host_pairs[e] & 0x7fff, b = host_pairs[e] >> 16;
      if (a >= 256 || b >= 256) fail(CA_ERR_UNSUPPORTED, "windowed pack with more than 256 modes per window");
      w[e] = pack_window_pair(a, b, (host_pairs[e] & 0x8000u) != 0);
    }
    pairs_win.upload(w.data(), w.size());
    CA_CUDA(cudaStreamSynchronize(nullptr));
  }
  return pairs_win.ptr;
}

// How many CTAs should share a block of states (each taking a range of tiles)?  With p parts, floor(slots / p) blocks
// are in flight at a time and every pass takes 1 / p of a whole block's time: minimise ceil(nblocks / floor(slots / p)) / p.
// p > 1 only if it saves at least 10 % (one CTA per block keeps all CTAs on the same part of the operator in L2).
int choose_parts(int64_t nblocks, int64_t slots, int max_parts) {
  if (nblocks <= 0 || slots <= 0 || nblocks >= 2 * slots || getenv("CA_WIN_NO_SPLIT")) return 1;
  auto cost = [&](int p) {
    const int64_t in_flight = slots / p;
    return in_flight < 1 ? 1e30 : (double)((nblocks + in_flight - 1) / in_flight) / p;
  };
  double lowest = cost(1);
  for (int p = 2; p <= max_parts; ++p) lowest = std::min(lowest, cost(p));
  if (lowest > 0.9 * cost(1)) return 1;
  for (int p = 2; p <= max_parts; ++p)  // the fewest parts within 6 % of the best: every part has its fixed costs
    if (cost(p) <= 1.06 * lowest) return p;
  return 1;
}

namespace {

struct LaunchCfg {
  int sms = 0;
  size_t smem_optin = 0;
};

const LaunchCfg& device_cfg() {
  static thread_local int cached_dev = -1;
  static thread_local LaunchCfg cfg;
  int dev = 0;
  CA_CUDA(cudaGetDevice(&dev));
  if (dev != cached_dev) {
    int v = 0;
    CA_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    cfg.sms = v;
    CA_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    cfg.smem_optin = (size_t)v;
    cached_dev = dev;
  }
  return cfg;
}

// whole tiles per warp when the packer's LPT schedule is balanced (no reduction, no per-tile barriers)
constexpr double kMaxSchedImbalance = 1.08;

template <int MT, int MODE, bool FACT = false>
void launch_warptile(const DevicePack& op, const double* dX, const double* dY, int64_t nstates, int64_t ldx,
                     double* dOUT, int64_t ldo, cudaStream_t stream) {
  auto kern = bilinear_warptile_kernel<MT, MODE, FACT>;
  const size_t smem = warptile_smem_bytes<MT, MODE>(op.nin);
  CA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  CA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
  if (per_sm < 1) fail(CA_ERR_UNSUPPORTED, "warp-tile kernel does not fit (smem %zu B)", smem);
  const int64_t nblocks = (nstates + 8 * MT - 1) / (8 * MT);
  const int grid = (int)std::min<int64_t>(nblocks, (int64_t)device_cfg().sms * per_sm);
  WarpSchedule ws;
  for (int w = 0; w < 9; ++w) ws.off[w] = op.sched_off[w];
  for (int w = 0; w < 8; ++w) {
    ws.nt2[w] = op.sched_nt2[w];
    ws.ks2[w] = op.sched_ks2[w];
    ws.ks1[w] = op.sched_ks1[w];
  }
  kern<<<grid, kThreads, smem, stream>>>(op.tiles.ptr, op.sched.ptr, ws, op.pairs_for(MT), op.vals.ptr, op.rows.ptr,
                                         op.nin, dX, dY, nstates, ldx, dOUT, ldo);
  CA_CUDA(cudaGetLastError());
  count_launch();
}

template <int MT, int MODE>
void launch_one(const DevicePack& op, const double* dX, const double* dY, int64_t nstates, int64_t ldx,
                double* dOUT, int64_t ldo, cudaStream_t stream) {
  auto kern = bilinear_batched_kernel<MT, MODE>;
  const size_t smem = batched_smem_bytes<MT, MODE>(op.nin);
  CA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  CA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
  const int64_t nblocks = (nstates + 8 * MT - 1) / (8 * MT);
  const int64_t slots = (int64_t)device_cfg().sms * std::max(per_sm, 1);
  // Ensembles with fewer blocks of states than CTA slots: the k-split kernel with the TILES (independent output rows)
  // divided among `parts` CTAs per block of states, instead of one CTA (whole tiles per warp) per block and an idle
  // machine -- 1,000 states of the 307-mode model otherwise take as long as 9,472
  WindowParts tp;
  tp.n = 1;
  if (per_sm >= 1 && nblocks * 2 <= slots && op.ntiles >= 2)
    tp.n = choose_parts(nblocks, slots, std::min<int>(WindowParts::kMax, op.ntiles));
  tp.first[0] = 0;
  tp.first[tp.n] = op.ntiles;
  if (tp.n > 1) {
    const int64_t total = op.dmma_before_tile_index[(size_t)op.ntiles];
    int t = 0;
    for (int p = 1; p < tp.n; ++p) {
      const int64_t want = total * p / tp.n;
      while (t < op.ntiles && op.dmma_before_tile_index[(size_t)t] < want) ++t;
      tp.first[p] = t;
    }
  }
  if (tp.n == 1 && op.stream_layout && op.sched_imbalance <= kMaxSchedImbalance && !getenv("CA_FORCE_KSPLIT")) {
    // packs in factored pair order: the quadratic evaluation needs no pair products (bilinear_kernels.cuh)
    if (MODE == MODE_QUAD && op.symmetric && op.pair_order == kOrderFactored && !getenv("CA_NO_FACTORED"))
      launch_warptile<MT, MODE_QUAD, true>(op, dX, dY, nstates, ldx, dOUT, ldo, stream);
    else
      launch_warptile<MT, MODE>(op, dX, dY, nstates, ldx, dOUT, ldo, stream);
    return;
  }
  if (per_sm < 1) fail(CA_ERR_UNSUPPORTED, "batched kernel does not fit (smem %zu B)", smem);
  const int grid = (int)std::min<int64_t>(nblocks, slots / tp.n) * tp.n;
  kern<<<grid, kThreads, smem, stream>>>(op.tiles.ptr, tp, op.nt2_tiles, op.pairs_for(MT), op.vals.ptr,
                                         op.rows.ptr, op.nin, dX, dY, nstates, ldx, dOUT, ldo);
  CA_CUDA(cudaGetLastError());
  count_launch();
}

template <int MODE>
void launch_mode(const DevicePack& op, const double* dX, const double* dY, int64_t nstates,
                 int64_t ldx, double* dOUT, int64_t ldo, cudaStream_t stream) {
  const size_t cap = device_cfg().smem_optin;
  // widest state tile that fits; narrow tiles for tiny ensembles
  int mt = 4;
  if (nstates <= 8) mt = 1;
  else if (nstates <= 16) mt = 2;
  while (mt > 1 && (mt == 4 ? batched_smem_bytes<4, MODE>(op.nin) : batched_smem_bytes<2, MODE>(op.nin)) > cap)
    mt /= 2;
  if (mt == 1 && batched_smem_bytes<1, MODE>(op.nin) > cap)
    fail(CA_ERR_UNSUPPORTED, "state tile for m = %d does not fit in shared memory", op.m);
  if (mt == 4) launch_one<4, MODE>(op, dX, dY, nstates, ldx, dOUT, ldo, stream);
  else if (mt == 2) launch_one<2, MODE>(op, dX, dY, nstates, ldx, dOUT, ldo, stream);
  else launch_one<1, MODE>(op, dX, dY, nstates, ldx, dOUT, ldo, stream);
}

}  // namespace

int window_modes_for(int32_t m) {
  if (const char* f = getenv("CA_FORCE_WINDOW_MODES")) {  // test hook: exercise the windowed kernel at small m
    const int v = atoi(f);
    if (v >= 4) return v;
  }
  // whole-tile packing while x and y/v tiles of 32 states fit next to the reduction buffer (m <= ~380)
  return batched_smem_bytes<4, MODE_JVP>(m + 1) <= (size_t)227 * 1024 ? 0 : kWindowModes;
}

int batched_variant(const DevicePack& op, int64_t nstates) {
  if (op.window_modes > 0) return 4;
  // small ensembles (fewer blocks of 32 states than half the CTA slots): k-split kernel with the tiles divided among CTAs
  const int64_t nblocks = (nstates + 31) / 32, slots = 2 * (int64_t)device_cfg().sms;
  if (nstates > 0 && nblocks * 2 <= slots && op.ntiles >= 2 && !getenv("CA_WIN_NO_SPLIT")) return 3;
  return op.stream_layout && op.sched_imbalance <= kMaxSchedImbalance && !getenv("CA_FORCE_KSPLIT") ? 0 : 3;
}

int nt_max_for(int window_modes) {
  if (window_modes <= 0) return kNtMax;
  const char* e = getenv("CA_NT3");  // "0": never, anything else: always, unset: the packer decides (pack.cpp)
  return e && e[0] == '0' ? kNtMax : 3;
}

bool batched_fits(const DevicePack& op, int mode) {
  if (op.window_modes > 0) return true;  // windowed packs have an m-independent footprint
  const size_t cap = device_cfg().smem_optin;
  return (mode == MODE_QUAD ? batched_smem_bytes<1, MODE_QUAD>(op.nin) : batched_smem_bytes<1, MODE_JVP>(op.nin)) <= cap;
}

namespace {
template <int MODE, bool NT3>
void launch_windowed_nt(const DevicePack& op, const double* dX, const double* dY, int64_t nstates, int64_t ldx,
                        double* dOUT, int64_t ldo, cudaStream_t stream) {
  auto kern = bilinear_windowed_kernel<MODE, NT3, false>;  // (same resources as the SPLIT instantiation)
  const int wmax = std::max(op.max_window_modes, 1);
  const size_t smem = windowed_smem_bytes<MODE, NT3>(wmax);
  CA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  constexpr int threads = WinCfg<MODE, NT3>::THREADS;
  CA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
  if (per_sm < 1) fail(CA_ERR_UNSUPPORTED, "windowed kernel does not fit (smem %zu B)", smem);
  constexpr int teams = WinCfg<MODE, NT3>::TEAMS;  // blocks of 32 states per CTA and pass
  const int64_t nblocks = ((nstates + 8 * kWinMT - 1) / (8 * kWinMT) + teams - 1) / teams;
  const int64_t slots = (int64_t)device_cfg().sms * per_sm;
  // ensembles too small to give every SM a block of states: the tiles (independent output rows) are split into `parts`
  // ranges of about equal DMMA count and every block of states is walked by `parts` CTAs, one range each
  // ensembles that do not fill whole passes of the machine: see choose_parts
  WindowParts wp;
  wp.n = choose_parts(nblocks, slots, std::min<int>(WindowParts::kMax, (int)op.tile_first_window.size() - 1));
  wp.first[0] = 0;
  wp.first[wp.n] = op.nwin;
  if (wp.n > 1) {
    const int ntl = (int)op.tile_first_window.size() - 1;
    const int64_t total = op.dmma_before_tile[(size_t)ntl];
    int t = 0;
    for (int p = 1; p < wp.n; ++p) {  // first tile boundary at or after p / n of the work
      const int64_t want = total * p / wp.n;
      while (t < ntl && op.dmma_before_tile[(size_t)t] < want) ++t;
      wp.first[p] = op.tile_first_window[(size_t)t];
    }
  }
  const int grid = (int)std::min<int64_t>(nblocks, std::max<int64_t>(slots / wp.n, 1)) * wp.n;
  if (wp.n > 1) {
    auto kern_split = bilinear_windowed_kernel<MODE, NT3, true>;
    CA_CUDA(cudaFuncSetAttribute(kern_split, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern_split<<<grid, threads, smem, stream>>>(op.windows.ptr, wp, op.nwin_nt3, op.nwin_nt2, op.modes.ptr, op.pairs_windowed(),
                                                op.vals.ptr, op.rows.ptr, wmax, op.nin - 1, dX, dY, nstates, ldx, dOUT, ldo);
  } else {
    kern<<<grid, threads, smem, stream>>>(op.windows.ptr, wp, op.nwin_nt3, op.nwin_nt2, op.modes.ptr, op.pairs_windowed(),
                                          op.vals.ptr, op.rows.ptr, wmax, op.nin - 1, dX, dY, nstates, ldx, dOUT, ldo);
  }
  CA_CUDA(cudaGetLastError());
  count_launch();
}

template <int MODE>
void launch_windowed(const DevicePack& op, const double* dX, const double* dY, int64_t nstates, int64_t ldx,
                     double* dOUT, int64_t ldo, cudaStream_t stream) {
  if (op.nwin_nt3 > 0) launch_windowed_nt<MODE, true>(op, dX, dY, nstates, ldx, dOUT, ldo, stream);
  else launch_windowed_nt<MODE, false>(op, dX, dY, nstates, ldx, dOUT, ldo, stream);
}
}  // namespace

void launch_batched(const DevicePack& op, int mode, const double* dX, const double* dY,
                    int64_t nstates, int64_t ldx, double* dOUT, int64_t ldo, cudaStream_t stream) {
  if (nstates <= 0) return;
  if (ldx < nstates || ldo < nstates) fail(CA_ERR_INVALID, "leading dimension smaller than nstates");
  if (mode != MODE_QUAD && !dY) fail(CA_ERR_INVALID, "second ensemble missing");
  if (mode != MODE_ORDERED && !op.symmetric) fail(CA_ERR_INVALID, "quadratic/JVP evaluation needs the symmetric pack");
  if (op.window_modes > 0) {
    switch (mode) {
      case MODE_QUAD: launch_windowed<MODE_QUAD>(op, dX, dY, nstates, ldx, dOUT, ldo, stream); break;
      case MODE_ORDERED: launch_windowed<MODE_ORDERED>(op, dX, dY, nstates, ldx, dOUT, ldo, stream); break;
      case MODE_JVP: launch_windowed<MODE_JVP>(op, dX, dY, nstates, ldx, dOUT, ldo, stream); break;
      default: fail(CA_ERR_INVALID, "unknown evaluation mode %d", mode);
    }
    return;
  }
  switch (mode) {
    case MODE_QUAD: launch_mode<MODE_QUAD>(op, dX, dY, nstates, ldx, dOUT, ldo, stream); break;
    case MODE_ORDERED: launch_mode<MODE_ORDERED>(op, dX, dY, nstates, ldx, dOUT, ldo, stream); break;
    case MODE_JVP: launch_mode<MODE_JVP>(op, dX, dY, nstates, ldx, dOUT, ldo, stream); break;
    default: fail(CA_ERR_INVALID, "unknown evaluation mode %d", mode);
  }
}

// ----------------------------------------------------------------------------------- Jacobian
// One thread per structurally non-zero Jacobian entry; entries of a 32-wide slice are stored
// column-major so a warp reads 32 consecutive (src,val) per step.  Slices are formed after sorting
// outputs by contribution count inside windows of 4096 outputs, which keeps padding small while
// neighbouring threads still write neighbouring addresses.
__global__ void __launch_bounds__(256)
jacobian_kernel(const int64_t* __restrict__ slice_off, const int32_t* __restrict__ out_pos,
                const uint16_t* __restrict__ src, const double* __restrict__ val, int64_t nslices,
                int m, const double* __restrict__ x, double* __restrict__ DN) {
  extern __shared__ double xs[];
  for (int j = threadIdx.x; j < m; j += blockDim.x) xs[j] = x[j];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t sl = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); sl < nslices;
       sl += warps_total) {
    const int64_t b = slice_off[sl], e = slice_off[sl + 1];
    double acc0 = 0.0, acc1 = 0.0;
    int64_t p = b + lane;
    for (; p + 32 < e; p += 64) {
      acc0 = fma(val[p], xs[src[p]], acc0);
      acc1 = fma(val[p + 32], xs[src[p + 32]], acc1);
    }
    if (p < e) acc0 = fma(val[p], xs[src[p]], acc0);
    const int32_t o = out_pos[sl * 32 + lane];
    if (o >= 0) DN[o] = acc0 + acc1;
  }
}

void JacobianPack::build(const std::vector<Term>& coo, int32_t m_) {
  m = m_;
  // contributions keyed by output entry (column j, row i) so that the threaded row-bucketed sort applies and
  // consecutive outputs are consecutive rows of one column (coalesced stores): Term{i = column, a = row, b = source}
  struct E { int64_t out; int32_t src; double v; };
  std::vector<Term> es;
  es.reserve(coo.size() * 2);
  for (const Term& t : coo) {
    es.push_back(Term{t.a, t.i, t.b, t.v});  // DN[i,j] += v * x[k]
    es.push_back(Term{t.b, t.i, t.a, t.v});  // DN[i,k] += v * x[j]
  }
  const std::vector<Term> merged = merge_terms(std::move(es), /*symmetric=*/false, m);
  std::vector<E> mg(merged.size());
  for (size_t r = 0; r < merged.size(); ++r)
    mg[r] = E{(int64_t)merged[r].a + (int64_t)m * merged[r].i, merged[r].b, merged[r].v};
  nentries = (int64_t)mg.size();
  // outputs and their ranges
  std::vector<int64_t> ostart;
  std::vector<int64_t> opos;
  for (size_t r = 0; r < mg.size();) {
    size_t e = r;
    while (e < mg.size() && mg[e].out == mg[r].out) ++e;
    ostart.push_back((int64_t)r);
    opos.push_back(mg[r].out);
    r = e;
  }
  ostart.push_back((int64_t)mg.size());
  nout = (int64_t)opos.size();
  std::vector<int64_t> ord(nout);
  std::iota(ord.begin(), ord.end(), 0);
  constexpr int64_t kWindow = 4096;
  for (int64_t w = 0; w < nout; w += kWindow) {
    auto b = ord.begin() + w, e = ord.begin() + std::min(nout, w + kWindow);
    std::stable_sort(b, e, [&](int64_t x, int64_t y) {
      return ostart[x + 1] - ostart[x] > ostart[y + 1] - ostart[y];
    });
  }
  nslices = (nout + 31) / 32;
  std::vector<int64_t> soff(nslices + 1, 0);
  std::vector<int32_t> pos((size_t)nslices * 32, -1);
  for (int64_t s = 0; s < nslices; ++s) {
    int64_t len = 0;
    for (int l = 0; l < 32 && s * 32 + l < nout; ++l) {
      const int64_t o = ord[s * 32 + l];
      len = std::max(len, ostart[o + 1] - ostart[o]);
      pos[s * 32 + l] = (int32_t)opos[o];
    }
    soff[s + 1] = soff[s] + len * 32;
  }
  std::vector<uint16_t> hsrc((size_t)soff[nslices], 0);
  std::vector<double> hval((size_t)soff[nslices], 0.0);
  for (int64_t s = 0; s < nslices; ++s)
    for (int l = 0; l < 32 && s * 32 + l < nout; ++l) {
      const int64_t o = ord[s * 32 + l];
      for (int64_t r = ostart[o]; r < ostart[o + 1]; ++r) {
        const size_t at = (size_t)soff[s] + (size_t)(r - ostart[o]) * 32 + l;
        hsrc[at] = (uint16_t)mg[r].src;
        hval[at] = mg[r].v;
      }
    }
  // host mirror for the ensemble kernel's output-major copy -- only where that kernel can run at all (its tile of
  // 32 states has to fit in shared memory)
  const bool mirror = (size_t)m * 32 * sizeof(double) <= (size_t)227 * 1024;
  h_out.resize(mirror ? mg.size() : 0);
  h_src.resize(h_out.size());
  h_val.resize(h_out.size());
  for (size_t r = 0; r < h_out.size(); ++r) {
    h_out[r] = mg[r].out;
    h_src[r] = (uint16_t)mg[r].src;
    h_val[r] = mg[r].v;
  }
  optr.release();
  slice_off.upload(soff.data(), soff.size());
  out_pos.upload(pos.data(), pos.size());
  src.upload(hsrc.data(), hsrc.size());
  val.upload(hval.data(), hval.size());
  CA_CUDA(cudaStreamSynchronize(nullptr));
}

// Ensemble version: a CTA holds the coefficients of 32*SPL states in shared memory ([mode][state]); a warp takes one
// output entry at a time, each lane carries SPL states (lane, lane + 32, ...): the entry's (source, value) list is a
// warp-uniform broadcast read of the same sliced-ELL structure -- amortised over SPL states, four contributions in
// flight -- and the state values and the output rows are 256-byte coalesced segments.
// Bound: HBM write of 8 m^2 bytes per state (small m) / issue rate of the contribution loop (m >~ 100).
template <int SPL, bool LINEAR>
__global__ void __launch_bounds__(256)
jacobian_batched_kernel(const int64_t* __restrict__ optr, const uint16_t* __restrict__ src,
                        const double* __restrict__ val, int m,
                        const double* __restrict__ X, int64_t nstates, int64_t ldx, const double* __restrict__ L1,
                        const double* __restrict__ L2, double invR, double* __restrict__ OUT, int64_t ldo,
                        int64_t ochunk) {
  constexpr int SB = 32 * SPL;
  extern __shared__ double xs[];  // [m][SB]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t s0 = (int64_t)blockIdx.x * SB;
  for (int idx = threadIdx.x; idx < m * SB; idx += 256) {
    const int j = idx / SB, s = idx % SB;
    xs[idx] = s0 + s < nstates ? X[(s0 + s) + ldx * (int64_t)j] : 0.0;
  }
  __syncthreads();
  const int64_t mm = (int64_t)m * m;
  const int64_t ob = (int64_t)blockIdx.y * ochunk, oe = min(mm, ob + ochunk);
  const double* xl = xs + lane;
  for (int64_t o = ob + warp; o < oe; o += 8) {
    double acc[SPL];
#pragma unroll
    for (int q = 0; q < SPL; ++q) acc[q] = 0.0;
    int64_t p = __ldg(optr + o);
    const int64_t end = __ldg(optr + o + 1);  // structurally zero outputs have empty lists and are written as zeros
    for (; p + 4 <= end; p += 4) {
      double v[4];
      int sc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        v[u] = __ldg(val + p + u);
        sc[u] = (int)__ldg(src + p + u) * SB;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int q = 0; q < SPL; ++q) acc[q] = fma(v[u], xl[sc[u] + 32 * q], acc[q]);
    }
    for (; p < end; ++p) {
      const double v = __ldg(val + p);
      const int sc = (int)__ldg(src + p) * SB;
#pragma unroll
      for (int q = 0; q < SPL; ++q) acc[q] = fma(v, xl[sc + 32 * q], acc[q]);
    }
    const double lin = LINEAR ? fma(__ldg(L2 + o), invR, __ldg(L1 + o)) : 0.0;
    double* out = OUT + s0 + lane + ldo * o;
#pragma unroll
    for (int q = 0; q < SPL; ++q)
      if (s0 + lane + 32 * q < nstates) out[32 * q] = acc[q] + lin;
  }
}

template <int SPL>
static void launch_jacobian_batched_spl(const JacobianPack& J, const double* dX, int64_t nstates, int64_t ldx,
                                        const double* L1, const double* L2, double invR, double* dOUT, int64_t ldo,
                                        cudaStream_t stream) {
  constexpr int SB = 32 * SPL;
  const size_t smem = (size_t)J.m * SB * sizeof(double);
  const int64_t nsb = (nstates + SB - 1) / SB, mm = (int64_t)J.m * J.m;
  if (nsb > 2147483647LL) fail(CA_ERR_UNSUPPORTED, "too many states for one launch");
  // enough CTAs to fill the machine a few times over, at least 64 outputs (8 per warp) each
  int64_t ychunks = std::max<int64_t>(1, std::min<int64_t>((4 * (int64_t)device_cfg().sms + nsb - 1) / nsb, (mm + 63) / 64));
  ychunks = std::min<int64_t>(ychunks, 65535);
  const int64_t ochunk = (mm + ychunks - 1) / ychunks;
  const dim3 grid((unsigned)nsb, (unsigned)((mm + ochunk - 1) / ochunk));
  if (L1) {
    auto kern = jacobian_batched_kernel<SPL, true>;
    CA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 256, smem, stream>>>(J.optr.ptr, J.osrc.ptr, J.oval.ptr, J.m, dX, nstates, ldx, L1, L2, invR, dOUT, ldo,
                                      ochunk);
  } else {
    auto kern = jacobian_batched_kernel<SPL, false>;
    CA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 256, smem, stream>>>(J.optr.ptr, J.osrc.ptr, J.oval.ptr, J.m, dX, nstates, ldx, nullptr, nullptr, 0.0,
                                      dOUT, ldo, ochunk);
  }
  CA_CUDA(cudaGetLastError());
  count_launch();
}

void JacobianPack::ensure_output_major() const {
  if (optr.ptr) return;
  if (h_out.size() != (size_t)nentries)
    fail(CA_ERR_UNSUPPORTED, "batched dense Jacobian: m = %d does not fit in shared memory", m);
  const int64_t mm = (int64_t)m * m;
  std::vector<int64_t> ptr((size_t)mm + 1, 0);
  for (int64_t o : h_out) ptr[(size_t)o + 1]++;  // h_out is sorted: the lists are already contiguous
  for (int64_t o = 0; o < mm; ++o) ptr[(size_t)o + 1] += ptr[(size_t)o];
  optr.upload(ptr.data(), ptr.size());
  osrc.upload(h_src.data(), h_src.size());
  oval.upload(h_val.data(), h_val.size());
  CA_CUDA(cudaStreamSynchronize(nullptr));
  std::vector<int64_t>().swap(h_out);
  std::vector<uint16_t>().swap(h_src);
  std::vector<double>().swap(h_val);
}

void launch_jacobian_batched(const JacobianPack& J, const double* dX, int64_t nstates, int64_t ldx, const double* L1,
                             const double* L2, double invR, double* dOUT, int64_t ldo, cudaStream_t stream) {
  if (nstates <= 0 || J.m <= 0) return;
  J.ensure_output_major();
  if (ldx < nstates || ldo < nstates) fail(CA_ERR_INVALID, "leading dimension smaller than nstates");
  const size_t per32 = (size_t)J.m * 32 * sizeof(double);
  if (per32 > device_cfg().smem_optin) fail(CA_ERR_UNSUPPORTED, "batched dense Jacobian: m = %d does not fit in shared memory", J.m);
  int spl = 8;  // states per lane: as many as keep >= 2 CTAs per SM and have states to work on
  if (const char* e = getenv("CA_JAC_SPL")) spl = atoi(e);  // tuning hook
  while (spl > 1 && (per32 * spl > (size_t)100 * 1024 || nstates <= 16 * spl)) spl /= 2;
  if (spl >= 8) launch_jacobian_batched_spl<8>(J, dX, nstates, ldx, L1, L2, invR, dOUT, ldo, stream);
  else if (spl >= 4) launch_jacobian_batched_spl<4>(J, dX, nstates, ldx, L1, L2, invR, dOUT, ldo, stream);
  else if (spl == 2) launch_jacobian_batched_spl<2>(J, dX, nstates, ldx, L1, L2, invR, dOUT, ldo, stream);
  else launch_jacobian_batched_spl<1>(J, dX, nstates, ldx, L1, L2, invR, dOUT, ldo, stream);
}

void launch_jacobian(const JacobianPack& J, const double* dx, double* dDN, cudaStream_t stream) {
  CA_CUDA(cudaMemsetAsync(dDN, 0, (size_t)J.m * J.m * sizeof(double), stream));
  if (J.nslices == 0) return;
  const int warps_per_block = 8;
  const int64_t want = (J.nslices + warps_per_block - 1) / warps_per_block;
  const int grid = (int)std::min<int64_t>(want, (int64_t)device_cfg().sms * 8);
  jacobian_kernel<<<grid, 256, (size_t)J.m * sizeof(double), stream>>>(
      J.slice_off.ptr, J.out_pos.ptr, J.src.ptr, J.val.ptr, J.nslices, J.m, dx, dDN);
  CA_CUDA(cudaGetLastError());
  count_launch();
}

}  // namespace ca
