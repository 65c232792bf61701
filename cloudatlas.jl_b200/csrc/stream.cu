// Few-state sparse-bilinear evaluation: coalesced CSR stream with warp-level reductions (HBM-bound).
//
// Replaces the reference's COO loop (src/SparseBilinear.jl:56-65) when there are too few states to
// amortise the operator over a DMMA tile: single-state N(x), N(x,y), f(x,R) and the matrix-free Jacobian
// action for Newton-Krylov.  Each warp owns one chunk of one output row; lanes read consecutive
// (a|b, value) entries (12 B, coalesced), gather the factors of up to NB states from shared memory,
// and the 32 partial sums are combined by shuffles; a second tiny kernel adds the chunk partials of every
// row in chunk order, so results do not depend on scheduling.
#include <algorithm>

#include "bilinear.h"
#include "bilinear_kernels.cuh"

namespace ca {

namespace {

constexpr int kStreamThreads = 512;
constexpr int kStreamUnroll = 4;  // independent entry loads in flight per lane

template <int NB, int MODE>
__global__ void __launch_bounds__(kStreamThreads)
stream_kernel(const uint32_t* __restrict__ ab, const double* __restrict__ val,
              const StreamChunk* __restrict__ chunks, int64_t chunk0, int64_t nchunks, int nin,
              const double* __restrict__ X, const double* __restrict__ Y, int64_t ldx, int nb,
              double* __restrict__ partial) {
  extern __shared__ double sm[];
  double* xs = sm;                                    // [NB][nin]: one contiguous vector per state, so that the
  double* ys = MODE == MODE_QUAD ? sm : sm + (size_t)nin * NB;  // consecutive second factors of a row hit consecutive banks
  for (int o = threadIdx.x; o < nin * NB; o += kStreamThreads) {
    const int j = o / NB, s = o % NB;
    double vx = 0.0, vy = 0.0;
    if (j == nin - 1) {
      vx = 1.0;
      vy = MODE == MODE_JVP ? 0.0 : 1.0;
    } else if (s < nb) {
      vx = X[s + ldx * (int64_t)j];
      if (MODE != MODE_QUAD) vy = Y[s + ldx * (int64_t)j];
    }
    xs[s * nin + j] = vx;
    if (MODE != MODE_QUAD) ys[s * nin + j] = vy;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * kStreamThreads) >> 5;
  for (int64_t c = chunk0 + (((int64_t)blockIdx.x * kStreamThreads + threadIdx.x) >> 5); c < chunk0 + nchunks; c += warps) {
    const StreamChunk ch = chunks[c];
    double acc[NB];
#pragma unroll
    for (int s = 0; s < NB; ++s) acc[s] = 0.0;
    const uint32_t* pab = ab + ch.begin;
    const double* pval = val + ch.begin;
    auto contribute = [&](uint32_t w, double v) {
      const int a = w & 0xffff, b = w >> 16;
#pragma unroll
      for (int s = 0; s < NB; ++s) {
        double p;
        if (MODE == MODE_QUAD) p = xs[s * nin + a] * xs[s * nin + b];
        else if (MODE == MODE_ORDERED) p = xs[s * nin + a] * ys[s * nin + b];
        else p = xs[s * nin + a] * ys[s * nin + b] + ys[s * nin + a] * xs[s * nin + b];
        acc[s] = fma(v, p, acc[s]);
      }
    };
    int e = lane;
    for (; e + 32 * (kStreamUnroll - 1) < ch.count; e += 32 * kStreamUnroll) {
      uint32_t w[kStreamUnroll];
      double v[kStreamUnroll];
#pragma unroll
      for (int u = 0; u < kStreamUnroll; ++u) {
        w[u] = __ldg(pab + e + 32 * u);
        v[u] = __ldg(pval + e + 32 * u);
      }
#pragma unroll
      for (int u = 0; u < kStreamUnroll; ++u) contribute(w[u], v[u]);
    }
    for (; e < ch.count; e += 32) contribute(__ldg(pab + e), __ldg(pval + e));
#pragma unroll
    for (int s = 0; s < NB; ++s) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[s] += __shfl_xor_sync(0xffffffffu, acc[s], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < NB; ++s) partial[c * 8 + s] = acc[s];
    }
  }
}

// Dense linear part (L1 + L2/R) v, one warp per (row, chunk of kLinCols columns): every row of L1r / L2r is read once,
// coalesced, and applied to all nb states; partials go next to the quadratic chunk partials.
constexpr int kLinCols = 512;

template <int MODE>
__global__ void __launch_bounds__(kStreamThreads)
stream_linear_kernel(const double* __restrict__ L1r, const double* __restrict__ L2r, double invR, int m, int nb,
                     int nlc, const double* __restrict__ X, const double* __restrict__ Y, int64_t ldx,
                     double* __restrict__ lin_partial, int row0, int nrows) {
  const int lane = threadIdx.x & 31;
  const int64_t wl = ((int64_t)blockIdx.x * kStreamThreads + threadIdx.x) >> 5;
  if (wl >= (int64_t)nrows * nlc) return;
  const int64_t w = wl + (int64_t)row0 * nlc;  // rows [row0, row0 + nrows) of the operator
  const int row = (int)(w / nlc), j0 = (int)(w % nlc) * kLinCols, j1 = min(m, j0 + kLinCols);
  const double* V = MODE == MODE_JVP ? Y : X;  // for the Jacobian action the linear part acts on the direction
  const double* l1 = L1r + (size_t)row * m;
  const double* l2 = L2r + (size_t)row * m;
  double acc[8];
#pragma unroll
  for (int s = 0; s < 8; ++s) acc[s] = 0.0;
#pragma unroll 4
  for (int j = j0 + lane; j < j1; j += 32) {
    const double l = fma(__ldg(l2 + j), invR, __ldg(l1 + j));
    const double* v = V + ldx * (int64_t)j;
#pragma unroll
    for (int s = 0; s < 8; ++s)
      if (s < nb) acc[s] = fma(l, v[s], acc[s]);
  }
#pragma unroll
  for (int s = 0; s < 8; ++s) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[s] += __shfl_xor_sync(0xffffffffu, acc[s], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < 8; ++s) lin_partial[w * 8 + s] = acc[s];
  }
}

// out[s, row] = sum of the row's chunk partials (+ the partials of the dense linear part), in fixed order
__global__ void __launch_bounds__(kStreamThreads)
stream_combine_kernel(const double* __restrict__ partial, const int32_t* __restrict__ row_chunk_ptr, int m,
                      int nb, const double* __restrict__ lin_partial, int nlc, double* __restrict__ OUT,
                      int64_t ldo, int row0, int nrows) {
  const int64_t t = (int64_t)blockIdx.x * kStreamThreads + threadIdx.x;
  const int row = row0 + (int)(t >> 3), s = (int)(t & 7);
  if (row >= row0 + nrows || row >= m || s >= nb) return;
  double acc = 0.0;
  if (lin_partial)
    for (int c = 0; c < nlc; ++c) acc += lin_partial[((int64_t)row * nlc + c) * 8 + s];
  for (int c = row_chunk_ptr[row]; c < row_chunk_ptr[row + 1]; ++c) acc += partial[(int64_t)c * 8 + s];
  OUT[s + ldo * (int64_t)row] = acc;
}

template <int NB, int MODE>
void launch_nb(const StreamPack& sp, const double* dX, const double* dY, int nb, int64_t ldx, double* dOUT,
               int64_t ldo, const double* L1r, const double* L2r, double invR, cudaStream_t st, int row0, int nrows) {
  // rows [row0, row0 + nrows) of the operator: their chunks are a contiguous range (chunks are listed row by row)
  const int64_t chunk0 = sp.h_row_chunk_ptr[(size_t)row0], nch = sp.h_row_chunk_ptr[(size_t)row0 + nrows] - chunk0;
  const size_t smem = (MODE == MODE_QUAD ? 1 : 2) * (size_t)sp.nin * NB * sizeof(double);
  auto kern = stream_kernel<NB, MODE>;
  CA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int sms = 0, dev = 0;
  CA_CUDA(cudaGetDevice(&dev));
  CA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int per_sm = 0;
  CA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kStreamThreads, smem));
  if (per_sm < 1) fail(CA_ERR_UNSUPPORTED, "streaming kernel does not fit (smem %zu B)", smem);
  double* partial = sp.partial_for(st);
  if (nch > 0) {
    const int64_t want = (nch + kStreamThreads / 32 - 1) / (kStreamThreads / 32);
    if (const char* e = getenv("CA_STREAM_CTAS")) per_sm = std::max(1, std::min(per_sm, atoi(e)));  // tuning hook
    const int grid = (int)std::min<int64_t>(want, (int64_t)sms * per_sm);
    kern<<<grid, kStreamThreads, smem, st>>>(sp.ab.ptr, sp.val.ptr, sp.chunks.ptr, chunk0, nch, sp.nin, dX, dY,
                                             ldx, nb, partial);
    CA_CUDA(cudaGetLastError());
    count_launch();
  }
  const int nlc = (sp.m + kLinCols - 1) / kLinCols;
  double* lin_partial = L1r ? partial + (size_t)std::max<int64_t>(sp.nchunks, 1) * 8 : nullptr;
  if (L1r) {
    const int64_t warps = (int64_t)nrows * nlc;
    const int gridl = (int)((warps * 32 + kStreamThreads - 1) / kStreamThreads);
    stream_linear_kernel<MODE><<<gridl, kStreamThreads, 0, st>>>(L1r, L2r, invR, sp.m, nb, nlc, dX, dY, ldx, lin_partial, row0, nrows);
    CA_CUDA(cudaGetLastError());
    count_launch();
  }
  const int grid2 = (nrows * 8 + kStreamThreads - 1) / kStreamThreads;
  stream_combine_kernel<<<grid2, kStreamThreads, 0, st>>>(partial, sp.row_chunk_ptr.ptr, sp.m, nb, lin_partial, nlc,
                                                          dOUT, ldo, row0, nrows);
  CA_CUDA(cudaGetLastError());
  count_launch();
}

template <int MODE>
void launch_mode(const StreamPack& sp, const double* dX, const double* dY, int64_t nstates, int64_t ldx,
                 double* dOUT, int64_t ldo, const double* L1r, const double* L2r, double invR, cudaStream_t st,
                 int row0, int nrows) {
  // widest batch whose state tile fits in shared memory
  int optin = 0, dev = 0;
  CA_CUDA(cudaGetDevice(&dev));
  CA_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const size_t per_state = (MODE == MODE_QUAD ? 1 : 2) * (size_t)sp.nin * sizeof(double);
  int nbmax = 8;
  while (nbmax > 1 && per_state * nbmax > (size_t)optin / 2) nbmax /= 2;  // keep two CTAs per SM
  if (per_state > (size_t)optin) fail(CA_ERR_UNSUPPORTED, "m = %d is too large for the streaming kernel", sp.m);
  for (int64_t s0 = 0; s0 < nstates;) {
    const int64_t left = nstates - s0;
    int nb = (int)std::min<int64_t>(left, nbmax);
    const double* x = dX + s0;
    const double* y = dY ? dY + s0 : nullptr;
    double* o = dOUT + s0;
    if (nb >= 8) { nb = 8; launch_nb<8, MODE>(sp, x, y, nb, ldx, o, ldo, L1r, L2r, invR, st, row0, nrows); }
    else if (nb >= 4) { nb = 4; launch_nb<4, MODE>(sp, x, y, nb, ldx, o, ldo, L1r, L2r, invR, st, row0, nrows); }
    else if (nb >= 2) { nb = 2; launch_nb<2, MODE>(sp, x, y, nb, ldx, o, ldo, L1r, L2r, invR, st, row0, nrows); }
    else { nb = 1; launch_nb<1, MODE>(sp, x, y, nb, ldx, o, ldo, L1r, L2r, invR, st, row0, nrows); }
    s0 += nb;
  }
}

}  // namespace

void StreamPack::build_rect(std::vector<Term> terms, int32_t m_, int32_t n_inputs) {
  build(std::move(terms), m_, true);
  nin = n_inputs + 1;
  if (nin > 65536) fail(CA_ERR_UNSUPPORTED, "%d inputs exceed the 16-bit mode index", n_inputs);
}

void StreamPack::build(std::vector<Term> terms, int32_t m_, bool symmetric_) {
  m = m_;
  nin = m + 1;
  symmetric = symmetric_;
  if (nin > 65536) fail(CA_ERR_UNSUPPORTED, "m = %d exceeds the 16-bit mode index", m);
  for (auto& t : terms)
    if (symmetric && t.a > t.b) std::swap(t.a, t.b);
  sort_terms(terms, m);
  std::vector<uint32_t> hab;
  std::vector<double> hval;
  std::vector<int64_t> rowptr(m + 1, 0);
  hab.reserve(terms.size());
  hval.reserve(terms.size());
  for (size_t r = 0; r < terms.size();) {
    size_t e = r;
    double s = 0;
    while (e < terms.size() && terms[e].i == terms[r].i && terms[e].a == terms[r].a && terms[e].b == terms[r].b)
      s += terms[e++].v;
    if (s != 0.0) {
      hab.push_back((uint32_t)terms[r].a | ((uint32_t)terms[r].b << 16));
      hval.push_back(s);
      rowptr[terms[r].i + 1]++;
    }
    r = e;
  }
  for (int i = 0; i < m; ++i) rowptr[i + 1] += rowptr[i];
  nentries = (int64_t)hab.size();
  // chunk length: enough chunks to fill the machine, long enough to amortise the reduction
  int64_t clen = nentries / (148 * 64);
  clen = std::max<int64_t>(128, std::min<int64_t>(4096, clen));
  clen = (clen + 31) / 32 * 32;
  std::vector<StreamChunk> hchunks;
  std::vector<int32_t> rcp(m + 1, 0);
  for (int i = 0; i < m; ++i) {
    for (int64_t b = rowptr[i]; b < rowptr[i + 1]; b += clen)
      hchunks.push_back(StreamChunk{i, (int32_t)std::min<int64_t>(clen, rowptr[i + 1] - b), b});
    rcp[i + 1] = (int32_t)hchunks.size();
  }
  nchunks = (int64_t)hchunks.size();
  ab.upload(hab.data(), hab.size());
  val.upload(hval.data(), hval.size());
  chunks.upload(hchunks.data(), hchunks.size());
  row_chunk_ptr.upload(rcp.data(), rcp.size());
  h_row_chunk_ptr = rcp;
  partials.clear();
  CA_CUDA(cudaStreamSynchronize(nullptr));
}

void StreamPack::adopt(DeviceBuffer<uint32_t>&& ab_, DeviceBuffer<double>&& val_, const std::vector<int64_t>& row_begin,
                       const std::vector<int64_t>& row_end, int32_t m_, bool symmetric_) {
  m = m_;
  nin = m + 1;
  symmetric = symmetric_;
  if (nin > 65536) fail(CA_ERR_UNSUPPORTED, "m = %d exceeds the 16-bit mode index", m);
  nentries = 0;
  for (int i = 0; i < m; ++i) nentries += row_end[(size_t)i] - row_begin[(size_t)i];
  int64_t clen = nentries / (148 * 64);  // same chunking rule as build()
  clen = std::max<int64_t>(128, std::min<int64_t>(4096, clen));
  clen = (clen + 31) / 32 * 32;
  std::vector<StreamChunk> hchunks;
  std::vector<int32_t> rcp((size_t)m + 1, 0);
  for (int i = 0; i < m; ++i) {
    for (int64_t b = row_begin[(size_t)i]; b < row_end[(size_t)i]; b += clen)
      hchunks.push_back(StreamChunk{i, (int32_t)std::min<int64_t>(clen, row_end[(size_t)i] - b), b});
    rcp[(size_t)i + 1] = (int32_t)hchunks.size();
  }
  nchunks = (int64_t)hchunks.size();
  std::swap(ab.ptr, ab_.ptr);
  std::swap(ab.count, ab_.count);
  std::swap(val.ptr, val_.ptr);
  std::swap(val.count, val_.count);
  chunks.upload(hchunks.data(), hchunks.size());
  row_chunk_ptr.upload(rcp.data(), rcp.size());
  h_row_chunk_ptr = rcp;
  partials.clear();
  CA_CUDA(cudaStreamSynchronize(nullptr));
}

double* StreamPack::partial_for(cudaStream_t st) const {
  auto& slot = partials[st];
  if (!slot) {
    slot.reset(new DeviceBuffer<double>);
    // chunk partials, then the partials of the dense linear part (stream_linear_kernel)
    slot->alloc(((size_t)std::max<int64_t>(nchunks, 1) + (size_t)m * ((m + 511) / 512)) * 8);
  }
  return slot->ptr;
}

void launch_stream(const StreamPack& sp, int mode, const double* dX, const double* dY, int64_t nstates,
                   int64_t ldx, double* dOUT, int64_t ldo, const double* L1r, const double* L2r, double invR,
                   cudaStream_t stream, int row_begin, int row_end) {
  if (nstates <= 0) return;
  if (row_end < 0) row_end = sp.m;
  if (row_begin < 0 || row_begin > row_end || row_end > sp.m) fail(CA_ERR_INVALID, "row range [%d,%d) outside 0:%d", row_begin, row_end, sp.m);
  if (row_begin == row_end) return;
  const int row0 = row_begin, nrows = row_end - row_begin;
  if (mode != MODE_QUAD && !dY) fail(CA_ERR_INVALID, "second ensemble missing");
  if (mode != MODE_ORDERED && !sp.symmetric) fail(CA_ERR_INVALID, "quadratic/JVP evaluation needs the symmetric stream");
  switch (mode) {
    case MODE_QUAD: launch_mode<MODE_QUAD>(sp, dX, dY, nstates, ldx, dOUT, ldo, L1r, L2r, invR, stream, row0, nrows); break;
    case MODE_ORDERED: launch_mode<MODE_ORDERED>(sp, dX, dY, nstates, ldx, dOUT, ldo, L1r, L2r, invR, stream, row0, nrows); break;
    case MODE_JVP: launch_mode<MODE_JVP>(sp, dX, dY, nstates, ldx, dOUT, ldo, L1r, L2r, invR, stream, row0, nrows); break;
    default: fail(CA_ERR_INVALID, "unknown evaluation mode %d", mode);
  }
}

}  // namespace ca
