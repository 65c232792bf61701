// C ABI: SparseBilinear entry points (reference: src/SparseBilinear.jl).
#include <algorithm>
#include <memory>
#include <vector>

#include "bilinear.h"
#include "bilinear_kernels.cuh"

using namespace ca;

struct ca_bilinear {
  int device = 0;
  int64_t m = 0, nnz = 0;
  std::vector<Term> coo;  // 0-based copy of the caller's COO, reference order
  DevicePack sym;         // (j,k) ~ (k,j): N(x), JVP
  std::unique_ptr<DevicePack> ord;      // ordered pairs: N(x,y), built on first use
  std::unique_ptr<JacobianPack> jac;    // built on first use
  std::unique_ptr<StreamPack> ssym, sord;  // few-state streaming form, built on first use
  // staging for the host-pointer entry points
  cudaStream_t streams[kStage] = {nullptr, nullptr, nullptr};
  DeviceBuffer<double> stage_in[kStage], stage_out[kStage], small_x, small_y, small_out, jac_out;
  ~ca_bilinear() {
    for (auto& s : streams)
      if (s) cudaStreamDestroy(s);
  }
  void bind() const { CA_CUDA(cudaSetDevice(device)); }
  void need_streams() {
    for (auto& s : streams)
      if (!s) CA_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  }
  DevicePack& ordered() {
    if (!ord) {
      ord.reset(new DevicePack);
      ord->upload(pack_terms(coo, (int32_t)m, /*symmetric=*/false, nt_max_for(window_modes_for((int32_t)m)), window_modes_for((int32_t)m)));
    }
    return *ord;
  }
  StreamPack& stream_sym() {
    if (!ssym) {
      ssym.reset(new StreamPack);
      ssym->build(coo, (int32_t)m, true);
    }
    return *ssym;
  }
  StreamPack& stream_ord() {
    if (!sord) {
      sord.reset(new StreamPack);
      sord->build(coo, (int32_t)m, false);
    }
    return *sord;
  }
  // mode-aware dispatch between the DMMA ensemble kernel and the streaming kernel
  void eval_dev(int mode, const double* dX, const double* dY, int64_t nstates, int64_t ldx, double* dOUT,
                int64_t ldo, cudaStream_t st) {
    DevicePack* dp = mode == MODE_ORDERED ? nullptr : &sym;
    const bool small = nstates <= kStreamMaxStates;
    if (!small) {
      if (mode == MODE_ORDERED) dp = &ordered();
      if (batched_fits(*dp, mode)) {
        launch_batched(*dp, mode, dX, dY, nstates, ldx, dOUT, ldo, st);
        return;
      }
    }
    launch_stream(mode == MODE_ORDERED ? stream_ord() : stream_sym(), mode, dX, dY, nstates, ldx, dOUT, ldo,
                  nullptr, nullptr, 0.0, st);
  }
  JacobianPack& jacobian() {
    if (!jac) {
      jac.reset(new JacobianPack);
      jac->build(coo, (int32_t)m);
    }
    return *jac;
  }
};

namespace ca {

// Shared with the model code: evaluates a packed operator for a host ensemble, chunked and
// double-buffered (H2D of chunk c+1 and D2H of chunk c-1 overlap the kernel of chunk c when the
// host buffers are page-locked, see ca_pin).
void eval_batched_host(const LaunchFn& launch, cudaStream_t streams[kStage], DeviceBuffer<double> in[kStage],
                       DeviceBuffer<double> out[kStage], const double* X, int64_t nstates, double* OUT,
                       int64_t m_in, int64_t m_out, int64_t ld_host) {
  if (nstates <= 0) return;
  if (ld_host < nstates) ld_host = nstates;
  const int64_t bytes_per_state = (m_in + m_out) * (int64_t)sizeof(double);
  // ~128 MB of traffic per chunk (short exposed head and tail of the copy/compute pipeline), rounded to whole
  // waves of the persistent ensemble kernel (SMs x 2 CTAs x 32 states) so that no chunk ends in a partial wave
  int sms = 148;
  {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int64_t wave = (int64_t)sms * 2 * 32;
  int64_t chunk = std::max<int64_t>(wave, (128ll << 20) / bytes_per_state / wave * wave);
  chunk = std::min(chunk, (nstates + 31) / 32 * 32);
  for (int b = 0; b < kStage; ++b) {
    if (in[b].count < (size_t)(chunk * m_in)) in[b].alloc((size_t)(chunk * m_in));
    if (out[b].count < (size_t)(chunk * m_out)) out[b].alloc((size_t)(chunk * m_out));
  }
  int c = 0;
  for (int64_t s0 = 0; s0 < nstates; s0 += chunk, ++c) {
    const int b = c % kStage;
    const int64_t n = std::min(chunk, nstates - s0);
    cudaStream_t st = streams[b];
    CA_CUDA(cudaMemcpy2DAsync(in[b].ptr, (size_t)n * sizeof(double), X + s0,
                              (size_t)ld_host * sizeof(double), (size_t)n * sizeof(double),
                              (size_t)m_in, cudaMemcpyHostToDevice, st));
    launch(in[b].ptr, n, out[b].ptr, st);
    CA_CUDA(cudaMemcpy2DAsync(OUT + s0, (size_t)ld_host * sizeof(double), out[b].ptr,
                              (size_t)n * sizeof(double), (size_t)n * sizeof(double), (size_t)m_out,
                              cudaMemcpyDeviceToHost, st));
  }
  for (int b = 0; b < kStage; ++b) CA_CUDA(cudaStreamSynchronize(streams[b]));
}

}  // namespace ca

extern "C" {

int32_t ca_bilinear_create(const int64_t* ijk, const double* val, int64_t nnz, int64_t m,
                           ca_bilinear** out) {
  return guarded([&] {
    if (!out) fail(CA_ERR_INVALID, "ca_bilinear_create: out is null");
    *out = nullptr;
    if (m <= 0) fail(CA_ERR_INVALID, "ca_bilinear_create: m must be positive");
    if (nnz < 0 || (nnz > 0 && (!ijk || !val)))
      fail(CA_ERR_INVALID, "ijk and val must have same number of rows");
    std::unique_ptr<ca_bilinear> h(new ca_bilinear);
    h->m = m;
    h->nnz = nnz;
    h->coo.resize((size_t)nnz);
    for (int64_t r = 0; r < nnz; ++r) {
      const int64_t i = ijk[r], j = ijk[r + nnz], k = ijk[r + 2 * nnz];
      if (i < 1 || i > m || j < 1 || j > m || k < 1 || k > m)
        fail(CA_ERR_INVALID, "ca_bilinear_create: entry %lld has index (%lld,%lld,%lld) outside 1:%lld",
             (long long)r + 1, (long long)i, (long long)j, (long long)k, (long long)m);
      h->coo[(size_t)r] = Term{(int32_t)(i - 1), (int32_t)(j - 1), (int32_t)(k - 1), val[r]};
    }
    require_device();
    CA_CUDA(cudaGetDevice(&h->device));
    h->sym.upload(pack_terms(h->coo, (int32_t)m, /*symmetric=*/true, nt_max_for(window_modes_for((int32_t)m)), window_modes_for((int32_t)m)));
    *out = h.release();
  });
}

int32_t ca_bilinear_destroy(ca_bilinear* h) {
  return guarded([&] {
    if (!h) return;
    cudaSetDevice(h->device);
    delete h;
  });
}

int32_t ca_bilinear_info(const ca_bilinear* h, int64_t* m, int64_t* nnz, int64_t* nnz_sym,
                         int64_t* npairs, int64_t* padded_fma_per_state) {
  return guarded([&] {
    if (!h) fail(CA_ERR_INVALID, "null handle");
    if (m) *m = h->m;
    if (nnz) *nnz = h->nnz;
    if (nnz_sym) *nnz_sym = h->sym.nterms;
    if (npairs) *npairs = h->sym.npairs_distinct;
    if (padded_fma_per_state) *padded_fma_per_state = h->sym.padded_fma;
  });
}

int32_t ca_bilinear_eval(ca_bilinear* h, const double* x, const double* y, double* out) {
  return guarded([&] {
    if (!h || !x || !out) fail(CA_ERR_INVALID, "ca_bilinear_eval: null argument");
    h->bind();
    const size_t m = (size_t)h->m;
    h->small_x.upload(x, m);
    if (y && y != x) h->small_y.upload(y, m);
    if (h->small_out.count < m) h->small_out.alloc(m);
    if (y && y != x) h->eval_dev(MODE_ORDERED, h->small_x.ptr, h->small_y.ptr, 1, 1, h->small_out.ptr, 1, nullptr);
    else h->eval_dev(MODE_QUAD, h->small_x.ptr, nullptr, 1, 1, h->small_out.ptr, 1, nullptr);
    CA_CUDA(cudaMemcpy(out, h->small_out.ptr, m * sizeof(double), cudaMemcpyDeviceToHost));
  });
}

int32_t ca_bilinear_eval_batched(ca_bilinear* h, const double* X, int64_t nstates, double* OUT) {
  return guarded([&] {
    if (!h || (nstates > 0 && (!X || !OUT))) fail(CA_ERR_INVALID, "ca_bilinear_eval_batched: null argument");
    if (nstates < 0) fail(CA_ERR_INVALID, "nstates must be non-negative");
    h->bind();
    h->need_streams();
    eval_batched_host([&](const double* dX, int64_t n, double* dOUT, cudaStream_t st) {
      h->eval_dev(MODE_QUAD, dX, nullptr, n, n, dOUT, n, st);
    }, h->streams, h->stage_in, h->stage_out, X, nstates, OUT, h->m, h->m);
  });
}

int32_t ca_bilinear_eval_batched_dev(ca_bilinear* h, const double* dX, int64_t nstates, int64_t ldx,
                                     double* dOUT, int64_t ldo, void* stream) {
  return guarded([&] {
    if (!h || (nstates > 0 && (!dX || !dOUT))) fail(CA_ERR_INVALID, "null argument");
    h->bind();
    h->eval_dev(MODE_QUAD, dX, nullptr, nstates, ldx, dOUT, ldo, (cudaStream_t)stream);
  });
}

int32_t ca_bilinear_eval2_batched_dev(ca_bilinear* h, const double* dX, const double* dY,
                                      int64_t nstates, int64_t ld, double* dOUT, int64_t ldo,
                                      void* stream) {
  return guarded([&] {
    if (!h || (nstates > 0 && (!dX || !dY || !dOUT))) fail(CA_ERR_INVALID, "null argument");
    h->bind();
    h->eval_dev(MODE_ORDERED, dX, dY, nstates, ld, dOUT, ldo, (cudaStream_t)stream);
  });
}

int32_t ca_bilinear_jvp_batched_dev(ca_bilinear* h, const double* dX, const double* dV,
                                    int64_t nstates, int64_t ld, double* dOUT, int64_t ldo,
                                    void* stream) {
  return guarded([&] {
    if (!h || (nstates > 0 && (!dX || !dV || !dOUT))) fail(CA_ERR_INVALID, "null argument");
    h->bind();
    h->eval_dev(MODE_JVP, dX, dV, nstates, ld, dOUT, ldo, (cudaStream_t)stream);
  });
}

int32_t ca_bilinear_jacobian_dev(ca_bilinear* h, const double* dx, double* dDN, void* stream) {
  return guarded([&] {
    if (!h || !dx || !dDN) fail(CA_ERR_INVALID, "null argument");
    h->bind();
    launch_jacobian(h->jacobian(), dx, dDN, (cudaStream_t)stream);
  });
}

int32_t ca_bilinear_jacobian_batched_dev(ca_bilinear* h, const double* dX, int64_t nstates, int64_t ldx, double* dDN,
                                         int64_t ldo, void* stream) {
  return guarded([&] {
    if (!h || (nstates > 0 && (!dX || !dDN))) fail(CA_ERR_INVALID, "null argument");
    h->bind();
    launch_jacobian_batched(h->jacobian(), dX, nstates, ldx, nullptr, nullptr, 0.0, dDN, ldo, (cudaStream_t)stream);
  });
}

int32_t ca_bilinear_jacobian_batched(ca_bilinear* h, const double* X, int64_t nstates, double* DN) {
  return guarded([&] {
    if (!h || (nstates > 0 && (!X || !DN))) fail(CA_ERR_INVALID, "null argument");
    if (nstates <= 0) return;
    h->bind();
    const size_t m = (size_t)h->m, n = (size_t)nstates;
    DeviceBuffer<double> dX, dD;
    dX.upload(X, n * m);
    dD.alloc(n * m * m);
    launch_jacobian_batched(h->jacobian(), dX.ptr, nstates, nstates, nullptr, nullptr, 0.0, dD.ptr, nstates, nullptr);
    CA_CUDA(cudaMemcpy(DN, dD.ptr, n * m * m * sizeof(double), cudaMemcpyDeviceToHost));
  });
}

int32_t ca_bilinear_jacobian(ca_bilinear* h, const double* x, double* DN) {
  return guarded([&] {
    if (!h || !x || !DN) fail(CA_ERR_INVALID, "null argument");
    h->bind();
    const size_t m = (size_t)h->m;
    h->small_x.upload(x, m);
    if (h->jac_out.count < m * m) h->jac_out.alloc(m * m);
    launch_jacobian(h->jacobian(), h->small_x.ptr, h->jac_out.ptr, nullptr);
    CA_CUDA(cudaMemcpy(DN, h->jac_out.ptr, m * m * sizeof(double), cudaMemcpyDeviceToHost));
  });
}

}  // extern "C"

extern "C" int32_t ca_selftest_choose_parts(int64_t nblocks, int64_t slots, int32_t max_parts, int32_t* parts) {
  return guarded([&] {
    if (!parts) fail(CA_ERR_INVALID, "ca_selftest_choose_parts: parts is null");
    *parts = choose_parts(nblocks, slots, max_parts);
  });
}

extern "C" int32_t ca_selftest_pack_digest(const int64_t* ijk, const double* val, int64_t nnz, int64_t m,
                                           int32_t symmetric, int32_t window_modes, uint64_t* digest8) {
  return guarded([&] {
    if (!digest8 || m <= 0 || (nnz > 0 && (!ijk || !val))) fail(CA_ERR_INVALID, "ca_selftest_pack_digest: bad arguments");
    std::vector<Term> coo((size_t)nnz);
    for (int64_t r = 0; r < nnz; ++r)
      coo[(size_t)r] = Term{(int32_t)(ijk[r] - 1), (int32_t)(ijk[r + nnz] - 1), (int32_t)(ijk[r + 2 * nnz] - 1), val[r]};
    const PackedHost P = pack_terms(std::move(coo), (int32_t)m, symmetric != 0, nt_max_for(window_modes), window_modes);
    pack_digest(P, digest8);
  });
}

extern "C" int32_t ca_selftest_pack(const int64_t* ijk, const double* val, int64_t nnz, int64_t m,
                                    int32_t symmetric, int32_t window_modes, int64_t* ntiles, int64_t* nterms,
                                    int64_t* padded_fma, int64_t* pair_products, int64_t* mismatches,
                                    int64_t* nwindows, int64_t* max_window_modes) {
  return guarded([&] {
    std::vector<Term> coo((size_t)nnz);
    for (int64_t r = 0; r < nnz; ++r)
      coo[(size_t)r] = Term{(int32_t)(ijk[r] - 1), (int32_t)(ijk[r + nnz] - 1), (int32_t)(ijk[r + 2 * nnz] - 1), val[r]};
    PackedHost P = pack_terms(coo, (int32_t)m, symmetric != 0, nt_max_for(window_modes), window_modes);
    // expected: merged terms
    struct K { int32_t i, a, b; double v; };
    auto less = [](const K& x, const K& y) {
      if (x.i != y.i) return x.i < y.i;
      if (x.a != y.a) return x.a < y.a;
      return x.b < y.b;
    };
    std::vector<K> want;
    for (auto t : coo) {
      if (symmetric && t.a > t.b) std::swap(t.a, t.b);
      want.push_back({t.i, t.a, t.b, t.v});
    }
    std::sort(want.begin(), want.end(), less);
    std::vector<K> wm;
    for (size_t r = 0; r < want.size();) {
      size_t e = r;
      double s = 0;
      while (e < want.size() && !less(want[r], want[e]) && !less(want[e], want[r])) s += want[e++].v;
      if (s != 0.0) wm.push_back({want[r].i, want[r].a, want[r].b, s});
      r = e;
    }
    std::vector<K> got;
    std::vector<char> row_seen((size_t)m, 0);
    int64_t bad = 0;
    // global mode index of both factors of the pair at (tile, position k)
    auto factors = [&](size_t tile_index, const TileDesc& T, int k, int32_t& a, int32_t& b) {
      const uint32_t pr = P.pairs[(size_t)T.ks_begin * 4 + k];
      a = (int32_t)(pr & 0x7fff);
      b = (int32_t)(pr >> 16);
      if (window_modes > 0) {
        const int ks = T.ks_begin + k / 4;
        const WindowDesc* W = nullptr;
        for (const WindowDesc& w : P.windows)
          if (w.tile == (int32_t)tile_index && ks >= w.ks_begin && ks < w.ks_begin + w.nks) W = &w;
        if (!W || a >= W->nmodes || b >= W->nmodes) { ++bad; a = b = 0; return; }
        a = P.modes[(size_t)W->modes_off + a];
        b = P.modes[(size_t)W->modes_off + b];
      }
    };
    for (size_t ti = 0; ti < P.tiles.size(); ++ti) {
      const TileDesc& T = P.tiles[ti];
      std::vector<int32_t> fa(4 * (size_t)T.nks), fb(4 * (size_t)T.nks);
      for (int k = 0; k < 4 * T.nks; ++k) factors(ti, T, k, fa[k], fb[k]);
      for (int n = 0; n < 8 * T.nt; ++n) {
        const int32_t row = P.rows[(size_t)T.rows_off + n];
        if (row < 0) continue;
        if (row_seen[(size_t)row]) ++bad;  // every output row is owned by exactly one tile
        row_seen[(size_t)row] = 1;
        for (int k = 0; k < 4 * T.nks; ++k) {
          const double v = P.vals[(size_t)T.vals_off + ((size_t)(k / 4) * T.nt + n / 8) * 32 + (n % 8) * 4 + k % 4];
          if (v != 0.0) got.push_back({row, fa[k], fb[k], v});
        }
      }
    }
    if (window_modes > 0) {  // windows tile the k-steps of their tile, in order, and respect the mode cap
      size_t w = 0;
      for (size_t ti = 0; ti < P.tiles.size(); ++ti) {
        int ks = P.tiles[ti].ks_begin;
        bool last_seen = false;
        for (; w < P.windows.size() && P.windows[w].tile == (int32_t)ti; ++w) {
          const WindowDesc& W = P.windows[w];
          const TileDesc& T = P.tiles[ti];
          bad += W.ks_begin != ks || W.nmodes > window_modes || last_seen;
          // what the kernel stages per window: at most kWindowMaxKs k-steps of pair words, and the tile's value /
          // row offsets repeated in the descriptor
          bad += W.nks > kWindowMaxKs || W.nt != T.nt || W.rows_off != T.rows_off ||
                 W.vals_off != T.vals_off + (int64_t)(W.ks_begin - T.ks_begin) * T.nt * 32;
          ks += P.windows[w].nks;
          last_seen = P.windows[w].last != 0;
        }
        bad += ks != P.tiles[ti].ks_begin + P.tiles[ti].nks || !last_seen;
      }
      bad += w != P.windows.size();
    }
    for (int64_t r = 0; r < m; ++r) bad += !row_seen[(size_t)r];
    std::sort(got.begin(), got.end(), less);
    if (got.size() != wm.size()) bad += (int64_t)std::max(got.size(), wm.size()) - (int64_t)std::min(got.size(), wm.size());
    for (size_t r = 0, shown = 0; r < std::min(got.size(), wm.size()); ++r)
      if (getenv("CA_DEBUG_PACK") && shown < 10 && (got[r].i != wm[r].i || got[r].a != wm[r].a || got[r].b != wm[r].b || got[r].v != wm[r].v)) { ++shown;
        fprintf(stderr, "r=%zu got (%d,%d,%d,%g) want (%d,%d,%d,%g) sizes %zu %zu\n", r, got[r].i, got[r].a, got[r].b, got[r].v, wm[r].i, wm[r].a, wm[r].b, wm[r].v, got.size(), wm.size()); }
    for (size_t r = 0; r < std::min(got.size(), wm.size()); ++r)
      bad += got[r].i != wm[r].i || got[r].a != wm[r].a || got[r].b != wm[r].b || got[r].v != wm[r].v;
    {  // the warp schedule covers every tile exactly once, nt == 2 tiles first, and its stream lengths add up
      std::vector<int> seen(P.tiles.size(), 0);
      for (int w = 0; w < 8; ++w) {
        int64_t ks2 = 0, ks1 = 0;
        for (int q = P.sched_off[w]; q < P.sched_off[w + 1]; ++q) {
          const int32_t t = P.sched[(size_t)q];
          if (t < 0 || t >= (int32_t)P.tiles.size()) { ++bad; continue; }
          seen[(size_t)t]++;
          const bool first_part = q < P.sched_off[w] + P.sched_nt2[w];
          bad += first_part != (P.tiles[(size_t)t].nt >= 2);  // (three n-tiles only occur in windowed packs, which do not use the schedule)
          (P.tiles[(size_t)t].nt == 2 ? ks2 : ks1) += (P.tiles[(size_t)t].nks + kStreamAlign - 1) / kStreamAlign * kStreamAlign;
          if (P.stream_layout) bad += P.tiles[(size_t)t].ks_begin % kStreamAlign != 0;
        }
        if (P.stream_layout) bad += ks2 != P.sched_ks2[w] || ks1 != P.sched_ks1[w];
      }
      for (int v : seen) bad += v != 1;
    }
    if (ntiles) *ntiles = (int64_t)P.tiles.size();
    if (nterms) *nterms = P.nterms;
    if (padded_fma) *padded_fma = P.padded_fma;
    if (pair_products) *pair_products = P.pair_products;
    if (mismatches) *mismatches = bad;
    if (nwindows) *nwindows = (int64_t)P.windows.size();
    if (max_window_modes) *max_window_modes = P.max_window_modes;
  });
}
