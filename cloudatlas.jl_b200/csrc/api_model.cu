// C ABI: ODEModel evaluation side (reference: src/ODEModels.jl:54-58).
#include <algorithm>
#include <cmath>
#include <atomic>
#include <functional>
#include <thread>
#include <map>
#include <memory>
#include <numeric>
#include <string>
#include <vector>

#include "assembly.h"
#include "bilinear.h"
#include "bilinear_kernels.cuh"
#include "dense_lu.cuh"
#include "model.h"

using namespace ca;


namespace {

// vals[pos[t]] = l1[t] + l2[t] * invR : re-targets the linear slots of the packed operator to a new R
__global__ void patch_linear_kernel(double* __restrict__ vals, const int64_t* __restrict__ pos,
                                    const double* __restrict__ l1, const double* __restrict__ l2,
                                    int64_t n, double invR) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) vals[pos[t]] = fma(l2[t], invR, l1[t]);
}

// out (column-major m x m) = in (row-major m x m), i.e. a plain transpose of the memory image
__global__ void transpose_kernel(const double* __restrict__ in, double* __restrict__ out, int m) {
  __shared__ double tile[32][33];
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8)
    if (y0 + r < m && x0 + threadIdx.x < m) tile[r][threadIdx.x] = in[(size_t)(y0 + r) * m + x0 + threadIdx.x];
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8)
    if (x0 + r < m && y0 + threadIdx.x < m) out[(size_t)(x0 + r) * m + y0 + threadIdx.x] = tile[threadIdx.x][r];
}

// Df[i,j] += L1[i,j] + L2[i,j] * invR
__global__ void add_linear_kernel(double* __restrict__ Df, const double* __restrict__ L1,
                                  const double* __restrict__ L2, int64_t n, double invR) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) Df[t] += fma(L2[t], invR, L1[t]);
}

// In-place LU with partial pivoting of a dense g x g block (row-major), then inverse by solving for
// the identity -- the per-block equivalent of  Bfact = lu(B); Bfact \ I  (ODEModels.jl:54).
void invert_block(std::vector<double>& A, int g, std::vector<double>& inv) {
  std::vector<int> piv(g);
  for (int c = 0; c < g; ++c) {
    int p = c;
    for (int r = c + 1; r < g; ++r)
      if (std::fabs(A[r * g + c]) > std::fabs(A[p * g + c])) p = r;
    if (A[p * g + c] == 0.0) fail(CA_ERR_INVALID, "ca_model_create: B is singular");
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
}

}  // namespace

void ca_model::set_R(double R, cudaStream_t stream) {
    if (!(R != 0.0)) fail(CA_ERR_INVALID, "R must be non-zero");
    const double invR = 1.0 / R;
    if (invR == cur_invR) return;
    if (nnz_lin > 0) {
      CA_CUDA(cudaDeviceSynchronize());
      patch_linear_kernel<<<(unsigned)((nnz_lin + 255) / 256), 256, 0, stream>>>(
          op.vals.ptr, lin_pos.ptr, lin_l1.ptr, lin_l2.ptr, nnz_lin, invR);
      CA_CUDA(cudaGetLastError());
      count_launch();
      CA_CUDA(cudaStreamSynchronize(stream));
    }
    cur_invR = invR;
  }

void ca_model::eval_dev(int mode, const double* dX, const double* dV, int64_t nstates, int64_t ldx, double R,
                        double* dOUT, int64_t ldo, cudaStream_t st) {
  if (!(R != 0.0)) fail(CA_ERR_INVALID, "R must be non-zero");
  if (mode == MODE_QUAD && m <= kJitMaxM && nstates >= kJitMinStates && !small.failed) {
    if (!small_tried) {
      small_tried = true;
      host_linear();
      host_qterms();
      std::vector<int32_t> li, lj;
      for (int j = 0; j < (int)m; ++j)
        for (int i = 0; i < (int)m; ++i)
          if (hL1[i + (size_t)m * j] != 0.0 || hL2[i + (size_t)m * j] != 0.0) { li.push_back(i); lj.push_back(j); }
      small.build(merge_terms(qterms, true), li, lj, (int)m);
    }
    if (small.ready) {
      const double invR = 1.0 / R;
      if (!small.lc_set || invR != small.cur_invR) {
        std::vector<double> lc(small.lin_i.size());
        for (size_t e = 0; e < lc.size(); ++e) {
          const size_t at = small.lin_i[e] + (size_t)m * small.lin_j[e];
          lc[e] = std::fma(hL2[at], invR, hL1[at]);
        }
        CA_CUDA(cudaDeviceSynchronize());  // LC is shared by every stream using this model
        small.set_linear(lc.data(), st);
        small.cur_invR = invR;
      }
      small.launch(dX, nstates, ldx, dOUT, ldo, st);
      return;
    }
  }
  if (nstates > kStreamMaxStates && batched_fits(op, mode)) {
    set_R(R, st);
    launch_batched(op, mode, dX, dV, nstates, ldx, dOUT, ldo, st);
    return;
  }
  launch_stream(streaming(), mode, dX, dV, nstates, ldx, dOUT, ldo, L1r.ptr, L2r.ptr, 1.0 / R, st);
}

extern "C" {

}  // extern "C"

// Solver CSR of the symmetrised folded Q by output row, for the in-kernel evaluation of the device-resident
// Newton-hookstep (only models small enough for that solver: 3 m^2 doubles must fit in shared memory)
static void build_solver_csr(ca_model* h) {
  const int m = (int)h->m;
  if (m > 96) return;
  const std::vector<Term> sym = merge_terms(h->host_qterms(), true);
  std::vector<int32_t> rowptr(m + 1, 0);
  std::vector<uint32_t> ab(sym.size());
  std::vector<double> qv(sym.size());
  for (size_t e = 0; e < sym.size(); ++e) {
    ab[e] = (uint32_t)sym[e].a | ((uint32_t)sym[e].b << 16);
    qv[e] = sym[e].v;
    rowptr[sym[e].i + 1]++;
  }
  for (int i = 0; i < m; ++i) rowptr[i + 1] += rowptr[i];
  h->q_rowptr.upload(rowptr.data(), rowptr.size());
  h->q_ab.upload(ab.data(), ab.size());
  h->q_val.upload(qv.data(), qv.size());
  CA_CUDA(cudaStreamSynchronize(nullptr));
}

bool build_model_on_device(ca_model* h, const double* dB, const double* dA1, const double* dA2, int64_t m64,
                           const int64_t* di, const int64_t* dj, const int64_t* dk, const double* dval, int64_t nnz,
                           cudaStream_t st);

// The host build: B^-1 folding, row clustering and fragment layout on worker threads (the round-1 path; kept for
// operators the device build does not take -- unsorted or duplicated COO, large blocks of B -- and as its checker).
static void build_model_on_host(ca_model* hmodel, const double* B, const double* A1, const double* A2, int64_t m64,
                                const int64_t* ijk, const double* val, int64_t nnz) {
  {
    ca_model* h = hmodel;
    const int m = (int)m64;
    const double t_start_host = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }();
    PhaseTimer tm("ca_model_create");
    // ---- blocks of B = connected components of its non-zero pattern
    std::vector<int> comp(m);
    std::iota(comp.begin(), comp.end(), 0);
    std::function<int(int)> find = [&](int a) { while (comp[a] != a) a = comp[a] = comp[comp[a]]; return a; };
    for (int j = 0; j < m; ++j)
      for (int i = 0; i < m; ++i)
        if (i != j && (B[i + (size_t)m * j] != 0.0 || B[j + (size_t)m * i] != 0.0)) {
          int a = find(i), b = find(j);
          if (a != b) comp[std::max(a, b)] = std::min(a, b);
        }
    std::map<int, std::vector<int>> blocks;
    for (int i = 0; i < m; ++i) blocks[find(i)].push_back(i);

    // ---- N grouped by output row
    std::vector<std::vector<std::pair<uint64_t, double>>> byrow(m);
    {
      // the reference's COO is sorted by (i, j, k) (SparseBilinear.jl:24-49): rows are contiguous ranges, filled on
      // worker threads; anything else goes through the serial scatter
      std::atomic<int> unsorted{0};
      parallel_chunks((size_t)nnz, 64, [&](size_t b, size_t e, size_t) {
        for (size_t r = std::max<size_t>(b, 1); r < e; ++r)
          if (ijk[r] < ijk[r - 1]) { unsorted = 1; return; }
      });
      const bool sorted_rows = unsorted == 0;
      if (sorted_rows && nnz > (1 << 16)) {
        std::vector<int64_t> row_start((size_t)m + 1, 0);  // sorted by i: the rows are found by bisection
        for (int i = 0; i <= m; ++i) row_start[(size_t)i] = std::lower_bound(ijk, ijk + nnz, (int64_t)i + 1) - ijk;
        parallel_chunks((size_t)m, (size_t)m, [&](size_t b, size_t e, size_t) {
          for (size_t i = b; i < e; ++i) {
            auto& row = byrow[i];
            row.reserve((size_t)(row_start[i + 1] - row_start[i]));
            for (int64_t r = row_start[i]; r < row_start[i + 1]; ++r)
              row.push_back({((uint64_t)ijk[r + nnz] - 1) * (uint64_t)m + ((uint64_t)ijk[r + 2 * nnz] - 1), val[r]});
          }
        });
      } else {
        for (int64_t r = 0; r < nnz; ++r) {
          const int i = (int)ijk[r] - 1;
          const uint64_t j = (uint64_t)ijk[r + nnz] - 1, k = (uint64_t)ijk[r + 2 * nnz] - 1;
          byrow[i].push_back({j * (uint64_t)m + k, val[r]});
        }
      }
    }
    tm.lap("blocks+byrow");
    std::vector<Term> qterms;  // folded Q = B^-1 N, ordered (j,k)
    std::vector<double> hL1((size_t)m * m, 0.0), hL2((size_t)m * m, 0.0);
    std::vector<const std::vector<int>*> block_list;
    for (auto& kv : blocks) block_list.push_back(&kv.second);
    std::vector<std::vector<Term>> block_terms(block_list.size());
    std::atomic<int> singular{0};
    // one block of B at a time (independent: rows of different blocks do not mix), on worker threads
    auto fold_block = [&](size_t bi) {
      const std::vector<int>& G = *block_list[bi];
      const int g = (int)G.size();
      std::vector<double> Bg((size_t)g * g), Binv;
      for (int a = 0; a < g; ++a)
        for (int b = 0; b < g; ++b) Bg[(size_t)a * g + b] = B[G[a] + (size_t)m * G[b]];
      try {
        invert_block(Bg, g, Binv);
      } catch (const Error&) {
        singular = 1;
        return;
      }
      // linear operators: rows G of B^-1 A
      for (int a = 0; a < g; ++a)
        for (int j = 0; j < m; ++j) {
          double s1 = 0, s2 = 0;
          for (int b = 0; b < g; ++b) {
            s1 += Binv[(size_t)a * g + b] * A1[G[b] + (size_t)m * j];
            s2 += Binv[(size_t)a * g + b] * A2[G[b] + (size_t)m * j];
          }
          hL1[G[a] + (size_t)m * j] = s1;
          hL2[G[a] + (size_t)m * j] = s2;
        }
      // quadratic operator: dense [g x |union of pairs|] times Binv
      std::vector<uint64_t> U;
      bool rows_sorted = true;  // (j,k)-sorted rows (the reference's COO order): the union is a sequence of merges
      for (int b = 0; b < g && rows_sorted; ++b) {
        const auto& row = byrow[G[b]];
        for (size_t e = 1; e < row.size() && rows_sorted; ++e) rows_sorted = row[e - 1].first <= row[e].first;
      }
      if (rows_sorted) {
        std::vector<uint64_t> keys, merged;
        for (int b = 0; b < g; ++b) {
          keys.clear();
          for (auto& e : byrow[G[b]])
            if (keys.empty() || keys.back() != e.first) keys.push_back(e.first);
          merged.clear();
          std::set_union(U.begin(), U.end(), keys.begin(), keys.end(), std::back_inserter(merged));
          U.swap(merged);
        }
      } else {
        for (int b = 0; b < g; ++b)
          for (auto& e : byrow[G[b]]) U.push_back(e.first);
        std::sort(U.begin(), U.end());
        U.erase(std::unique(U.begin(), U.end()), U.end());
      }
      if (U.empty()) return;
      std::vector<double> D((size_t)g * U.size(), 0.0);
      for (int b = 0; b < g; ++b) {
        size_t p = 0;
        for (auto& e : byrow[G[b]]) {
          if (rows_sorted) {
            while (U[p] != e.first) ++p;  // both ascending; the row's keys are a subset of U
          } else {
            p = std::lower_bound(U.begin(), U.end(), e.first) - U.begin();
          }
          D[(size_t)b * U.size() + p] += e.second;
        }
      }
      // row a of B^-1 times D, one row of the result at a time: the pair index runs innermost (contiguous, vectorised);
      // every (a, p) still sums over b in ascending order, and skipped zeros of B^-1 or D only ever added exact zeros
      std::vector<Term>& out_terms = block_terms[bi];
      const size_t np = U.size();
      std::vector<double> Rm((size_t)g * np, 0.0);
      constexpr size_t kChunk = 1024;  // pairs per pass: the g rows of D and of the result stay in cache (8 KB each)
      for (size_t p0 = 0; p0 < np; p0 += kChunk) {
        const size_t p1 = std::min(np, p0 + kChunk);
        for (int a = 0; a < g; ++a) {
          double* Ra = Rm.data() + (size_t)a * np;
          for (int b = 0; b < g; ++b) {
            const double w = Binv[(size_t)a * g + b];
            if (w == 0.0) continue;
            const double* Db = D.data() + (size_t)b * np;
            for (size_t p = p0; p < p1; ++p) Ra[p] += w * Db[p];
          }
        }
      }
      {
        size_t keep = 0;
        for (double v : Rm) keep += v != 0.0;
        out_terms.reserve(keep);
      }
      std::vector<int32_t> uj(np), uk(np);  // the pair of every column, split once (64-bit divisions)
      for (size_t p = 0; p < np; ++p) {
        uj[p] = (int32_t)(U[p] / m);
        uk[p] = (int32_t)(U[p] % m);
      }
      for (int a = 0; a < g; ++a) {
        const double* Ra = Rm.data() + (size_t)a * np;
        for (size_t p = 0; p < np; ++p)
          if (Ra[p] != 0.0) out_terms.push_back(Term{G[a], uj[p], uk[p], Ra[p]});
      }
    };
    {
      std::atomic<size_t> next{0};
      auto work = [&]() {
        for (size_t bi = next.fetch_add(1); bi < block_list.size(); bi = next.fetch_add(1)) fold_block(bi);
      };
      const unsigned nthreads = ca::worker_threads();
      std::vector<std::thread> pool;
      for (unsigned t = 1; t < nthreads; ++t) pool.emplace_back(work);
      work();
      for (auto& th : pool) th.join();
    }
    if (singular) fail(CA_ERR_INVALID, "ca_model_create: B is singular");
    {
      size_t total = 0;
      for (auto& v : block_terms) total += v.size();
      std::vector<size_t> at(block_terms.size() + 1, 0);
      for (size_t bi = 0; bi < block_terms.size(); ++bi) at[bi + 1] = at[bi] + block_terms[bi].size();
      qterms.resize(total);  // (Term() does not zero-fill) the blocks are copied into place on worker threads
      parallel_chunks(block_terms.size(), block_terms.size(), [&](size_t b, size_t e, size_t) {
        for (size_t bi = b; bi < e; ++bi) {
          std::copy(block_terms[bi].begin(), block_terms[bi].end(), qterms.begin() + (ptrdiff_t)at[bi]);
          std::vector<Term>().swap(block_terms[bi]);
        }
      });
    }
    tm.lap("fold B^-1");
    h->nblocks = (int64_t)blocks.size();
    // linear placeholders (value 1) mark the slots that set_R patches
    std::vector<Term> all;
    all.reserve(qterms.size() + (size_t)m * m);  // one allocation: growing a multi-gigabyte copy by push_back copies it again
    all.resize(qterms.size());
    parallel_chunks(qterms.size(), 64, [&](size_t b, size_t e, size_t) {
      std::copy(qterms.begin() + (ptrdiff_t)b, qterms.begin() + (ptrdiff_t)e, all.begin() + (ptrdiff_t)b);
    });
    for (int j = 0; j < m; ++j)
      for (int i = 0; i < m; ++i)
        if (hL1[i + (size_t)m * j] != 0.0 || hL2[i + (size_t)m * j] != 0.0) all.push_back(Term{i, j, m, 1.0});
    tm.lap("term list");
    PackedHost P = pack_terms(std::move(all), m, /*symmetric=*/true, nt_max_for(window_modes_for(m)), window_modes_for(m));
    h->qterms = std::move(qterms);
    tm.lap("pack");
    std::vector<int64_t> pos;
    std::vector<double> l1, l2;
    // global (a, b) of the pair at position k of a tile; windowed packs store window-local indices
    std::vector<const WindowDesc*> win_of_kstep;
    if (P.window_modes > 0) {
      win_of_kstep.assign(P.pairs.size() / 4, nullptr);
      for (const WindowDesc& W : P.windows)
        for (int ks = 0; ks < W.nks; ++ks) win_of_kstep[(size_t)W.ks_begin + ks] = &W;
    }
    for (const TileDesc& T : P.tiles)
      for (int k = 0; k < 4 * T.nks; ++k) {
        const uint32_t pr = P.pairs[(size_t)T.ks_begin * 4 + k];
        int a = pr & 0x7fff, b = pr >> 16;
        if (P.window_modes > 0) {
          const WindowDesc* W = win_of_kstep[(size_t)T.ks_begin + k / 4];
          a = P.modes[(size_t)W->modes_off + a];
          b = P.modes[(size_t)W->modes_off + b];
        }
        if (b != m || a == m) continue;
        for (int n = 0; n < T.nrows; ++n) {
          const size_t at = (size_t)T.vals_off + ((size_t)(k / 4) * T.nt + n / 8) * 32 + (n % 8) * 4 + k % 4;
          if (P.vals[at] == 0.0) continue;
          const int row = P.rows[(size_t)T.rows_off + n];
          pos.push_back((int64_t)at);
          l1.push_back(hL1[row + (size_t)m * a]);
          l2.push_back(hL2[row + (size_t)m * a]);
        }
      }
    tm.lap("slots");
    h->nnz_lin = (int64_t)pos.size();
    h->op.upload(P);
    h->lin_pos.upload(pos.data(), pos.size());
    h->lin_l1.upload(l1.data(), l1.size());
    h->lin_l2.upload(l2.data(), l2.size());
    h->L1.upload(hL1.data(), hL1.size());
    h->L2.upload(hL2.data(), hL2.size());
    h->hL1 = hL1;
    h->hL2 = hL2;
    h->hB.assign(B, B + (size_t)m * m);
    h->B_on_device = false;
    h->dB.release();
    h->hA1.assign(A1, A1 + (size_t)m * m);
    h->hA2.assign(A2, A2 + (size_t)m * m);
    h->ncoo.resize((size_t)nnz);
    {
      ca_model* hp = h;
      parallel_chunks((size_t)nnz, 64, [&](size_t b, size_t e, size_t) {
        for (size_t r = b; r < e; ++r)
          hp->ncoo[r] = Term{(int32_t)(ijk[r] - 1), (int32_t)(ijk[r + nnz] - 1), (int32_t)(ijk[r + 2 * nnz] - 1), val[r]};
      });
    }
    {
      std::vector<double> r1((size_t)m * m), r2((size_t)m * m);
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) {
          r1[(size_t)i * m + j] = hL1[i + (size_t)m * j];
          r2[(size_t)i * m + j] = hL2[i + (size_t)m * j];
        }
      h->L1r.upload(r1.data(), r1.size());
      h->L2r.upload(r2.data(), r2.size());
      CA_CUDA(cudaStreamSynchronize(nullptr));
    }
    CA_CUDA(cudaStreamSynchronize(nullptr));
    tm.lap("upload");
    h->build_stats = ModelBuildStats();
    h->build_stats.total = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }() - t_start_host;
    h->build_stats.ntiles = (int64_t)P.tiles.size();
    h->build_stats.nterms = P.nterms;
  }
}

// column-major m x m device copy of a host matrix etc.: the host entry point feeds the device build
extern "C" {

int32_t ca_model_create(const double* B, const double* A1, const double* A2, int64_t m64,
                        const int64_t* ijk, const double* val, int64_t nnz, ca_model** out) {
  return guarded([&] {
    if (!out) fail(CA_ERR_INVALID, "ca_model_create: out is null");
    *out = nullptr;
    if (m64 <= 0 || !B || !A1 || !A2) fail(CA_ERR_INVALID, "ca_model_create: bad operator arguments");
    if (nnz < 0 || (nnz > 0 && (!ijk || !val))) fail(CA_ERR_INVALID, "ijk and val must have same number of rows");
    require_device();
    std::unique_ptr<ca_model> h(new ca_model);
    CA_CUDA(cudaGetDevice(&h->device));
    h->m = m64;
    bool built = false;
    if (!getenv("CA_BUILD_HOST") && m64 + 1 <= 32768) {
      const size_t mm = (size_t)m64 * (size_t)m64;
      DeviceBuffer<double> dB, dA1, dA2, dv;
      DeviceBuffer<int64_t> dijk;
      dB.upload(B, mm);
      dA1.upload(A1, mm);
      dA2.upload(A2, mm);
      dijk.upload(ijk, (size_t)nnz * 3);
      dv.upload(val, (size_t)nnz);
      built = build_model_on_device(h.get(), dB.ptr, dA1.ptr, dA2.ptr, m64, dijk.ptr, dijk.ptr + nnz, dijk.ptr + 2 * nnz,
                                    dv.ptr, nnz, nullptr);
      if (!built) {
        h.reset(new ca_model);
        CA_CUDA(cudaGetDevice(&h->device));
        h->m = m64;
      }
    }
    if (!built) {
      for (int64_t r = 0; r < nnz; ++r)  // (the device build checks the indices itself)
        for (int c = 0; c < 3; ++c)
          if (ijk[r + c * nnz] < 1 || ijk[r + c * nnz] > m64)
            fail(CA_ERR_INVALID, "ca_model_create: N entry %lld has an index outside 1:%lld", (long long)r + 1, (long long)m64);
      build_model_on_host(h.get(), B, A1, A2, m64, ijk, val, nnz);
    }
    build_solver_csr(h.get());
    *out = h.release();
  });
}

int32_t ca_model_create_dev(const double* dB, const double* dA1, const double* dA2, int64_t m64, const int64_t* d_i,
                            const int64_t* d_j, const int64_t* d_k, const double* d_val, int64_t nnz, void* stream,
                            ca_model** out) {
  return guarded([&] {
    if (!out) fail(CA_ERR_INVALID, "ca_model_create_dev: out is null");
    *out = nullptr;
    if (m64 <= 0 || !dB || !dA1 || !dA2) fail(CA_ERR_INVALID, "ca_model_create_dev: bad operator arguments");
    if (nnz < 0 || (nnz > 0 && (!d_i || !d_j || !d_k || !d_val))) fail(CA_ERR_INVALID, "ijk and val must have same number of rows");
    require_device();
    std::unique_ptr<ca_model> h(new ca_model);
    CA_CUDA(cudaGetDevice(&h->device));
    h->m = m64;
    cudaStream_t st = (cudaStream_t)stream;
    bool built = !getenv("CA_BUILD_HOST") &&
                 build_model_on_device(h.get(), dB, dA1, dA2, m64, d_i, d_j, d_k, d_val, nnz, st);
    if (!built) {  // operators the device build does not take: through the host build
      h.reset(new ca_model);
      CA_CUDA(cudaGetDevice(&h->device));
      h->m = m64;
      const size_t mm = (size_t)m64 * (size_t)m64;
      std::vector<double> B(mm), A1(mm), A2(mm), val((size_t)nnz);
      std::vector<int64_t> ijk((size_t)nnz * 3);
      CA_CUDA(cudaStreamSynchronize(st));
      CA_CUDA(cudaMemcpy(B.data(), dB, mm * sizeof(double), cudaMemcpyDeviceToHost));
      CA_CUDA(cudaMemcpy(A1.data(), dA1, mm * sizeof(double), cudaMemcpyDeviceToHost));
      CA_CUDA(cudaMemcpy(A2.data(), dA2, mm * sizeof(double), cudaMemcpyDeviceToHost));
      if (nnz > 0) {
        CA_CUDA(cudaMemcpy(ijk.data(), d_i, (size_t)nnz * sizeof(int64_t), cudaMemcpyDeviceToHost));
        CA_CUDA(cudaMemcpy(ijk.data() + nnz, d_j, (size_t)nnz * sizeof(int64_t), cudaMemcpyDeviceToHost));
        CA_CUDA(cudaMemcpy(ijk.data() + 2 * nnz, d_k, (size_t)nnz * sizeof(int64_t), cudaMemcpyDeviceToHost));
        CA_CUDA(cudaMemcpy(val.data(), d_val, (size_t)nnz * sizeof(double), cudaMemcpyDeviceToHost));
      }
      build_model_on_host(h.get(), B.data(), A1.data(), A2.data(), m64, ijk.data(), val.data(), nnz);
    }
    build_solver_csr(h.get());
    *out = h.release();
  });
}

// Digests of the model's device-resident build products (diagnostics: device build vs host build).  [0..7]: the packed
// operator as pack_digest(); [8], [9]: L1, L2; [10]: linear slots (sorted by position); [11]: streaming form of Q, row
// by row (built if absent).
int32_t ca_model_digest(ca_model* h, uint64_t* digest12) {
  return guarded([&] {
    if (!h || !digest12) fail(CA_ERR_INVALID, "null argument");
    h->bind();
    CA_CUDA(cudaDeviceSynchronize());
    auto fetch = [](auto& vec, const auto& buf, size_t n) {
      vec.resize(n);
      if (n) CA_CUDA(cudaMemcpy(vec.data(), buf.ptr, n * sizeof(vec[0]), cudaMemcpyDeviceToHost));
    };
    const DevicePack& op = h->op;
    PackedHost P;
    P.m = op.m;
    P.nin = op.nin;
    P.symmetric = op.symmetric;
    P.window_modes = op.window_modes;
    P.max_window_modes = op.max_window_modes;
    P.nterms = op.nterms;
    P.npairs_distinct = op.npairs_distinct;
    P.padded_fma = op.padded_fma;
    P.pair_products = op.pair_products;
    P.stream_layout = op.stream_layout;
    for (int w = 0; w < 9; ++w) P.sched_off[w] = op.sched_off[w];
    for (int w = 0; w < 8; ++w) {
      P.sched_nt2[w] = op.sched_nt2[w];
      P.sched_ks2[w] = op.sched_ks2[w];
      P.sched_ks1[w] = op.sched_ks1[w];
    }
    fetch(P.tiles, op.tiles, (size_t)op.ntiles);
    P.pairs = op.host_pairs;
    fetch(P.rows, op.rows, op.rows.count);
    fetch(P.windows, op.windows, (size_t)op.nwin);
    fetch(P.modes, op.modes, op.modes.count);
    fetch(P.sched, op.sched, op.sched.count);
    int64_t nvals = 0;
    for (const TileDesc& T : P.tiles) nvals = std::max<int64_t>(nvals, T.vals_off + (int64_t)T.nks * T.nt * 32);
    fetch(P.vals, op.vals, (size_t)nvals);
    // R-dependent linear slots: digest the pack as built (slots hold the placeholder 1) -- restore them first
    std::vector<int64_t> pos;
    std::vector<double> l1, l2;
    fetch(pos, h->lin_pos, (size_t)h->nnz_lin);
    fetch(l1, h->lin_l1, (size_t)h->nnz_lin);
    fetch(l2, h->lin_l2, (size_t)h->nnz_lin);
    for (int64_t p : pos) P.vals[(size_t)p] = 1.0;
    pack_digest(P, digest12);
    std::vector<double> M;
    fetch(M, h->L1, (size_t)h->m * h->m);
    digest12[8] = fnv1a(M.data(), M.size() * sizeof(double));
    fetch(M, h->L2, (size_t)h->m * h->m);
    digest12[9] = fnv1a(M.data(), M.size() * sizeof(double));
    std::vector<size_t> ord(pos.size());
    std::iota(ord.begin(), ord.end(), 0);
    std::sort(ord.begin(), ord.end(), [&](size_t x, size_t y) { return pos[x] < pos[y]; });
    uint64_t hs = fnv1a(nullptr, 0);
    for (size_t o : ord) {
      hs = fnv1a(&pos[o], 8, hs);
      hs = fnv1a(&l1[o], 8, hs);
      hs = fnv1a(&l2[o], 8, hs);
    }
    digest12[10] = hs;
    const StreamPack& sp = h->streaming();
    std::vector<StreamChunk> ch;
    std::vector<uint32_t> ab;
    std::vector<double> sv;
    fetch(ch, sp.chunks, (size_t)sp.nchunks);
    fetch(ab, sp.ab, sp.ab.count);
    fetch(sv, sp.val, sp.val.count);
    hs = fnv1a(nullptr, 0);
    for (const StreamChunk& c : ch) {
      const int64_t f[2] = {c.row, c.count};
      hs = fnv1a(f, sizeof f, hs);
      hs = fnv1a(ab.data() + c.begin, (size_t)c.count * sizeof(uint32_t), hs);
      hs = fnv1a(sv.data() + c.begin, (size_t)c.count * sizeof(double), hs);
    }
    digest12[11] = hs;
  });
}

// ODEModel straight from a full assembly (ca_assemble over all rows, or the gathered result of ca_assemble_sharded)
// on the assembly's device: the operator never leaves the GPU.  Also installs the observables (shear vector, D).
int32_t ca_model_create_from_assembly(ca_assembly* a, ca_model** out) {
  return guarded([&] {
    if (!a || !out) fail(CA_ERR_INVALID, "ca_model_create_from_assembly: null argument");
    *out = nullptr;
    if (a->r0 != 0 || a->r1 != a->m) fail(CA_ERR_INVALID, "ca_model_create_from_assembly: the assembly covers rows [%lld,%lld) of %lld; gather the slabs first",
                                          (long long)a->r0, (long long)a->r1, (long long)a->m);
    a->bind();
    const int64_t m = a->m;
    const size_t mm = (size_t)m * m;
    DeviceBuffer<double> cB, cA1, cA2, cD;  // column-major copies (the slabs are row-major)
    const double* src[4] = {a->B.ptr, a->A1.ptr, a->A2.ptr, a->D.ptr};
    DeviceBuffer<double>* dst[4] = {&cB, &cA1, &cA2, &cD};
    const dim3 grid((unsigned)((m + 31) / 32), (unsigned)((m + 31) / 32));
    for (int q = 0; q < 4; ++q) {
      dst[q]->alloc(mm);
      transpose_kernel<<<grid, dim3(32, 8)>>>(src[q], dst[q]->ptr, (int)m);
      CA_CUDA(cudaGetLastError());
      count_launch();
    }
    ca_model* h = nullptr;
    const int32_t st = ca_model_create_dev(cB.ptr, cA1.ptr, cA2.ptr, m, a->oi.ptr, a->oj.ptr, a->ok.ptr, a->oval.ptr, a->nnz, nullptr, &h);
    if (st != CA_OK) fail(st, "%s", ca_last_error());
    std::unique_ptr<ca_model> guard(h);
    std::vector<double> D(mm);
    CA_CUDA(cudaMemcpy(D.data(), cD.ptr, mm * sizeof(double), cudaMemcpyDeviceToHost));
    if (a->shear.size() == (size_t)m) {
      h->obs_shear = a->shear;
      h->obs_D.swap(D);
      h->obs_set = true;
      h->obs_built = false;
    }
    *out = guard.release();
  });
}

int32_t ca_model_build_info(const ca_model* h, int32_t* on_device, double* seconds9, int64_t* nclusters,
                            int64_t* ntiles, int64_t* nterms) {
  return guarded([&] {
    if (!h) fail(CA_ERR_INVALID, "null handle");
    const ModelBuildStats& S = h->build_stats;
    if (on_device) *on_device = S.on_device;
    if (seconds9) {
      const double v[9] = {S.total, S.blocks, S.fold, S.symmetrise, S.cluster, S.tile_unions, S.layout, S.place, S.stream};
      for (int q = 0; q < 9; ++q) seconds9[q] = v[q];
    }
    if (nclusters) *nclusters = S.nclusters;
    if (ntiles) *ntiles = S.ntiles;
    if (nterms) *nterms = S.nterms;
  });
}

int32_t ca_model_destroy(ca_model* h) {
  return guarded([&] {
    if (!h) return;
    cudaSetDevice(h->device);
    delete h;
  });
}

int32_t ca_model_info(const ca_model* h, int64_t* m, int64_t* nnz_q_sym, int64_t* nnz_lin,
                      int64_t* npairs, int64_t* padded_fma_per_state, int64_t* nblocks_B) {
  return guarded([&] {
    if (!h) fail(CA_ERR_INVALID, "null handle");
    if (m) *m = h->m;
    if (nnz_q_sym) *nnz_q_sym = h->op.nterms - h->nnz_lin;
    if (nnz_lin) *nnz_lin = h->nnz_lin;
    if (npairs) *npairs = h->op.npairs_distinct;
    if (padded_fma_per_state) *padded_fma_per_state = h->op.padded_fma;
    if (nblocks_B) *nblocks_B = h->nblocks;
  });
}

int32_t ca_model_rhs_batched_dev(ca_model* h, const double* dX, int64_t nstates, int64_t ldx, double R,
                                 double* dOUT, int64_t ldo, void* stream) {
  return guarded([&] {
    if (!h || (nstates > 0 && (!dX || !dOUT))) fail(CA_ERR_INVALID, "null argument");
    h->bind();
    h->eval_dev(MODE_QUAD, dX, nullptr, nstates, ldx, R, dOUT, ldo, (cudaStream_t)stream);
  });
}

int32_t ca_model_jvp_batched_dev(ca_model* h, const double* dX, const double* dV, int64_t nstates,
                                 int64_t ld, double R, double* dOUT, int64_t ldo, void* stream) {
  return guarded([&] {
    if (!h || (nstates > 0 && (!dX || !dV || !dOUT))) fail(CA_ERR_INVALID, "null argument");
    h->bind();
    h->eval_dev(MODE_JVP, dX, dV, nstates, ld, R, dOUT, ldo, (cudaStream_t)stream);
  });
}

int32_t ca_model_rhs_batched(ca_model* h, const double* X, int64_t nstates, double R, double* OUT) {
  return guarded([&] {
    if (!h || (nstates > 0 && (!X || !OUT))) fail(CA_ERR_INVALID, "null argument");
    if (nstates < 0) fail(CA_ERR_INVALID, "nstates must be non-negative");
    h->bind();
    h->need_streams();
    h->set_R(R, h->streams[0]);  // before the two streams start, so no re-targeting happens mid-flight
    eval_batched_host([&](const double* dX, int64_t n, double* dOUT, cudaStream_t st) {
      h->eval_dev(MODE_QUAD, dX, nullptr, n, n, R, dOUT, n, st);
    }, h->streams, h->stage_in, h->stage_out, X, nstates, OUT, h->m, h->m);
  });
}

// f(x_s, R) of one host ensemble on several devices of this process: models[d] is the same model built on device d
// (ca_assemble_sharded + ca_model_create_from_assembly); the states are split evenly, every device runs the
// host-pointer pipeline of ca_model_rhs_batched on its range from its own host thread.  No collective: outputs are
// disjoint ranges of OUT ("ensemble RHS shards by state").
int32_t ca_models_rhs_batched(ca_model** models, int32_t nmodels, const double* X, int64_t nstates, double R, double* OUT) {
  return guarded([&] {
    if (!models || nmodels < 1 || (nstates > 0 && (!X || !OUT))) fail(CA_ERR_INVALID, "ca_models_rhs_batched: bad arguments");
    for (int d = 0; d < nmodels; ++d)
      if (!models[d] || models[d]->m != models[0]->m) fail(CA_ERR_INVALID, "ca_models_rhs_batched: models must be replicas of one model");
    if (nstates <= 0) return;
    std::vector<int32_t> codes((size_t)nmodels, CA_OK);
    std::vector<std::string> errors((size_t)nmodels);
    std::vector<std::thread> pool;
    // ranges in whole 32-state tiles so that only the last device sees a ragged tail
    const int64_t tiles = (nstates + 31) / 32;
    for (int d = 0; d < nmodels; ++d) {
      const int64_t s0 = std::min(nstates, tiles * d / nmodels * 32), s1 = std::min(nstates, tiles * (d + 1) / nmodels * 32);
      if (s1 <= s0) continue;
      pool.emplace_back([&, d, s0, s1] {
        try {
          ca_model* h = models[d];
          h->bind();
          h->need_streams();
          h->set_R(R, h->streams[0]);
          eval_batched_host([&](const double* dX, int64_t n, double* dOUT, cudaStream_t st) {
            h->eval_dev(MODE_QUAD, dX, nullptr, n, n, R, dOUT, n, st);
          }, h->streams, h->stage_in, h->stage_out, X + s0, s1 - s0, OUT + s0, h->m, h->m, nstates);
        } catch (const Error& e) {
          codes[(size_t)d] = e.code;
          errors[(size_t)d] = e.what();
        } catch (const std::exception& e) {
          codes[(size_t)d] = CA_ERR_INVALID;
          errors[(size_t)d] = e.what();
        }
      });
    }
    for (auto& t : pool) t.join();
    for (int d = 0; d < nmodels; ++d)
      if (codes[(size_t)d] != CA_OK) fail(codes[(size_t)d], "device of model %d: %s", d, errors[(size_t)d].c_str());
  });
}

int32_t ca_model_rhs(ca_model* h, const double* x, double R, double* out) {
  return guarded([&] {
    if (!h || !x || !out) fail(CA_ERR_INVALID, "null argument");
    h->bind();
    const size_t m = (size_t)h->m;
    h->small_x.upload(x, m);
    if (h->small_out.count < m) h->small_out.alloc(m);
    h->eval_dev(MODE_QUAD, h->small_x.ptr, nullptr, 1, 1, R, h->small_out.ptr, 1, nullptr);
    CA_CUDA(cudaMemcpy(out, h->small_out.ptr, m * sizeof(double), cudaMemcpyDeviceToHost));
  });
}

int32_t ca_model_jac_dev(ca_model* h, const double* dx, double R, double* dDf, void* stream) {
  return guarded([&] {
    if (!h || !dx || !dDf) fail(CA_ERR_INVALID, "null argument");
    if (!(R != 0.0)) fail(CA_ERR_INVALID, "R must be non-zero");
    h->bind();
    cudaStream_t st = (cudaStream_t)stream;
    launch_jacobian(h->jacobian(), dx, dDf, st);
    const int64_t n = h->m * h->m;
    add_linear_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dDf, h->L1.ptr, h->L2.ptr, n, 1.0 / R);
    CA_CUDA(cudaGetLastError());
    count_launch();
  });
}

int32_t ca_model_jac_batched_dev(ca_model* h, const double* dX, int64_t nstates, int64_t ldx, double R, double* dDf,
                                 int64_t ldo, void* stream) {
  return guarded([&] {
    if (!h || (nstates > 0 && (!dX || !dDf))) fail(CA_ERR_INVALID, "null argument");
    if (!(R != 0.0)) fail(CA_ERR_INVALID, "R must be non-zero");
    h->bind();
    launch_jacobian_batched(h->jacobian(), dX, nstates, ldx, h->L1.ptr, h->L2.ptr, 1.0 / R, dDf, ldo, (cudaStream_t)stream);
  });
}

int32_t ca_model_jac_batched(ca_model* h, const double* X, int64_t nstates, double R, double* Df) {
  return guarded([&] {
    if (!h || (nstates > 0 && (!X || !Df))) fail(CA_ERR_INVALID, "null argument");
    if (!(R != 0.0)) fail(CA_ERR_INVALID, "R must be non-zero");
    if (nstates <= 0) return;
    h->bind();
    const size_t m = (size_t)h->m, n = (size_t)nstates;
    DeviceBuffer<double> dX, dD;
    dX.upload(X, n * m);
    dD.alloc(n * m * m);
    launch_jacobian_batched(h->jacobian(), dX.ptr, nstates, nstates, h->L1.ptr, h->L2.ptr, 1.0 / R, dD.ptr, nstates, nullptr);
    CA_CUDA(cudaMemcpy(Df, dD.ptr, n * m * m * sizeof(double), cudaMemcpyDeviceToHost));
  });
}

int32_t ca_model_jac(ca_model* h, const double* x, double R, double* Df) {
  return guarded([&] {
    if (!h || !x || !Df) fail(CA_ERR_INVALID, "null argument");
    h->bind();
    const size_t m = (size_t)h->m;
    h->small_x.upload(x, m);
    if (h->jac_out.count < m * m) h->jac_out.alloc(m * m);
    int32_t st = ca_model_jac_dev(h, h->small_x.ptr, R, h->jac_out.ptr, nullptr);
    if (st != CA_OK) fail(st, "%s", ca_last_error());
    CA_CUDA(cudaMemcpy(Df, h->jac_out.ptr, m * m * sizeof(double), cudaMemcpyDeviceToHost));
  });
}

}  // extern "C"

// Host-only diagnostic: generates and NVRTC-compiles the model-specialised kernel for a symmetrised quadratic operator
// (+ a dense linear pattern) without creating a model; loading the cubin needs a device and is not attempted here.
extern "C" int32_t ca_selftest_small_kernel(const int64_t* ijk, const double* val, int64_t nnz, int64_t m,
                                            int64_t* cubin_bytes, int32_t* threads, int32_t* stages,
                                            int64_t* source_bytes) {
  return guarded([&] {
    if (m <= 0 || m > kJitMaxM || (nnz > 0 && (!ijk || !val))) fail(CA_ERR_INVALID, "ca_selftest_small_kernel: bad arguments");
    std::vector<Term> q((size_t)nnz);
    for (int64_t r = 0; r < nnz; ++r)
      q[(size_t)r] = Term{(int32_t)(ijk[r] - 1), (int32_t)(ijk[r + nnz] - 1), (int32_t)(ijk[r + 2 * nnz] - 1), val[r]};
    std::vector<int32_t> li, lj;
    for (int j = 0; j < (int)m; ++j)
      for (int i = 0; i < (int)m; ++i) { li.push_back(i); lj.push_back(j); }
    SmallKernel k;
    k.build(merge_terms(std::move(q), true, (int32_t)m), li, lj, (int)m);
    if (k.cubin_bytes == 0) fail(CA_ERR_UNSUPPORTED, "generated kernel did not compile: %s", k.log.c_str());
    if (cubin_bytes) *cubin_bytes = (int64_t)k.cubin_bytes;
    if (threads) *threads = k.threads;
    if (stages) *stages = k.stages;
    if (source_bytes) *source_bytes = (int64_t)k.source.size();
  });
}

extern "C" int32_t ca_model_kernel_path(ca_model* h, int64_t nstates, int32_t* path, const char** source,
                                        const char** log) {
  return guarded([&] {
    if (!h || !path) fail(CA_ERR_INVALID, "null argument");
    if (h->m <= kJitMaxM && nstates >= kJitMinStates && h->small.ready) *path = 2;
    else if (nstates > kStreamMaxStates && batched_fits(h->op, MODE_QUAD)) *path = batched_variant(h->op, nstates);
    else *path = 1;
    if (source) *source = h->small.source.c_str();
    if (log) *log = h->small.log.c_str();
  });
}

// ------------------------------------------------------------------------------------------ cnab2
namespace {

// g = c1 * n1 + c0 * n0, written into columns [m, 2m) of the stacked ensemble buffer
__global__ void cnab2_combine_kernel(const double* __restrict__ n1, const double* __restrict__ n0, double c1,
                                     double c0, int64_t count, double* __restrict__ g) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < count) g[t] = fma(c1, n1[t], c0 * n0[t]);
}

__global__ void set_identity_kernel(double* __restrict__ A, int m) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < (int64_t)m * m) A[t] = (t % m) == (t / m) ? 1.0 : 0.0;
}

void eval_pack(const DevicePack& dp, const StreamPack& sp, const double* dX, int64_t nstates, int64_t ldx,
               double* dOUT, int64_t ldo, cudaStream_t st) {
  if (nstates > kStreamMaxStates && batched_fits(dp, MODE_QUAD)) launch_batched(dp, MODE_QUAD, dX, nullptr, nstates, ldx, dOUT, ldo, st);
  else launch_stream(sp, MODE_QUAD, dX, nullptr, nstates, ldx, dOUT, ldo, nullptr, nullptr, 0.0, st);
}

}  // namespace

extern "C" int32_t ca_model_cnab2_batched_dev(ca_model* h, const double* dX0, int64_t nstates, int64_t ld, double R,
                                              double dt, int64_t nsteps, int64_t nsave, double* dXsave,
                                              double* dXfinal, void* stream) {
  return guarded([&] {
    if (!h || !dX0) fail(CA_ERR_INVALID, "null argument");
    if (nstates <= 0 || ld < nstates) fail(CA_ERR_INVALID, "bad ensemble shape");
    if (!(R != 0.0) || !(dt > 0.0) || nsteps < 0 || nsave <= 0) fail(CA_ERR_INVALID, "bad time-stepping parameters");
    h->bind();
    cudaStream_t st = (cudaStream_t)stream;
    const int m = (int)h->m;
    const bool few = nstates <= kStreamMaxStates;
    h->host_linear();
    // ---- N(x): raw operator
    if (few ? !h->nraw_stream_built : !h->nraw_built) {
      if (few) {
        h->nraw_stream.build(h->host_ncoo(), m, true);
        h->nraw_stream_built = true;
      } else {
        h->nraw.upload(pack_terms(h->host_ncoo(), m, true, nt_max_for(window_modes_for(m)), window_modes_for(m)));
        h->nraw_built = true;
        if (!batched_fits(h->nraw, MODE_QUAD) && !h->nraw_stream_built) {
          h->nraw_stream.build(h->host_ncoo(), m, true);
          h->nraw_stream_built = true;
        }
      }
    }
    // ---- [M1 M2] for this (R, dt):  LHS = B - dt/2 A, RHS = B + dt/2 A   (ODEModels.jl:149-155).  The reference keeps
    // lu(LHS); here M2 = LHS^-1 and M1 = M2 RHS are formed once per (R, dt) ON THE DEVICE (blocked LU of dense_lu.cuh,
    // multi-right-hand-side substitution, one DGEMM) and applied every step: as a packed dense operator on the DMMA
    // ensemble kernel up to kCnDenseMax modes, as two DGEMMs beyond (the term packer is limited to 32767 inputs and
    // packing 2 m^2 terms costs more than it saves at m in the thousands).
    constexpr int kCnDenseMax = 1024;
    if (!h->cn_built || h->cn_R != R || h->cn_dt != dt) {
      CA_CUDA(cudaDeviceSynchronize());
      const size_t mm = (size_t)m * m;
      std::vector<double> lhs(mm), rhs(mm);
      const std::vector<double>& hB = h->host_B();
      for (size_t e = 0; e < mm; ++e) {
        const double a = h->hA1[e] + (1.0 / R) * h->hA2[e];
        lhs[e] = hB[e] - (dt / 2) * a;
        rhs[e] = hB[e] + (dt / 2) * a;
      }
      DeviceBuffer<double> dlhs, drhs;
      DeviceBuffer<int32_t> piv, flag;
      dlhs.upload(lhs.data(), mm, st);
      drhs.upload(rhs.data(), mm, st);
      piv.alloc((size_t)m);
      flag.alloc(1);
      h->cn_M1.alloc(mm);
      h->cn_M2.alloc(mm);
      CA_CUDA(cudaMemsetAsync(flag.ptr, 0, sizeof(int32_t), st));
      set_identity_kernel<<<(unsigned)((mm + 255) / 256), 256, 0, st>>>(h->cn_M2.ptr, m);
      CA_CUDA(cudaGetLastError());
      count_launch();
      lu::factor(dlhs.ptr, m, piv.ptr, flag.ptr, true, st);
      lu::solve_many(dlhs.ptr, m, piv.ptr, h->cn_M2.ptr, m, st);                                   // M2 = LHS^-1
      launch_dgemm(m, m, m, 1.0, colmajor(h->cn_M2.ptr, m), colmajor(drhs.ptr, m), 0.0, h->cn_M1.ptr, 1, m, st);  // M1 = M2 RHS
      int32_t sing = 0;
      CA_CUDA(cudaMemcpyAsync(&sing, flag.ptr, sizeof sing, cudaMemcpyDeviceToHost, st));
      CA_CUDA(cudaStreamSynchronize(st));
      if (sing) fail(CA_ERR_INVALID, "cnab2: B - dt/2 A is singular");
      h->cn_dense = m > kCnDenseMax || getenv("CA_CNAB2_DENSE") != nullptr;  // (test hook: the DGEMM form at any size)
      if (!h->cn_dense) {
        std::vector<double> M1(mm), M2(mm);
        CA_CUDA(cudaMemcpy(M1.data(), h->cn_M1.ptr, mm * sizeof(double), cudaMemcpyDeviceToHost));
        CA_CUDA(cudaMemcpy(M2.data(), h->cn_M2.ptr, mm * sizeof(double), cudaMemcpyDeviceToHost));
        std::vector<Term> lin;
        lin.reserve((size_t)2 * mm);
        for (int j = 0; j < m; ++j)
          for (int i = 0; i < m; ++i) {
            if (M1[i + (size_t)m * j] != 0.0) lin.push_back(Term{i, j, 2 * m, M1[i + (size_t)m * j]});
            if (M2[i + (size_t)m * j] != 0.0) lin.push_back(Term{i, m + j, 2 * m, M2[i + (size_t)m * j]});
          }
        h->cn_lin_stream.build_rect(lin, m, 2 * m);
        h->cn_lin.upload(pack_terms(std::move(lin), m, true, nt_max_for(window_modes_for(2 * m)), window_modes_for(2 * m), 2 * m));
      }
      h->cn_R = R;
      h->cn_dt = dt;
      h->cn_built = true;
    }
    // ---- buffers: stacked [x; g] ensembles (nstates x 2m), N(x_n), N(x_{n-1})
    const size_t per = (size_t)nstates * m;
    for (int b = 0; b < 2; ++b) {
      if (h->cn_buf[b].count < 2 * per) h->cn_buf[b].alloc(2 * per);
      if (h->cn_n[b].count < per) h->cn_n[b].alloc(per);
    }
    double* cur = h->cn_buf[0].ptr;
    double* nxt = h->cn_buf[1].ptr;
    double* ncur = h->cn_n[0].ptr;
    double* nprev = h->cn_n[1].ptr;
    CA_CUDA(cudaMemcpy2DAsync(cur, (size_t)nstates * 8, dX0, (size_t)ld * 8, (size_t)nstates * 8, (size_t)m,
                              cudaMemcpyDeviceToDevice, st));
    auto evalN = [&](const double* x, double* out) {
      eval_pack(h->nraw, h->nraw_stream, x, nstates, nstates, out, nstates, st);
    };
    evalN(cur, ncur);  // Nxn = N(x0); Nxn1 = N(x0)
    int64_t k = 0;
    const int64_t nsaves = nsteps / nsave;
    for (int64_t n = 1; n <= nsteps; ++n) {
      if (n > 1) {
        std::swap(ncur, nprev);  // Nxn1 <- Nxn
        evalN(cur, ncur);        // Nxn <- N(xn)
      }
      const double* n0 = n > 1 ? nprev : ncur;
      cnab2_combine_kernel<<<(unsigned)((per + 255) / 256), 256, 0, st>>>(ncur, n0, 1.5 * dt, -0.5 * dt, (int64_t)per,
                                                                        cur + per);
      CA_CUDA(cudaGetLastError());
      count_launch();
      if (h->cn_dense) {  // x_{n+1} = M1 x_n + M2 g for every state: [nstates x m] = [nstates x m] [m x m]'
        launch_dgemm((int)nstates, m, m, 1.0, colmajor(cur, nstates), transposed(colmajor(h->cn_M1.ptr, m)), 0.0, nxt, 1, nstates, st);
        launch_dgemm((int)nstates, m, m, 1.0, colmajor(cur + per, nstates), transposed(colmajor(h->cn_M2.ptr, m)), 1.0, nxt, 1, nstates, st);
      } else {
        eval_pack(h->cn_lin, h->cn_lin_stream, cur, nstates, nstates, nxt, nstates, st);
      }
      std::swap(cur, nxt);
      if (n % nsave == 0 && dXsave && k < nsaves) {
        CA_CUDA(cudaMemcpyAsync(dXsave + (size_t)k * per, cur, per * 8, cudaMemcpyDeviceToDevice, st));
        ++k;
      }
    }
    if (dXfinal) CA_CUDA(cudaMemcpyAsync(dXfinal, cur, per * 8, cudaMemcpyDeviceToDevice, st));
    CA_CUDA(cudaStreamSynchronize(st));  // the stacked buffers belong to the handle
  });
}

// ------------------------------------------------------------------------------------ observables
// The pack of the two observable rows is built on the first evaluation, not here: at m ~ 3000 packing the m^2 terms
// of D costs more than the whole device model build, and constructors set the observables unconditionally.
extern "C" int32_t ca_model_set_observables(ca_model* h, const double* shear, const double* D) {
  return guarded([&] {
    if (!h || !shear || !D) fail(CA_ERR_INVALID, "null argument");
    const size_t m = (size_t)h->m;
    h->obs_shear.assign(shear, shear + m);
    h->obs_D.assign(D, D + m * m);
    h->obs_set = true;
    h->obs_built = false;
  });
}

static void build_observables(ca_model* h) {
  if (h->obs_built) return;
  const int m = (int)h->m;
  const double* shear = h->obs_shear.data();
  const double* D = h->obs_D.data();
  std::vector<Term> t;
  t.push_back(Term{0, m, m, 1.0});  // the constant 1 of shear(x) = 1 + dot(x, shear)
  t.push_back(Term{1, m, m, 1.0});  // and of dissipation(x) = 1 + x' D x
  for (int i = 0; i < m; ++i)
    if (shear[i] != 0.0) t.push_back(Term{0, i, m, shear[i]});
  for (int j = 0; j < m; ++j)
    for (int i = 0; i < m; ++i)
      if (D[i + (size_t)m * j] != 0.0) t.push_back(Term{1, i, j, D[i + (size_t)m * j]});
  h->obs_stream.build_rect(t, 2, m);
  h->obs.upload(pack_terms(std::move(t), 2, true, nt_max_for(window_modes_for(m)), window_modes_for(m), m));
  h->obs_built = true;
}

extern "C" int32_t ca_model_observables_batched_dev(ca_model* h, const double* dX, int64_t nstates, int64_t ld,
                                                    double* dOUT, void* stream) {
  return guarded([&] {
    if (!h || !dX || !dOUT) fail(CA_ERR_INVALID, "null argument");
    if (!h->obs_set) fail(CA_ERR_INVALID, "observables not set (ca_model_set_observables)");
    h->bind();
    build_observables(h);
    eval_pack(h->obs, h->obs_stream, dX, nstates, ld, dOUT, nstates, (cudaStream_t)stream);
  });
}

// ------------------------------------------------------------------ host-pointer convenience forms
extern "C" int32_t ca_model_jvp_batched(ca_model* h, const double* X, const double* V, int64_t nstates, double R,
                                        double* OUT) {
  return guarded([&] {
    if (!h || !X || !V || !OUT || nstates <= 0) fail(CA_ERR_INVALID, "bad argument");
    h->bind();
    const size_t n = (size_t)nstates * h->m;
    DeviceBuffer<double> dX, dV, dO;
    dX.upload(X, n);
    dV.upload(V, n);
    dO.alloc(n);
    h->eval_dev(MODE_JVP, dX.ptr, dV.ptr, nstates, nstates, R, dO.ptr, nstates, nullptr);
    CA_CUDA(cudaMemcpy(OUT, dO.ptr, n * sizeof(double), cudaMemcpyDeviceToHost));
  });
}

extern "C" int32_t ca_model_cnab2_batched(ca_model* h, const double* X0, int64_t nstates, double R, double dt,
                                          int64_t nsteps, int64_t nsave, double* Xsave, double* Xfinal) {
  return guarded([&] {
    if (!h || !X0 || nstates <= 0 || nsave <= 0 || nsteps < 0) fail(CA_ERR_INVALID, "bad argument");
    h->bind();
    const size_t n = (size_t)nstates * h->m, nsaves = (size_t)(nsteps / nsave);
    DeviceBuffer<double> dX, dS, dF;
    dX.upload(X0, n);
    if (Xsave && nsaves) dS.alloc(n * nsaves);
    if (Xfinal) dF.alloc(n);
    const int32_t st = ca_model_cnab2_batched_dev(h, dX.ptr, nstates, nstates, R, dt, nsteps, nsave, dS.ptr, dF.ptr, nullptr);
    if (st != CA_OK) fail(st, "%s", ca_last_error());
    if (Xsave && nsaves) CA_CUDA(cudaMemcpy(Xsave, dS.ptr, n * nsaves * sizeof(double), cudaMemcpyDeviceToHost));
    if (Xfinal) CA_CUDA(cudaMemcpy(Xfinal, dF.ptr, n * sizeof(double), cudaMemcpyDeviceToHost));
  });
}

extern "C" int32_t ca_model_observables_batched(ca_model* h, const double* X, int64_t nstates, double* OBS) {
  return guarded([&] {
    if (!h || !X || !OBS || nstates <= 0) fail(CA_ERR_INVALID, "bad argument");
    h->bind();
    DeviceBuffer<double> dX, dO;
    dX.upload(X, (size_t)nstates * h->m);
    dO.alloc((size_t)nstates * 2);
    const int32_t st = ca_model_observables_batched_dev(h, dX.ptr, nstates, nstates, dO.ptr, nullptr);
    if (st != CA_OK) fail(st, "%s", ca_last_error());
    CA_CUDA(cudaMemcpy(OBS, dO.ptr, (size_t)nstates * 2 * sizeof(double), cudaMemcpyDeviceToHost));
  });
}
