// Matrix-free Newton-GMRES for f(x, R) = 0 at model sizes where the dense m x m Jacobian and its LU dominate the
// reference's hookstepsolve (BASELINE.json configs[4]: "batched Jacobian DQ(x) evaluation for Newton-Krylov" at
// m ~ 3000), device resident and sharded by ROW SLAB over the GPUs of a communicator (last row).
//
// One Krylov vector costs one Jacobian action  Df(x,R) v = (L1 + L2/R) v + Q(x,v) + Q(v,x)  by the streaming kernel
// (stream.cu), which is HBM-bound on the operator: with N GPUs every device holds the full model but streams only the
// rows [r N^-1 m, (r+1) N^-1 m) -- 1/N of the bytes -- and one ncclAllGather of m doubles completes the vector on every
// device.  The Arnoldi process is REPLICATED: every device runs the same orthogonalisation on the same numbers (the
// kernels are deterministic, so the replicas stay bit-identical) and no vector is ever broadcast; the host reads the
// Hessenberg column of replica 0 and drives all replicas with the same scalars.  With one GPU there is no collective.
// One process driving several devices (ca_comm_init_all) and one process per GPU (ca_comm_init_rank) use the same code:
// a process owns the replicas of its local devices.
//
// Algorithm: restarted GMRES with classical Gram-Schmidt applied twice, least squares on the small Hessenberg matrix by
// Givens rotations on the host, backtracking line search on |f|.  The system is LEFT-PRECONDITIONED with the linear
// operator  M = L1 + L2/R  (dense, factored once per solve by the blocked device LU of dense_lu.cuh):
//   M^-1 Df(x) = I + M^-1 DQ(x)
// is a bounded perturbation of the identity, whereas Df itself carries the viscous spectrum of L2/R (eigenvalues
// spread over five decades at 300+ modes) on which unpreconditioned GMRES stagnates -- measured: no convergence in
// 40 x 300 Jacobian actions at 307 / 999 / 2962 modes even one continuation step away from an equilibrium.
#include <algorithm>
#include <cmath>
#include <memory>
#include <vector>

#include "bilinear.h"
#include "bilinear_kernels.cuh"
#include "comm.h"
#include "dense_lu.cuh"
#include "dgemm.cuh"
#include "model.h"

using namespace ca;

namespace {

__global__ void axpby_kernel(double a, const double* __restrict__ x, double b, const double* __restrict__ y, int n,
                             double* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = a * x[t] + (b != 0.0 ? b * y[t] : 0.0);
}

// out[i] = V[:, i] . w for i < k: one CTA per basis vector, fixed reduction order
__global__ void __launch_bounds__(256)
multi_dot_kernel(const double* __restrict__ V, const double* __restrict__ w, int m, int k, double* __restrict__ out) {
  __shared__ double red[8];
  const int i = blockIdx.x;
  if (i >= k) return;
  const double* v = V + (size_t)i * m;
  double s = 0.0;
  for (int r = threadIdx.x; r < m; r += 256) s = fma(v[r], w[r], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int q = 0; q < 8; ++q) t += red[q];
    out[i] = t;
  }
}

// w[r] += sign * sum_i V[r, i] h[i], i ascending
__global__ void __launch_bounds__(256)
combine_kernel(const double* __restrict__ V, const double* __restrict__ h, int m, int k, double sign, double* __restrict__ w) {
  const int r = blockIdx.x * 256 + threadIdx.x;
  if (r >= m) return;
  double s = 0.0;
  for (int i = 0; i < k; ++i) s = fma(V[(size_t)i * m + r], h[i], s);
  w[r] += sign * s;
}

__global__ void identity_kernel(double* __restrict__ A, int m) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < (int64_t)m * m) A[t] = (t % m) == (t / m) ? 1.0 : 0.0;
}

struct Replica {
  ca_model* h = nullptr;
  int device = 0, rank = 0, world = 1;
  ncclComm_t comm = nullptr;
  cudaStream_t st = nullptr;
  int m = 0, slab = 0, r0 = 0, r1 = 0, restart = 0;
  DeviceBuffer<double> x, xt, fx, ft, dx, w, tmp, pre, V, hcol, padded;  // padded: world * slab doubles for the all-gather
  DeviceBuffer<double> Minv;   // the preconditioner (L1 + L2/R)^-1, explicit, one ROW per contiguous m doubles
  bool precond = false;

  void init(ca_model* h_, int rank_, int world_, ncclComm_t comm_, cudaStream_t st_, int restart_) {
    h = h_;
    device = h->device;
    rank = rank_;
    world = world_;
    comm = comm_;
    st = st_;
    m = (int)h->m;
    restart = restart_;
    slab = (m + world - 1) / world;
    r0 = std::min(m, rank * slab);
    r1 = std::min(m, r0 + slab);
    CA_CUDA(cudaSetDevice(device));
    const size_t n = (size_t)m;
    for (DeviceBuffer<double>* v : {&x, &xt, &fx, &ft, &dx, &w, &tmp, &pre}) v->alloc(n);
    V.alloc(n * (size_t)(restart + 1));
    hcol.alloc((size_t)2 * (restart + 2));
    padded.alloc((size_t)world * slab);
  }
  void bind() const { CA_CUDA(cudaSetDevice(device)); }

  // this replica's rows of f (mode QUAD) or of the Jacobian action (mode JVP) into `padded`
  void eval_slab(int mode, const double* xin, const double* vin, double R) {
    bind();
    if (r1 > r0)
      launch_stream(h->streaming(), mode, xin, vin, 1, 1, padded.ptr, 1, h->L1r.ptr, h->L2r.ptr, 1.0 / R, st, r0, r1);
  }
  void gather_issue() {  // inside an ncclGroup when a process owns several replicas
    if (world > 1)
      CA_NCCL(nccl().AllGather(padded.ptr + (size_t)rank * slab, padded.ptr, (size_t)slab, ncclFloat64, comm, st));
  }
  void copy_out(double* dst) {
    bind();
    CA_CUDA(cudaMemcpyAsync(dst, padded.ptr, (size_t)m * sizeof(double), cudaMemcpyDeviceToDevice, st));
  }
  // M^-1 for M = L1 + L2/R, explicitly (replicated on every device; O(m^3) once per solve): the transpose M' --
  // the row-major copies L1r, L2r read as column-major -- is factored by the blocked device LU and M' X = I solved
  // for all columns at once, so that X (column-major) holds M^-1 ROW by row.  Applying it is then one dot product per
  // output entry over contiguous memory (multi_dot_kernel): ~15 us at m = 2962, against 2.5 ms for the single-CTA
  // triangular substitutions.
  bool build_preconditioner(double R) {
    bind();
    const size_t nn = (size_t)m * m;
    DeviceBuffer<double> Mt;
    DeviceBuffer<int32_t> piv, flag;
    Mt.alloc(nn);
    Minv.alloc(nn);
    piv.alloc((size_t)m);
    flag.alloc(1);
    CA_CUDA(cudaMemsetAsync(flag.ptr, 0, sizeof(int32_t), st));
    axpby_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, st>>>(1.0, h->L1r.ptr, 1.0 / R, h->L2r.ptr, (int)nn, Mt.ptr);
    identity_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, st>>>(Minv.ptr, m);
    CA_CUDA(cudaGetLastError());
    count_launch(2);
    lu::factor(Mt.ptr, m, piv.ptr, flag.ptr, true, st);
    lu::solve_many(Mt.ptr, m, piv.ptr, Minv.ptr, m, st);
    int32_t sing = 0;
    CA_CUDA(cudaMemcpyAsync(&sing, flag.ptr, sizeof sing, cudaMemcpyDeviceToHost, st));
    CA_CUDA(cudaStreamSynchronize(st));
    precond = sing == 0;
    return precond;
  }
  // v <- M^-1 v (through tmp2)
  void apply_preconditioner(double* v) {
    if (!precond) return;
    bind();
    multi_dot_kernel<<<(unsigned)m, 256, 0, st>>>(Minv.ptr, v, m, m, pre.ptr);
    CA_CUDA(cudaGetLastError());
    count_launch();
    CA_CUDA(cudaMemcpyAsync(v, pre.ptr, (size_t)m * sizeof(double), cudaMemcpyDeviceToDevice, st));
  }
  void axpby(double a, const double* xv, double b, const double* yv, double* out) {
    bind();
    axpby_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(a, xv, b, yv ? yv : xv, m, out);
    CA_CUDA(cudaGetLastError());
    count_launch();
  }
  // out[0..k) = V[:, 0..k)' w, asynchronously
  void project(int k, const double* wv, double* out) {
    bind();
    multi_dot_kernel<<<(unsigned)k, 256, 0, st>>>(V.ptr, wv, m, k, out);
    CA_CUDA(cudaGetLastError());
    count_launch();
  }
  // w -= V[:, 0..k) hvec
  void subtract(int k, const double* hvec, double* wv) {
    bind();
    combine_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(V.ptr, hvec, m, k, -1.0, wv);
    CA_CUDA(cudaGetLastError());
    count_launch();
  }
  void dot(const double* a, const double* b, double* out) {
    bind();
    multi_dot_kernel<<<1, 256, 0, st>>>(a, b, m, 1, out);  // one basis vector: out[0] = a . b
    CA_CUDA(cudaGetLastError());
    count_launch();
  }
  void sync() {
    bind();
    CA_CUDA(cudaStreamSynchronize(st));
  }
};

struct ReplicaRange {
  std::vector<Replica*>* v;
  struct It {
    Replica** p;
    Replica& operator*() const { return **p; }
    It& operator++() { ++p; return *this; }
    bool operator!=(const It& o) const { return p != o.p; }
  };
  It begin() const { return It{v->data()}; }
  It end() const { return It{v->data() + v->size()}; }
};

struct Solver {
  std::vector<std::unique_ptr<Replica>> owned;
  std::vector<Replica*> rp;
  int m = 0;
  int jvps = 0, fevals = 0;
  ReplicaRange reps{&rp};

  // y = f(x) or Df(x) v on every replica (row slabs + all-gather), result in rep.padded[0..m)
  void eval_all(int mode, double R, DeviceBuffer<double> Replica::*xin, const double* (*vsel)(Replica&), DeviceBuffer<double> Replica::*out) {
    for (Replica& r : reps) r.eval_slab(mode, (r.*xin).ptr, vsel ? vsel(r) : nullptr, R);
    if (rp[0]->world > 1) {
      CA_NCCL(nccl().GroupStart());
      for (Replica& r : reps) r.gather_issue();
      CA_NCCL(nccl().GroupEnd());
    }
    for (Replica& r : reps) r.copy_out((r.*out).ptr);
  }
  double norm_of(DeviceBuffer<double> Replica::*v) {
    for (Replica& r : reps) r.dot((r.*v).ptr, (r.*v).ptr, r.hcol.ptr);
    double s = 0;
    Replica& r0 = *rp[0];
    r0.bind();
    CA_CUDA(cudaMemcpyAsync(&s, r0.hcol.ptr, sizeof(double), cudaMemcpyDeviceToHost, r0.st));
    for (Replica& r : reps) r.sync();
    return std::sqrt(s);
  }
};

const double* v_is_w(Replica& r) { return r.w.ptr; }
const double* v_is_dx(Replica& r) { return r.dx.ptr; }

}  // namespace

extern "C" int32_t ca_newton_krylov(ca_model** models, int32_t nmodels, ca_comm* comm, double R, const double* xguess,
                                    const ca_krylov_params* params, double* x_out, int32_t* success,
                                    ca_krylov_info* info) {
  return guarded([&] {
    if (!models || nmodels < 1 || !xguess || !x_out || !success) fail(CA_ERR_INVALID, "ca_newton_krylov: null argument");
    if (!(R != 0.0)) fail(CA_ERR_INVALID, "R must be non-zero");
    ca_krylov_params p{1e-10, 1e-8, 30, 60, 5, 0, 1};
    if (params) p = *params;
    if (p.max_newton < 0 || p.restart < 1 || p.max_restarts < 1 || !(p.gmres_rtol > 0)) fail(CA_ERR_INVALID, "bad Krylov parameters");
    ca_comm* c = comm;
    if (!c && nmodels > 1 && default_comm() && (int)default_comm()->devices.size() == nmodels) c = default_comm();
    const int world = c ? c->nranks : 1;
    if (c && (int)c->devices.size() != nmodels)
      fail(CA_ERR_INVALID, "ca_newton_krylov: %d models for a communicator with %zu local devices", nmodels, c->devices.size());
    if (!c && nmodels != 1) fail(CA_ERR_INVALID, "ca_newton_krylov: several models need a communicator");
    int prev = 0;
    CA_CUDA(cudaGetDevice(&prev));
    Solver S;
    S.m = (int)models[0]->m;
    const int m = S.m;
    for (int d = 0; d < nmodels; ++d) {
      S.owned.emplace_back(new Replica);
      S.rp.push_back(S.owned.back().get());
    }
    std::vector<cudaStream_t> own_streams;
    for (int d = 0; d < nmodels; ++d) {
      if (!models[d] || models[d]->m != m) fail(CA_ERR_INVALID, "ca_newton_krylov: models must be replicas of one model");
      if (c && models[d]->device != c->devices[(size_t)d])
        fail(CA_ERR_INVALID, "ca_newton_krylov: model %d lives on device %d, the communicator's device %d is %d", d, models[d]->device, d, c->devices[(size_t)d]);
      cudaStream_t st = nullptr;
      if (c) st = c->streams[(size_t)d];
      else {
        CA_CUDA(cudaSetDevice(models[d]->device));
        CA_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        own_streams.push_back(st);
      }
      S.rp[(size_t)d]->init(models[d], c ? c->ranks[(size_t)d] : 0, world, (c && world > 1) ? c->comms[(size_t)d] : nullptr, st, p.restart);
    }
    cudaEvent_t e0, e1;
    Replica& lead = *S.rp[0];
    lead.bind();
    CA_CUDA(cudaEventCreate(&e0));
    CA_CUDA(cudaEventCreate(&e1));
    CA_CUDA(cudaEventRecord(e0, lead.st));
    for (Replica& r : S.reps) {
      r.bind();
      CA_CUDA(cudaMemcpyAsync(r.x.ptr, xguess, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, r.st));
    }
    if (p.precondition)
      for (Replica& r : S.reps)
        if (!r.build_preconditioner(R)) fail(CA_ERR_INVALID, "ca_newton_krylov: L1 + L2/R is singular, no preconditioner (set precondition = 0)");
    ca_krylov_info I{};
    const int k = p.restart;
    std::vector<double> Hm((size_t)(k + 1) * k, 0.0), cs((size_t)k), sn((size_t)k), g((size_t)k + 1), hh((size_t)2 * (k + 2)), y((size_t)k);
    auto Hat = [&](int i, int j) -> double& { return Hm[(size_t)j * (k + 1) + i]; };

    S.eval_all(MODE_QUAD, R, &Replica::x, nullptr, &Replica::fx);
    ++S.fevals;
    bool converged = false;
    double nf = 0.0;
    for (int it = 0; it <= p.max_newton; ++it) {
      nf = S.norm_of(&Replica::fx);
      if (it < CA_LOG_STEPS) I.residual[it] = nf;
      if (nf <= p.ftol) { converged = true; break; }
      if (it == p.max_newton) break;
      // ---- restarted GMRES for Df dx = -fx
      for (Replica& r : S.reps) {
        r.bind();
        CA_CUDA(cudaMemsetAsync(r.dx.ptr, 0, (size_t)m * sizeof(double), r.st));
      }
      // preconditioned right-hand side  -M^-1 fx  (kept in ft until the line search overwrites it)
      for (Replica& r : S.reps) {
        r.axpby(-1.0, r.fx.ptr, 0.0, nullptr, r.ft.ptr);
        r.apply_preconditioner(r.ft.ptr);
      }
      const double nb = S.norm_of(&Replica::ft);  // GMRES tolerances are relative to the preconditioned residual
      double rnorm = nb;
      for (int rs = 0; rs < p.max_restarts; ++rs) {
        // r0 = M^-1 (-fx - Df dx)  -> V[:, 0]
        if (rs > 0) {
          S.eval_all(MODE_JVP, R, &Replica::x, v_is_dx, &Replica::tmp);
          ++S.jvps;
          for (Replica& r : S.reps) {
            r.apply_preconditioner(r.tmp.ptr);
            r.axpby(1.0, r.ft.ptr, -1.0, r.tmp.ptr, r.w.ptr);
          }
        } else {
          for (Replica& r : S.reps) r.axpby(1.0, r.ft.ptr, 0.0, nullptr, r.w.ptr);
        }
        const double beta = S.norm_of(&Replica::w);
        if (beta <= p.gmres_rtol * nb) break;
        for (Replica& r : S.reps) r.axpby(1.0 / beta, r.w.ptr, 0.0, nullptr, r.V.ptr);
        std::fill(g.begin(), g.end(), 0.0);
        g[0] = beta;
        int used = 0;
        for (int j = 0; j < k; ++j) {
          // w = Df V_j, orthogonalised against V_0..V_j twice
          for (Replica& r : S.reps) {
            r.bind();
            CA_CUDA(cudaMemcpyAsync(r.tmp.ptr, r.V.ptr + (size_t)j * m, (size_t)m * sizeof(double), cudaMemcpyDeviceToDevice, r.st));
          }
          S.eval_all(MODE_JVP, R, &Replica::x, [](Replica& r) -> const double* { return r.tmp.ptr; }, &Replica::w);
          ++S.jvps;
          for (Replica& r : S.reps) {
            r.apply_preconditioner(r.w.ptr);
            r.project(j + 1, r.w.ptr, r.hcol.ptr);
            r.subtract(j + 1, r.hcol.ptr, r.w.ptr);
            r.project(j + 1, r.w.ptr, r.hcol.ptr + (k + 2));
            r.subtract(j + 1, r.hcol.ptr + (k + 2), r.w.ptr);
            r.dot(r.w.ptr, r.w.ptr, r.hcol.ptr + (j + 1));
          }
          lead.bind();
          CA_CUDA(cudaMemcpyAsync(hh.data(), lead.hcol.ptr, (size_t)2 * (k + 2) * sizeof(double), cudaMemcpyDeviceToHost, lead.st));
          for (Replica& r : S.reps) r.sync();
          for (int i = 0; i <= j; ++i) Hat(i, j) = hh[(size_t)i] + hh[(size_t)(k + 2) + i];
          const double hn = std::sqrt(std::max(0.0, hh[(size_t)j + 1]));
          Hat(j + 1, j) = hn;
          // Givens rotations: previous ones on the new column, then a new one
          for (int i = 0; i < j; ++i) {
            const double t = cs[(size_t)i] * Hat(i, j) + sn[(size_t)i] * Hat(i + 1, j);
            Hat(i + 1, j) = -sn[(size_t)i] * Hat(i, j) + cs[(size_t)i] * Hat(i + 1, j);
            Hat(i, j) = t;
          }
          const double den = std::hypot(Hat(j, j), Hat(j + 1, j));
          cs[(size_t)j] = den > 0 ? Hat(j, j) / den : 1.0;
          sn[(size_t)j] = den > 0 ? Hat(j + 1, j) / den : 0.0;
          Hat(j, j) = den;
          Hat(j + 1, j) = 0.0;
          g[(size_t)j + 1] = -sn[(size_t)j] * g[(size_t)j];
          g[(size_t)j] = cs[(size_t)j] * g[(size_t)j];
          rnorm = std::fabs(g[(size_t)j + 1]);
          used = j + 1;
          if (hn == 0.0 || rnorm <= p.gmres_rtol * nb) break;
          for (Replica& r : S.reps) r.axpby(1.0 / hn, r.w.ptr, 0.0, nullptr, r.V.ptr + (size_t)(j + 1) * m);
        }
        // back substitution, dx += V y
        for (int i = used - 1; i >= 0; --i) {
          double s = g[(size_t)i];
          for (int q = i + 1; q < used; ++q) s -= Hat(i, q) * y[(size_t)q];
          y[(size_t)i] = s / Hat(i, i);
        }
        for (Replica& r : S.reps) {
          r.bind();
          CA_CUDA(cudaMemcpyAsync(r.hcol.ptr, y.data(), (size_t)used * sizeof(double), cudaMemcpyHostToDevice, r.st));
          combine_kernel<<<(unsigned)((m + 255) / 256), 256, 0, r.st>>>(r.V.ptr, r.hcol.ptr, m, used, 1.0, r.dx.ptr);
          CA_CUDA(cudaGetLastError());
          count_launch();
        }
        for (Replica& r : S.reps) r.sync();  // y is reused on the host
        if (rnorm <= p.gmres_rtol * nb) break;
      }
      // ---- backtracking line search on |f|
      double lam = 1.0;
      for (;;) {
        for (Replica& r : S.reps) r.axpby(1.0, r.x.ptr, lam, r.dx.ptr, r.xt.ptr);
        S.eval_all(MODE_QUAD, R, &Replica::xt, nullptr, &Replica::ft);
        ++S.fevals;
        const double nt = S.norm_of(&Replica::ft);
        if (nt < (1 - 1e-4 * lam) * nf || lam < 1e-4) break;
        lam *= 0.5;
      }
      for (Replica& r : S.reps) {
        r.bind();
        CA_CUDA(cudaMemcpyAsync(r.x.ptr, r.xt.ptr, (size_t)m * sizeof(double), cudaMemcpyDeviceToDevice, r.st));
        CA_CUDA(cudaMemcpyAsync(r.fx.ptr, r.ft.ptr, (size_t)m * sizeof(double), cudaMemcpyDeviceToDevice, r.st));
      }
      I.newton_steps = it + 1;
    }
    lead.bind();
    CA_CUDA(cudaEventRecord(e1, lead.st));
    CA_CUDA(cudaMemcpyAsync(x_out, lead.x.ptr, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, lead.st));
    for (Replica& r : S.reps) r.sync();
    float ms = 0;
    CA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    for (auto& r : S.owned) {  // buffers are released while their device is current
      cudaSetDevice(r->device);
      r.reset();
    }
    for (size_t d = 0; d < own_streams.size(); ++d) {
      cudaSetDevice(models[d]->device);
      cudaStreamDestroy(own_streams[d]);
    }
    *success = converged ? 1 : 0;
    I.converged = converged ? 1 : 0;
    I.jvps = S.jvps;
    I.f_evals = S.fevals;
    I.final_residual = nf;
    I.device_seconds = ms * 1e-3;
    I.world = world;
    if (info) *info = I;
    CA_CUDA(cudaSetDevice(prev));
  });
}
