// Newton-hookstep for models too large for the one-CTA solver (m > 95): hookstepsolve / hookstep of
// src/Hookstep.jl:21-74,109-352 with f, Df, Df'Df, the LU factors and every vector RESIDENT ON THE DEVICE.
//
// The host runs the reference's decision tree (same constants, same exits) on a handful of scalars per step --
// norms and dot products that come back as one small Gram matrix -- and launches:
//   f(x)            the model's few-state streaming kernel (stream.cu)
//   Df(x)           m <= 400: the per-entry Jacobian kernel (bilinear.cu); larger: the Jacobian ACTION on the m unit
//                   vectors through the DMMA ensemble kernel (one launch, no Jacobian structure to build)
//   dx = -Df \ fx   blocked right-looking LU with partial pivoting in global memory (panel of 16 columns factored by
//                   one CTA, row swaps, triangular solve of the block row, trailing update by the tiled DGEMM of
//                   dgemm.cuh), then blocked forward / backward substitution
//   hookstep        H = Df'Df once per Newton step (DGEMM), (H + mu I) factored with the same LU, one factorisation
//                   per mu iteration (the reference solves twice per iteration with the same matrix across iterations,
//                   Hookstep.jl:45,60)
// Nothing of size m or m^2 crosses PCIe inside the search; the result and the log are copied out at the end.
#include <algorithm>
#include <cmath>
#include <vector>

#include "bilinear.h"
#include "bilinear_kernels.cuh"
#include "dense_lu.cuh"
#include "dgemm.cuh"
#include "model.h"

using namespace ca;

namespace {

__global__ void add_diag_kernel(const double* __restrict__ H, double mu, int m, double* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)m * m) return;
  out[t] = H[t] + ((t % m) == (t / m) ? mu : 0.0);
}

// out = a + s * b
__global__ void axpy_kernel(const double* __restrict__ a, double s, const double* __restrict__ b, int n, double* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = fma(s, b[t], a[t]);
}

// X[s, j] = x[j] for s < n and V = identity: the ensemble whose Jacobian actions are the columns of Df
__global__ void unit_ensemble_kernel(const double* __restrict__ x, int m, double* __restrict__ X, double* __restrict__ V) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)m * m) return;
  const int s = (int)(t % m), j = (int)(t / m);
  X[t] = x[j];
  V[t] = s == j ? 1.0 : 0.0;
}

// out[j + m * i] = in[i + m * j]
__global__ void transpose_sq_kernel(const double* __restrict__ in, double* __restrict__ out, int m) {
  __shared__ double tile[32][33];
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8)
    if (y0 + r < m && x0 + threadIdx.x < m) tile[r][threadIdx.x] = in[(size_t)(y0 + r) * m + x0 + threadIdx.x];
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8)
    if (x0 + r < m && y0 + threadIdx.x < m) out[(size_t)(x0 + r) * m + y0 + threadIdx.x] = tile[threadIdx.x][r];
}

struct LargeSolver {
  ca_model* h;
  int m;
  cudaStream_t st = nullptr;
  DeviceBuffer<double> x, fx, dx, dxn, xh, fh, Dfdx, g, tmp, Df, LU, H, HLU, gram, ensX, ensV, ensO;
  DeviceBuffer<int32_t> piv, hpiv, flag;
  bool H_ready = false;
  int nf = 0, ndf = 0;

  explicit LargeSolver(ca_model* h_) : h(h_), m((int)h_->m) {
    const size_t n = (size_t)m, nn = n * n;
    for (DeviceBuffer<double>* v : {&x, &fx, &dx, &dxn, &xh, &fh, &Dfdx, &g, &tmp}) v->alloc(n);
    Df.alloc(nn);
    LU.alloc(nn);
    H.alloc(nn);
    HLU.alloc(nn);
    gram.alloc(16);
    piv.alloc(n);
    hpiv.alloc(n);
    flag.alloc(1);
  }

  void f(const double* xin, double R, double* out) {
    h->eval_dev(MODE_QUAD, xin, nullptr, 1, 1, R, out, 1, st);
    ++nf;
  }
  void jac(const double* xin, double R) {
    ++ndf;
    if (m <= 400) {
      const int32_t s = ca_model_jac_dev(h, xin, R, Df.ptr, st);
      if (s != CA_OK) fail(s, "%s", ca_last_error());
      return;
    }
    const size_t nn = (size_t)m * m;
    if (ensX.count < nn) {
      ensX.alloc(nn);
      ensV.alloc(nn);
      ensO.alloc(nn);
    }
    unit_ensemble_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, st>>>(xin, m, ensX.ptr, ensV.ptr);
    CA_CUDA(cudaGetLastError());
    count_launch();
    // OUT[s, i] = (Df e_s)_i = Df[i, s]: the ensemble output is Df row-major; transpose into column-major
    h->eval_dev(MODE_JVP, ensX.ptr, ensV.ptr, m, m, R, ensO.ptr, m, st);
    transpose_sq_kernel<<<dim3((unsigned)((m + 31) / 32), (unsigned)((m + 31) / 32)), dim3(32, 8), 0, st>>>(ensO.ptr, Df.ptr, m);
    CA_CUDA(cudaGetLastError());
    count_launch();
  }
  // A (column-major m x m, overwritten by its factors)
  void factor(double* A, int32_t* p, bool pivoting) {
    CA_CUDA(cudaMemsetAsync(flag.ptr, 0, sizeof(int32_t), st));
    lu::factor(A, m, p, flag.ptr, pivoting, st);
  }
  bool singular() {
    int32_t s = 0;
    CA_CUDA(cudaMemcpyAsync(&s, flag.ptr, sizeof s, cudaMemcpyDeviceToHost, st));
    CA_CUDA(cudaStreamSynchronize(st));
    return s != 0;
  }
  // b <- scale * A^-1 b with the factors of factor()
  void solve(const double* A, const int32_t* p, double* b, double scale) { lu::solve(A, m, p, b, scale, st); }
  void gemv(const double* A, bool transpose, const double* v, double* out) {
    const MatView a = transpose ? transposed(colmajor(A, m)) : colmajor(A, m);
    launch_dgemm(m, 1, m, 1.0, a, colmajor(v, m), 0.0, out, 1, m, st);
  }
  // Gram matrix of up to 4 vectors -> host (row-major k x k)
  void dots(std::initializer_list<const double*> vs, double* out) {
    const int k = (int)vs.size();
    int a = 0;
    for (const double* u : vs) {
      int b = 0;
      for (const double* w : vs) {
        if (b >= a) launch_dgemm(1, 1, m, 1.0, rowmajor(u, m), colmajor(w, m), 0.0, gram.ptr + a * k + b, 1, 1, st);
        ++b;
      }
      ++a;
    }
    CA_CUDA(cudaMemcpyAsync(out, gram.ptr, (size_t)k * k * sizeof(double), cudaMemcpyDeviceToHost, st));
    CA_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < k; ++i)
      for (int j = 0; j < i; ++j) out[i * k + j] = out[j * k + i];
  }
  void copy(double* dst, const double* src, size_t n) {
    CA_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  }
  void axpy(const double* a, double s, const double* b, double* out) {
    axpy_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(a, s, b, m, out);
    CA_CUDA(cudaGetLastError());
    count_launch();
  }

  // hookstep(fx, Dfx, delta, dx_newton) -- Hookstep.jl:21-74; result in dx, returns its norm
  double hookstep(double delta, double norm_dx_newt, int Nmusearch, bool& sing) {
    const double dtol = 1e-4;
    double norm_dx = norm_dx_newt, mu = 1e-6;
    copy(dx.ptr, dxn.ptr, (size_t)m);
    if (!H_ready) {
      launch_dgemm(m, m, m, 1.0, transposed(colmajor(Df.ptr, m)), colmajor(Df.ptr, m), 0.0, H.ptr, 1, m, st);  // H = Df'Df :32
      gemv(Df.ptr, true, fx.ptr, g.ptr);                                                                         // Df' fx
      H_ready = true;
    }
    const int64_t nn = (int64_t)m * m;
    auto refactor = [&](double mu_) {
      add_diag_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, st>>>(H.ptr, mu_, m, HLU.ptr);
      count_launch();
      factor(HLU.ptr, hpiv.ptr, true);
    };
    refactor(mu);
    for (int it = 0; it < Nmusearch; ++it) {
      const double phi = norm_dx - delta;
      copy(tmp.ptr, dx.ptr, (size_t)m);
      solve(HLU.ptr, hpiv.ptr, tmp.ptr, 1.0);  // (H + mu I) \ dx                      :45,:53
      double d2[4];
      dots({dx.ptr, tmp.ptr}, d2);
      const double phiprime = -(1.0 / norm_dx) * d2[1];
      mu = mu - (norm_dx / delta) * (phi / phiprime);
      refactor(mu);
      if (singular()) { sing = true; return norm_dx; }
      copy(dx.ptr, g.ptr, (size_t)m);
      solve(HLU.ptr, hpiv.ptr, dx.ptr, -1.0);  // dx = -(H + mu I) \ (Df' fx)           :60
      double n2[1];
      dots({dx.ptr}, n2);
      norm_dx = std::sqrt(n2[0]);
      if (std::fabs(norm_dx - delta) / delta <= dtol) break;
    }
    return norm_dx;
  }

  // hookstepsolve -- Hookstep.jl:109-352
  void run(double R, const double* xguess_host, const ca_search_params& p, double* x_out_host, int32_t* success,
           ca_solve_log* L) {
    cudaEvent_t e0, e1;
    CA_CUDA(cudaEventCreate(&e0));
    CA_CUDA(cudaEventCreate(&e1));
    CA_CUDA(cudaEventRecord(e0, st));
    CA_CUDA(cudaMemcpyAsync(x.ptr, xguess_host, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, st));
    double delta = p.delta, last_res = 0.0;
    int exit_reason = CA_EXIT_MAX_NEWTON, nsteps = 0;
    nf = ndf = 0;
    const double* result = x.ptr;
    for (int n = 0; n < p.Nnewton; ++n) {
      f(x.ptr, R, fx.ptr);
      double s1[1];
      dots({fx.ptr}, s1);
      const double rx = 0.5 * s1[0];
      last_res = std::sqrt(2 * rx);
      if (L && n < CA_LOG_STEPS) { L->residual[n] = last_res; L->delta[n] = delta; L->hooksteps[n] = 0; }
      if (last_res <= p.ftol) { exit_reason = CA_EXIT_CONVERGED; break; }
      nsteps = n + 1;
      jac(x.ptr, R);
      H_ready = false;
      copy(LU.ptr, Df.ptr, (size_t)m * m);
      factor(LU.ptr, piv.ptr, true);
      if (singular()) { exit_reason = CA_EXIT_SINGULAR; break; }
      copy(dx.ptr, fx.ptr, (size_t)m);
      solve(LU.ptr, piv.ptr, dx.ptr, -1.0);  // dx = -Dfx \ fx   :158
      gemv(Df.ptr, false, dx.ptr, Dfdx.ptr);
      double s3[9];
      dots({fx.ptr, Dfdx.ptr, dx.ptr}, s3);
      double norm_dx = std::sqrt(s3[8]);
      if (s3[1] >= 0) { exit_reason = CA_EXIT_RESIDUAL_UP; break; }
      if (norm_dx < p.xtol) {
        axpy(x.ptr, 1.0, dx.ptr, xh.ptr);
        result = xh.ptr;
        exit_reason = CA_EXIT_XTOL_NEWTON;
        break;
      }
      copy(dxn.ptr, dx.ptr, (size_t)m);
      const double norm_dx_newt = norm_dx;
      axpy(x.ptr, 1.0, dx.ptr, xh.ptr);
      int nh = 0;
      bool xtol_exit = false, sing = false;
      for (int hk = 0; hk < p.Nhook; ++hk) {
        ++nh;
        if (norm_dx_newt > delta) {
          norm_dx = hookstep(delta, norm_dx_newt, p.Nmusearch, sing);
          if (sing) break;
        }
        gemv(Df.ptr, false, dx.ptr, Dfdx.ptr);
        if (norm_dx < p.xtol) {
          axpy(x.ptr, 1.0, dx.ptr, xh.ptr);
          xtol_exit = true;
          break;
        }
        const bool within = norm_dx_newt <= delta;
        axpy(x.ptr, 1.0, dx.ptr, xh.ptr);
        f(xh.ptr, R, fh.ptr);
        axpy(fx.ptr, 1.0, Dfdx.ptr, tmp.ptr);  // fx + Df dx
        double s4[9];
        dots({fh.ptr, tmp.ptr, fx.ptr}, s4);
        double s5[4];
        dots({fx.ptr, Dfdx.ptr}, s5);
        const double r_hook = 0.5 * s4[0], r_linear = rx + s5[1], r_quadratic = 0.5 * s4[4];
        const double dr_hook = rx - r_hook, dr_linear = rx - r_linear, dr_quadratic = rx - r_quadratic;
        if (dr_hook > dr_linear) {
          if (within) break;
          delta = 3 * delta / 2;
          continue;
        } else if (dr_hook < 0.01 * dr_quadratic) {
          delta = delta / 2;
          continue;
        } else if (dr_hook < 0.10 * dr_quadratic) {
          delta = delta / 2;
          break;
        } else if (0.8 * dr_quadratic < dr_hook && dr_hook < 1.2 * dr_quadratic) {
          const double mm = -dr_linear / delta;
          const double c = (r_hook - rx - mm * delta) / (delta * delta);
          const double dnew = -mm / (2 * c);
          if (within) {
          } else if (dnew < 2 * delta / 3) {
            delta = 2 * delta / 3;
          } else if (dnew > 4 * delta / 3) {
            delta = 4 * delta / 3;
          } else {
            delta = dnew;
          }
          break;
        } else {
          break;
        }
      }
      if (L && n < CA_LOG_STEPS) { L->delta[n] = delta; L->hooksteps[n] = nh; }
      if (sing) { exit_reason = CA_EXIT_SINGULAR; break; }
      if (xtol_exit) {
        result = xh.ptr;
        exit_reason = CA_EXIT_XTOL_HOOK;
        break;
      }
      copy(x.ptr, xh.ptr, (size_t)m);
    }
    CA_CUDA(cudaEventRecord(e1, st));
    CA_CUDA(cudaMemcpyAsync(x_out_host, result, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, st));
    CA_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    CA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *success = exit_reason == CA_EXIT_CONVERGED ? 1 : 0;
    if (L) {
      L->exit_reason = exit_reason;
      L->newton_steps = nsteps;
      L->f_evals = nf;
      L->df_evals = ndf;
      L->final_residual = last_res;
      L->final_delta = delta;
      L->device_seconds = ms * 1e-3;
    }
  }
};

}  // namespace

// called by ca_hookstep_solve_model for models beyond the one-CTA solver's shared-memory limit
void hookstep_solve_large(ca_model* h, const double* R, const double* Xguess, int64_t nsolve, const ca_search_params& p,
                          double* X_out, int32_t* success, ca_solve_log* logs) {
  const int m = (int)h->m;
  LargeSolver S(h);
  std::vector<double> xg((size_t)m), xo((size_t)m);
  for (int64_t s = 0; s < nsolve; ++s) {
    for (int i = 0; i < m; ++i) xg[(size_t)i] = Xguess[s + nsolve * i];
    if (logs) logs[s] = ca_solve_log{};
    S.run(R[s], xg.data(), p, xo.data(), &success[s], logs ? &logs[s] : nullptr);
    for (int i = 0; i < m; ++i) X_out[s + nsolve * i] = xo[(size_t)i];
  }
}

// LU solve of a dense column-major system on the device (diagnostics / tests of the blocked LU): A is overwritten by
// its factors, b by the solution; returns 1 in *singular if a pivot was exactly zero
extern "C" int32_t ca_selftest_lu_solve(double* A_host, int64_t m64, double* b_host, int32_t* singular) {
  return guarded([&] {
    if (!A_host || !b_host || m64 <= 0 || m64 > 32768) fail(CA_ERR_INVALID, "ca_selftest_lu_solve: bad arguments");
    require_device();
    const int m = (int)m64;
    DeviceBuffer<double> A, b;
    DeviceBuffer<int32_t> piv, flag;
    A.upload(A_host, (size_t)m * m);
    b.upload(b_host, (size_t)m);
    piv.alloc((size_t)m);
    flag.alloc(1);
    CA_CUDA(cudaMemset(flag.ptr, 0, sizeof(int32_t)));
    lu::factor(A.ptr, m, piv.ptr, flag.ptr, true, nullptr);
    lu::solve(A.ptr, m, piv.ptr, b.ptr, 1.0, nullptr);
    int32_t s = 0;
    CA_CUDA(cudaMemcpy(&s, flag.ptr, sizeof s, cudaMemcpyDeviceToHost));
    if (singular) *singular = s;
    CA_CUDA(cudaMemcpy(A_host, A.ptr, (size_t)m * m * sizeof(double), cudaMemcpyDeviceToHost));
    CA_CUDA(cudaMemcpy(b_host, b.ptr, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost));
  });
}
