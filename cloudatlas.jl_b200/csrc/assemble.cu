// Galerkin assembly of B, A1, A2 and the quadratic tensor N on the device (separable fast path).
//
// Replaces the m^2 and m^3 loops of the reference constructor (src/ODEModels.jl:31-49), which call the
// symbolic inner-product algebra (src/BasisFunctions.jl:87-99,278-298,328-356,444-463) once per entry.
// Every basis component is coef * E_jx(x) E_kz(z) S_l^(d)(y), so
//   <Psi_i, (Psi_j . grad) Psi_k> = sum_{c,a} coef_ic coef_ja coef_kc * X3 * Z3 * Y3
// with X3/Z3 closed-form Fourier triple averages (tables of (2J+1)^3 x 2 entries) and
//   Y3[p,q,r] = (1/2) int S_p S_q S_r dy  = sum_n w_n F_p(y_n) F_q(y_n) F_r(y_n)
// an exact Gauss-Legendre contraction done once per basis on the device.  N is produced directly in
// the reference's COO order by a count pass, a device scan and a fill pass: no m^3 scratch array.
#include <cub/device/device_scan.cuh>

#include <memory>
#include <vector>

#include "basis.h"
#include "common.h"

using namespace ca;

namespace {

constexpr double kAbsZero = 1e-15;  // ODEModels.jl:46
constexpr double kRelZero = 1e-11;  // cancellation-aware: |sum| vs sum|terms|
constexpr double kQuadZero = 1e-15; // quadrature result vs sum|w f g h|: below this it is an analytic zero

struct DevBasis {
  int m, J, K, npid, nx, nz;
  double alpha, gamma;
  const double* coef;   // [m][3]
  const int32_t* jx;    // [m][3], shifted by +J
  const int32_t* kz;    // [m][3], shifted by +K
  const int32_t* pid;   // [m][3]
  const double* k2;     // [m][3]  (alpha jx)^2 + (gamma kz)^2
  const double *X2, *Z2, *X3, *Z3, *Y2, *Y2y, *Y3;
};

// double-double helpers (error-free transformations): the wall-normal integrals must be accurate
// relative to THEMSELVES, not to the size of the quadrature's terms, because N entries that vanish
// by cancellation between components are recognised by |sum| <= 1e-11 * sum|terms|.
struct dd { double hi, lo; };
__device__ __forceinline__ dd dd_mul(dd a, dd b) {
  const double p = a.hi * b.hi;
  double e = fma(a.hi, b.hi, -p);
  e = fma(a.hi, b.lo, e);
  e = fma(a.lo, b.hi, e);
  const double s = p + e;
  return dd{s, e - (s - p)};
}
__device__ __forceinline__ dd dd_add(dd a, dd b) {
  const double s = a.hi + b.hi, bb = s - a.hi;
  double e = (a.hi - (s - bb)) + (b.hi - bb);
  e += a.lo + b.lo;
  const double t = s + e;
  return dd{t, e - (t - s)};
}

__global__ void y3_kernel(const double* __restrict__ F, const double* __restrict__ Flo,
                          const double* __restrict__ w, const double* __restrict__ wlo,
                          const int8_t* __restrict__ parity, int npid, int ny, double* __restrict__ Y3) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)npid * npid * npid;
  if (t >= total) return;
  const int r = (int)(t % npid), q = (int)((t / npid) % npid), p = (int)(t / ((int64_t)npid * npid));
  dd s{0.0, 0.0};
  double mag = 0.0;
  if (parity[p] * parity[q] * parity[r] == 1) {  // odd integrands vanish exactly
    for (int n = 0; n < ny; ++n) {
      const dd fp{F[p * ny + n], Flo[p * ny + n]}, fq{F[q * ny + n], Flo[q * ny + n]},
          fr{F[r * ny + n], Flo[r * ny + n]}, wn{w[n], wlo[n]};
      const dd term = dd_mul(dd_mul(wn, fp), dd_mul(fq, fr));
      s = dd_add(s, term);
      mag += fabs(term.hi);
    }
  }
  // Integrals that vanish analytically without being odd (orthogonality of the underlying Legendre
  // factors) come out as ~1e-19 * sum|terms| (the node values carry a 64-bit mantissa): flush them to the
  // exact zero the reference's polynomial algebra produces.
  const double v = s.hi + s.lo;
  Y3[t] = fabs(v) > kQuadZero * mag ? v : 0.0;
}

// <d_p f, d_q g> for component ca of mode i and component cb of mode j; p, q in {0: none, 1: d/dx, 2: d/dy,
// 3: d/dz}.  d/dx E_a = alpha a E_{-a}; the shifted index of -a is 2J - (a+J).
__device__ __forceinline__ double comp_ip(const DevBasis& b, int i, int ca, int p, int j, int cb, int q) {
  const double ci = b.coef[i * 3 + ca], cj = b.coef[j * 3 + cb];
  if (ci == 0.0 || cj == 0.0) return 0.0;
  int xi = b.jx[i * 3 + ca], xj = b.jx[j * 3 + cb], zi = b.kz[i * 3 + ca], zj = b.kz[j * 3 + cb];
  double f = ci * cj;
  if (p == 1) { f *= b.alpha * (xi - b.J); xi = 2 * b.J - xi; }
  if (q == 1) { f *= b.alpha * (xj - b.J); xj = 2 * b.J - xj; }
  if (p == 3) { f *= b.gamma * (zi - b.K); zi = 2 * b.K - zi; }
  if (q == 3) { f *= b.gamma * (zj - b.K); zj = 2 * b.K - zj; }
  const double xz = b.X2[(xi * b.nx + xj) * 2] * b.Z2[(zi * b.nz + zj) * 2];
  if (f == 0.0 || xz == 0.0) return 0.0;
  return f * xz * b.Y2[(b.pid[i * 3 + ca] + (p == 2)) * b.npid + b.pid[j * 3 + cb] + (q == 2)];
}

// one thread per (i,j): rows slab [r0,r1) x m, row-major
__global__ void linear_kernel(DevBasis b, int r0, int r1, double* __restrict__ B, double* __restrict__ A1,
                              double* __restrict__ A2, double* __restrict__ D) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)(r1 - r0) * b.m) return;
  const int i = r0 + (int)(t / b.m), j = (int)(t % b.m);
  double vb = 0, va1 = 0, va2 = 0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double ci = b.coef[i * 3 + c], cj = b.coef[j * 3 + c];
    if (ci == 0.0 || cj == 0.0) continue;
    const int xi = b.jx[i * 3 + c], xj = b.jx[j * 3 + c], zi = b.kz[i * 3 + c], zj = b.kz[j * 3 + c];
    const int pi = b.pid[i * 3 + c], pj = b.pid[j * 3 + c];
    const double z2 = b.Z2[(zi * b.nz + zj) * 2];
    if (z2 == 0.0) continue;
    const double x2 = b.X2[(xi * b.nx + xj) * 2], x2d = b.X2[(xi * b.nx + xj) * 2 + 1];
    if (x2 != 0.0) {
      const double f = ci * cj * x2 * z2, y2 = b.Y2[pi * b.npid + pj];
      vb += f * y2;
      va2 += f * (b.Y2[pi * b.npid + pj + 2] - b.k2[j * 3 + c] * y2);  // p'' - (a^2 j^2 + g^2 k^2) p
    }
    if (x2d != 0.0) va1 -= ci * cj * x2d * z2 * b.Y2y[pi * b.npid + pj];  // -<Psi_i, y d/dx Psi_j>
  }
  {  // -<Psi_i, [v_j, 0, 0]>
    const double ci = b.coef[i * 3 + 0], cj = b.coef[j * 3 + 1];
    if (ci != 0.0 && cj != 0.0) {
      const double f = b.X2[(b.jx[i * 3] * b.nx + b.jx[j * 3 + 1]) * 2] * b.Z2[(b.kz[i * 3] * b.nz + b.kz[j * 3 + 1]) * 2];
      if (f != 0.0) va1 -= ci * cj * f * b.Y2[b.pid[i * 3] * b.npid + b.pid[j * 3 + 1]];
    }
  }
  B[t] = vb;
  A1[t] = va1;
  A2[t] = va2;
  // D_ij = <curl Psi_i, curl Psi_j>, expanded as in dissipation_term (ODEModels.jl:87-127); u,v,w = 0,1,2
  const double wx = comp_ip(b, i, 2, 2, j, 2, 2) - comp_ip(b, i, 2, 2, j, 1, 3) - comp_ip(b, i, 1, 3, j, 2, 2) +
                    comp_ip(b, i, 1, 3, j, 1, 3);
  const double wy = comp_ip(b, i, 0, 3, j, 0, 3) - comp_ip(b, i, 0, 3, j, 2, 1) - comp_ip(b, i, 2, 1, j, 0, 3) +
                    comp_ip(b, i, 2, 1, j, 2, 1);
  const double wz = comp_ip(b, i, 1, 1, j, 1, 1) - comp_ip(b, i, 1, 1, j, 0, 2) - comp_ip(b, i, 0, 2, j, 1, 1) +
                    comp_ip(b, i, 0, 2, j, 0, 2);
  D[t] = wx + wy + wz;
}

struct IJ {  // warp-uniform factors of one (i,j) pair: the 9 (c,a) combinations
  double cij[9];
  const double* x3[9];  // X3 + ((xi*nx + xj)*nx)*2 + dx, indexed by xk*2
  const double* z3[9];
  const double* y3[9];  // Y3 + (pi*npid + pj)*npid + dy, indexed by pk
  unsigned live;        // bitmask of combinations with non-zero coefficient product
};

__device__ __forceinline__ void load_ij(const DevBasis& b, int i, int j, IJ& u) {
  u.live = 0;
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int q = c * 3 + a;
      const double ci = b.coef[i * 3 + c], cj = b.coef[j * 3 + a];
      u.cij[q] = ci * cj;
      if (u.cij[q] != 0.0) u.live |= 1u << q;
      u.x3[q] = b.X3 + ((size_t)(b.jx[i * 3 + c] * b.nx + b.jx[j * 3 + a]) * b.nx) * 2 + (a == 0 ? 1 : 0);
      u.z3[q] = b.Z3 + ((size_t)(b.kz[i * 3 + c] * b.nz + b.kz[j * 3 + a]) * b.nz) * 2 + (a == 2 ? 1 : 0);
      u.y3[q] = b.Y3 + ((size_t)(b.pid[i * 3 + c] * b.npid + b.pid[j * 3 + a]) * b.npid) + (a == 1 ? 1 : 0);
    }
}

__device__ __forceinline__ bool entry(const DevBasis& b, const IJ& u, int k, double& val) {
  double s = 0.0, mag = 0.0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double ck = b.coef[k * 3 + c];
    if (ck == 0.0) continue;
    const int xk = b.jx[k * 3 + c], zk = b.kz[k * 3 + c], pk = b.pid[k * 3 + c];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int q = c * 3 + a;
      if (!(u.live >> q & 1)) continue;
      const double x3 = u.x3[q][xk * 2];
      if (x3 == 0.0) continue;
      const double z3 = u.z3[q][zk * 2];
      if (z3 == 0.0) continue;
      const double t = u.cij[q] * ck * x3 * z3 * u.y3[q][pk];
      s -= t;
      mag += fabs(t);
    }
  }
  val = s;
  return fabs(s) > kAbsZero && fabs(s) > kRelZero * mag;
}

// FILL = false: counts[(i-r0)*m + j] = number of kept k;  FILL = true: writes the entries at offs[...]
template <bool FILL>
__global__ void __launch_bounds__(256)
n_kernel(DevBasis b, int r0, int r1, int64_t* __restrict__ counts, const int64_t* __restrict__ offs,
         int64_t* __restrict__ oi, int64_t* __restrict__ oj, int64_t* __restrict__ ok,
         double* __restrict__ oval) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t npairs = (int64_t)(r1 - r0) * b.m;
  for (int64_t t = warp; t < npairs; t += nwarps) {
    const int i = r0 + (int)(t / b.m), j = (int)(t % b.m);
    IJ u;
    load_ij(b, i, j, u);
    int64_t n = 0;
    const int64_t base = FILL ? offs[t] : 0;
    if (u.live) {
      for (int k0 = 0; k0 < b.m; k0 += 32) {
        const int k = k0 + lane;
        double v = 0.0;
        const bool keep = k < b.m && entry(b, u, k, v);
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        if (FILL && keep) {
          const int64_t at = base + n + __popc(mask & ((1u << lane) - 1));
          oi[at] = i + 1;
          oj[at] = j + 1;
          ok[at] = k + 1;
          oval[at] = v;
        }
        n += __popc(mask);
      }
    }
    if (!FILL && lane == 0) counts[t] = n;
  }
}

}  // namespace

struct ca_assembly {
  int device = 0;
  int64_t m = 0, r0 = 0, r1 = 0, nnz = 0;
  double seconds = 0;
  DeviceBuffer<double> B, A1, A2, D;  // (r1-r0) x m row-major; D = <curl, curl> (dissipation)
  DeviceBuffer<int64_t> oi, oj, ok;
  DeviceBuffer<double> oval;
  void bind() const { CA_CUDA(cudaSetDevice(device)); }
};

extern "C" {

int32_t ca_assemble(double alpha, double gamma, const int64_t* ijkl, int64_t m64, int32_t normalize,
                    int64_t row_begin, int64_t row_end, ca_assembly** out) {
  return guarded([&] {
    if (!out) fail(CA_ERR_INVALID, "ca_assemble: out is null");
    *out = nullptr;
    if (m64 <= 0 || !ijkl) fail(CA_ERR_INVALID, "ca_assemble: empty index set");
    if (row_begin < 0 || row_end > m64 || row_begin > row_end)
      fail(CA_ERR_INVALID, "ca_assemble: row range [%lld,%lld) outside 0:%lld", (long long)row_begin, (long long)row_end, (long long)m64);
    if (m64 > 46000) fail(CA_ERR_UNSUPPORTED, "ca_assemble: m too large");
    BasisTables T = build_tables(alpha, gamma, ijkl, m64, normalize != 0);
    require_device();
    std::unique_ptr<ca_assembly> h(new ca_assembly);
    CA_CUDA(cudaGetDevice(&h->device));
    const int m = (int)m64, r0 = (int)row_begin, r1 = (int)row_end, rows = r1 - r0;
    h->m = m;
    h->r0 = r0;
    h->r1 = r1;

    // ---- upload the tabular basis
    std::vector<double> coef((size_t)m * 3), k2((size_t)m * 3);
    std::vector<int32_t> jx((size_t)m * 3), kz((size_t)m * 3), pid((size_t)m * 3);
    for (size_t n = 0; n < (size_t)m * 3; ++n) {
      const Component& c = T.comp[n];
      coef[n] = c.coef;
      jx[n] = c.jx + T.J;
      kz[n] = c.kz + T.K;
      pid[n] = 4 * c.l + c.d;
      const double ax = T.alpha * c.jx, gz = T.gamma * c.kz;
      k2[n] = ax * ax + gz * gz;
    }
    DeviceBuffer<double> dcoef, dk2, dX2, dZ2, dX3, dZ3, dY2, dY2y, dY3, dF, dFlo, dw, dwlo;
    DeviceBuffer<int32_t> djx, dkz, dpid;
    DeviceBuffer<int8_t> dpar;
    dcoef.upload(coef.data(), coef.size());
    dk2.upload(k2.data(), k2.size());
    djx.upload(jx.data(), jx.size());
    dkz.upload(kz.data(), kz.size());
    dpid.upload(pid.data(), pid.size());
    dX2.upload(T.X2.data(), T.X2.size());
    dZ2.upload(T.Z2.data(), T.Z2.size());
    dX3.upload(T.X3.data(), T.X3.size());
    dZ3.upload(T.Z3.data(), T.Z3.size());
    dY2.upload(T.Y2.data(), T.Y2.size());
    dY2y.upload(T.Y2y.data(), T.Y2y.size());
    dF.upload(T.F.data(), T.F.size());
    dw.upload(T.weights.data(), T.weights.size());
    dFlo.upload(T.Flo.data(), T.Flo.size());
    dwlo.upload(T.wlo.data(), T.wlo.size());
    dpar.upload(T.parity.data(), T.parity.size());
    const int64_t ny3 = (int64_t)T.npid * T.npid * T.npid;
    dY3.alloc((size_t)ny3);

    cudaEvent_t e0, e1;
    CA_CUDA(cudaEventCreate(&e0));
    CA_CUDA(cudaEventCreate(&e1));
    CA_CUDA(cudaEventRecord(e0));
    y3_kernel<<<(unsigned)((ny3 + 255) / 256), 256>>>(dF.ptr, dFlo.ptr, dw.ptr, dwlo.ptr, dpar.ptr, T.npid, T.ny, dY3.ptr);
    CA_CUDA(cudaGetLastError());
    count_launch();

    DevBasis b{m, T.J, T.K, T.npid, 2 * T.J + 1, 2 * T.K + 1, T.alpha, T.gamma, dcoef.ptr, djx.ptr, dkz.ptr, dpid.ptr, dk2.ptr,
               dX2.ptr, dZ2.ptr, dX3.ptr, dZ3.ptr, dY2.ptr, dY2y.ptr, dY3.ptr};
    const int64_t nij = (int64_t)rows * m;
    if (nij > 0) {
      h->B.alloc((size_t)nij);
      h->A1.alloc((size_t)nij);
      h->A2.alloc((size_t)nij);
      h->D.alloc((size_t)nij);
      linear_kernel<<<(unsigned)((nij + 255) / 256), 256>>>(b, r0, r1, h->B.ptr, h->A1.ptr, h->A2.ptr, h->D.ptr);
      CA_CUDA(cudaGetLastError());
      count_launch();

      DeviceBuffer<int64_t> counts, offs;
      counts.alloc((size_t)nij + 1);
      offs.alloc((size_t)nij + 1);
      CA_CUDA(cudaMemsetAsync(counts.ptr, 0, ((size_t)nij + 1) * sizeof(int64_t)));
      int sms = 0;
      CA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
      const unsigned grid = (unsigned)std::min<int64_t>((nij + 7) / 8, (int64_t)sms * 16);
      n_kernel<false><<<grid, 256>>>(b, r0, r1, counts.ptr, nullptr, nullptr, nullptr, nullptr, nullptr);
      CA_CUDA(cudaGetLastError());
      count_launch();
      size_t tmp_bytes = 0;
      CA_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts.ptr, offs.ptr, (int)(nij + 1)));
      DeviceBuffer<char> tmp;
      tmp.alloc(tmp_bytes);
      CA_CUDA(cub::DeviceScan::ExclusiveSum(tmp.ptr, tmp_bytes, counts.ptr, offs.ptr, (int)(nij + 1)));
      count_launch();
      int64_t nnz = 0;
      CA_CUDA(cudaMemcpy(&nnz, offs.ptr + nij, sizeof(int64_t), cudaMemcpyDeviceToHost));
      h->nnz = nnz;
      if (nnz > 0) {
        h->oi.alloc((size_t)nnz);
        h->oj.alloc((size_t)nnz);
        h->ok.alloc((size_t)nnz);
        h->oval.alloc((size_t)nnz);
        n_kernel<true><<<grid, 256>>>(b, r0, r1, nullptr, offs.ptr, h->oi.ptr, h->oj.ptr, h->ok.ptr, h->oval.ptr);
        CA_CUDA(cudaGetLastError());
        count_launch();
      }
    }
    CA_CUDA(cudaEventRecord(e1));
    CA_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    CA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    h->seconds = ms * 1e-3;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *out = h.release();
  });
}

int32_t ca_assembly_destroy(ca_assembly* h) {
  return guarded([&] {
    if (!h) return;
    cudaSetDevice(h->device);
    delete h;
  });
}

int32_t ca_assembly_info(const ca_assembly* h, int64_t* m, int64_t* row_begin, int64_t* row_end,
                         int64_t* nnz, double* device_seconds) {
  return guarded([&] {
    if (!h) fail(CA_ERR_INVALID, "null handle");
    if (m) *m = h->m;
    if (row_begin) *row_begin = h->r0;
    if (row_end) *row_end = h->r1;
    if (nnz) *nnz = h->nnz;
    if (device_seconds) *device_seconds = h->seconds;
  });
}

int32_t ca_assembly_get_linear_dev(ca_assembly* h, double* dB, double* dA1, double* dA2, void* stream) {
  return guarded([&] {
    if (!h) fail(CA_ERR_INVALID, "null handle");
    h->bind();
    const size_t bytes = (size_t)(h->r1 - h->r0) * h->m * sizeof(double);
    if (bytes == 0) return;
    cudaStream_t st = (cudaStream_t)stream;
    if (dB) CA_CUDA(cudaMemcpyAsync(dB, h->B.ptr, bytes, cudaMemcpyDeviceToDevice, st));
    if (dA1) CA_CUDA(cudaMemcpyAsync(dA1, h->A1.ptr, bytes, cudaMemcpyDeviceToDevice, st));
    if (dA2) CA_CUDA(cudaMemcpyAsync(dA2, h->A2.ptr, bytes, cudaMemcpyDeviceToDevice, st));
  });
}

int32_t ca_assembly_get_dissipation_dev(ca_assembly* h, double* dD, void* stream) {
  return guarded([&] {
    if (!h || !dD) fail(CA_ERR_INVALID, "null argument");
    h->bind();
    const size_t bytes = (size_t)(h->r1 - h->r0) * h->m * sizeof(double);
    if (bytes) CA_CUDA(cudaMemcpyAsync(dD, h->D.ptr, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  });
}

int32_t ca_assembly_get_coo_dev(ca_assembly* h, int64_t* d_i, int64_t* d_j, int64_t* d_k, double* d_val,
                                void* stream) {
  return guarded([&] {
    if (!h) fail(CA_ERR_INVALID, "null handle");
    h->bind();
    if (h->nnz == 0) return;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)h->nnz;
    if (d_i) CA_CUDA(cudaMemcpyAsync(d_i, h->oi.ptr, n * 8, cudaMemcpyDeviceToDevice, st));
    if (d_j) CA_CUDA(cudaMemcpyAsync(d_j, h->oj.ptr, n * 8, cudaMemcpyDeviceToDevice, st));
    if (d_k) CA_CUDA(cudaMemcpyAsync(d_k, h->ok.ptr, n * 8, cudaMemcpyDeviceToDevice, st));
    if (d_val) CA_CUDA(cudaMemcpyAsync(d_val, h->oval.ptr, n * 8, cudaMemcpyDeviceToDevice, st));
  });
}

int32_t ca_assembly_get_linear(ca_assembly* h, double* B, double* A1, double* A2) {
  return guarded([&] {
    if (!h) fail(CA_ERR_INVALID, "null handle");
    h->bind();
    const int64_t rows = h->r1 - h->r0, m = h->m;
    std::vector<double> slab((size_t)rows * m);
    double* dst[3] = {B, A1, A2};
    const double* src[3] = {h->B.ptr, h->A1.ptr, h->A2.ptr};
    for (int q = 0; q < 3; ++q) {
      if (!dst[q] || rows == 0) continue;
      CA_CUDA(cudaMemcpy(slab.data(), src[q], slab.size() * sizeof(double), cudaMemcpyDeviceToHost));
      for (int64_t r = 0; r < rows; ++r)
        for (int64_t j = 0; j < m; ++j) dst[q][(h->r0 + r) + m * j] = slab[(size_t)r * m + j];
    }
  });
}

int32_t ca_assembly_get_dissipation(ca_assembly* h, double* D) {
  return guarded([&] {
    if (!h || !D) fail(CA_ERR_INVALID, "null argument");
    h->bind();
    const int64_t rows = h->r1 - h->r0, m = h->m;
    if (rows == 0) return;
    std::vector<double> slab((size_t)rows * m);
    CA_CUDA(cudaMemcpy(slab.data(), h->D.ptr, slab.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (int64_t r = 0; r < rows; ++r)
      for (int64_t j = 0; j < m; ++j) D[(h->r0 + r) + m * j] = slab[(size_t)r * m + j];
  });
}

int32_t ca_assembly_get_coo(ca_assembly* h, int64_t* ijk, double* val) {
  return guarded([&] {
    if (!h) fail(CA_ERR_INVALID, "null handle");
    h->bind();
    const size_t n = (size_t)h->nnz;
    if (n == 0) return;
    if (ijk) {
      CA_CUDA(cudaMemcpy(ijk, h->oi.ptr, n * 8, cudaMemcpyDeviceToHost));
      CA_CUDA(cudaMemcpy(ijk + n, h->oj.ptr, n * 8, cudaMemcpyDeviceToHost));
      CA_CUDA(cudaMemcpy(ijk + 2 * n, h->ok.ptr, n * 8, cudaMemcpyDeviceToHost));
    }
    if (val) CA_CUDA(cudaMemcpy(val, h->oval.ptr, n * 8, cudaMemcpyDeviceToHost));
  });
}

}  // extern "C"
