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
#include <algorithm>
#include <array>
#include <numeric>

#include <memory>
#include <vector>

#include "assembly.h"
#include "basis.h"
#include "common.h"
#include "scan.cuh"
#include "dgemm.cuh"

using namespace ca;

namespace {

constexpr double kAbsZero = 1e-15;  // ODEModels.jl:46
constexpr double kRelZero = 1e-11;  // cancellation-aware: |sum| vs sum|terms|
constexpr double kDenseRelZero = 1e-13;  // dense path: |N| vs max|N| (quadrature noise ~1e-16, true entries >= ~1e-10)
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
  // component-major copies for the lanes of n_kernel (consecutive modes -> consecutive addresses):
  const double* coefT;    // [3][m]
  const uint32_t* packT;  // [3][m]  jx | kz << 10 | pid << 20
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

struct IJ {  // warp-uniform data of one (i,j) pair, kept small (30 registers): the kernel is latency-bound on the
             // per-pair set-up and wants occupancy more than it wants precomputed table pointers
  double ci[3], cj[3];
  int xi[3], xj[3], zi[3], zj[3], pi[3], pj[3];
  unsigned live;  // bitmask over (c, a) of combinations with non-zero coefficient product
};

__device__ __forceinline__ void load_ij(const DevBasis& b, int i, int j, IJ& u) {
  u.live = 0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    u.ci[c] = b.coef[i * 3 + c];
    u.cj[c] = b.coef[j * 3 + c];
    u.xi[c] = b.jx[i * 3 + c];
    u.xj[c] = b.jx[j * 3 + c];
    u.zi[c] = b.kz[i * 3 + c];
    u.zj[c] = b.kz[j * 3 + c];
    u.pi[c] = b.pid[i * 3 + c];
    u.pj[c] = b.pid[j * 3 + c];
  }
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int a = 0; a < 3; ++a)
      if (u.ci[c] * u.cj[a] != 0.0) u.live |= 1u << (c * 3 + a);
}

// predicated load: issues only for pred != 0, and never turns into a branch
__device__ __forceinline__ double ldg_if(const double* p, bool pred) {
  double v = 0.0;
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q ld.global.nc.f64 %0, [%1];\n\t}" : "+d"(v) : "l"(p), "r"((int)pred));
  return v;
}

// Straight-line: the 9 x 2 Fourier factors (small tables, L1) decide which (c, a) combinations survive, the survivors'
// wall-normal integrals (Y3 is megabytes: L2) are fetched with predicated loads, all nine chains in flight together.
// As a chain of "if (x3 == 0) continue; if (z3 == 0) continue; ... Y3 ..." the same lookups were up to 27 dependent
// round trips per candidate, nine of them to L2.  Dead combinations add an exact zero, so the sum is unchanged.
__device__ __forceinline__ bool entry(const DevBasis& b, const IJ& u, int k, double& val) {
  double s = 0.0, mag = 0.0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (!(u.live >> (3 * c) & 7u)) continue;  // warp-uniform: component c of Psi_i (or all of Psi_j) is absent
    const double ck = __ldg(b.coefT + (size_t)c * b.m + k);
    const uint32_t pk3 = __ldg(b.packT + (size_t)c * b.m + k);
    const int xk = pk3 & 1023, zk = (pk3 >> 10) & 1023, pk = pk3 >> 20;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int q = c * 3 + a;
      if (!(u.live >> q & 1)) continue;  // warp-uniform
      // <E_i, E_j d^dx E_k>: derivative flag set for the advection direction a == 0 (x), a == 2 (z), a == 1 (y: pid + 1)
      const double x3 = b.X3[((size_t)(u.xi[c] * b.nx + u.xj[a]) * b.nx + xk) * 2 + (a == 0 ? 1 : 0)];
      const double z3 = b.Z3[((size_t)(u.zi[c] * b.nz + u.zj[a]) * b.nz + zk) * 2 + (a == 2 ? 1 : 0)];
      const bool alive = ck != 0.0 && x3 != 0.0 && z3 != 0.0;
      const double y3 = ldg_if(b.Y3 + ((size_t)(u.pi[c] * b.npid + u.pj[a]) * b.npid) + (a == 1 ? 1 : 0) + pk, alive);
      const double t = alive ? (u.ci[c] * u.cj[a]) * ck * x3 * z3 * y3 : 0.0;
      s -= t;
      mag += fabs(t);
    }
  }
  val = s;
  return fabs(s) > kAbsZero && fabs(s) > kRelZero * mag;
}

// Candidate third factors of a pair (i, j).  The Fourier triple averages vanish unless the wavenumber magnitudes of
// k are |j_i| +- |j_j| in x and |k_i| +- |k_j| in z, so only the modes of at most four (|j|, |k|) groups can
// contribute; the index set lists the modes of one signed (j, k) contiguously, which makes those candidates a few
// short contiguous ranges of k.  Per pair of groups the host stores the candidates as one ascending list (lanes take
// consecutive list positions: a range per signed (j, k) is only a handful of modes and would leave most of a warp
// idle); scanning it in order keeps the (i, j, k) order of the output.  The lists are further split by the wall-normal
// parity classes of i and j: the y integral of an odd integrand vanishes, which removes about half of the remaining
// candidates (46 % of the wavenumber-compatible ones survive otherwise).  (Scanning all m values of k made the O(m^3) selection-rule lookups the
// whole cost of the assembly: 0.62 s at m = 2962.)
struct Candidates {
  const int32_t* group;     // [m] group of every mode: |j| * (K + 1) + |k|
  const int32_t* pair_off;  // [G * G * S * S + 1] offsets into list
  const int32_t* list;      // ascending candidate k of every (group of i, group of j, parity class of i, of j)
  const int32_t* pclass;    // [m] wall-normal parity class of every mode
  int G, S;
};

// Two-pass form, kept for candidate sets too large for a scratch slot each (no pruning: every k is a candidate).
// FILL = false: counts[(i-r0)*m + j] = number of kept k;  FILL = true: writes the entries at offs[...]
template <bool FILL>
__global__ void __launch_bounds__(128)
n_kernel(DevBasis b, Candidates cd, int r0, int r1, int64_t* __restrict__ counts, const int64_t* __restrict__ offs,
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
      const int gp = ((cd.group[i] * cd.G + cd.group[j]) * cd.S + cd.pclass[i]) * cd.S + cd.pclass[j];
      const int qe = cd.pair_off[gp + 1];
      for (int q0 = cd.pair_off[gp]; q0 < qe; q0 += 64) {  // two chunks of 32 candidates, evaluated before either is emitted
        const int qa = q0 + lane, qb = qa + 32;
        const int ka = qa < qe ? __ldg(cd.list + qa) : -1, kb = qb < qe ? __ldg(cd.list + qb) : -1;
        double va = 0.0, vb = 0.0;
        const bool keepa = ka >= 0 && entry(b, u, ka, va);
        const bool keepb = kb >= 0 && entry(b, u, kb, vb);
        const unsigned maska = __ballot_sync(0xffffffffu, keepa), maskb = __ballot_sync(0xffffffffu, keepb);
        if (FILL && keepa) {
          const int64_t at = base + n + __popc(maska & ((1u << lane) - 1));
          oi[at] = i + 1;
          oj[at] = j + 1;
          ok[at] = ka + 1;
          oval[at] = va;
        }
        n += __popc(maska);
        if (FILL && keepb) {
          const int64_t at = base + n + __popc(maskb & ((1u << lane) - 1));
          oi[at] = i + 1;
          oj[at] = j + 1;
          ok[at] = kb + 1;
          oval[at] = vb;
        }
        n += __popc(maskb);
      }
    }
    if (!FILL && lane == 0) counts[t] = n;
  }
}

// Every candidate is evaluated ONCE: the kept entries of a pair (i, j) are written, compacted, into the pair's segment
// of a scratch array (its offset is the running sum of the pairs' candidate counts, known without evaluating anything),
// counts[(i-r0)*m + j] gets their number; after the scan of the counts a copy kernel moves the segments to their final
// places in (i, j, k) order.  (Two evaluating passes -- count, then fill -- cost twice the table lookups.)
__global__ void __launch_bounds__(256)
cand_count_kernel(Candidates cd, int m, int r0, int64_t npairs, int64_t* __restrict__ cnum) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npairs) return;
  const int i = r0 + (int)(t / m), j = (int)(t % m);
  const int gp = ((cd.group[i] * cd.G + cd.group[j]) * cd.S + cd.pclass[i]) * cd.S + cd.pclass[j];
  cnum[t] = cd.pair_off[gp + 1] - cd.pair_off[gp];
}

__global__ void __launch_bounds__(128)
n_eval_kernel(DevBasis b, Candidates cd, int r0, int r1, const int64_t* __restrict__ coffs, int64_t* __restrict__ counts,
              int32_t* __restrict__ tk, double* __restrict__ tv) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t npairs = (int64_t)(r1 - r0) * b.m;
  for (int64_t t = warp; t < npairs; t += nwarps) {
    const int i = r0 + (int)(t / b.m), j = (int)(t % b.m);
    IJ u;
    load_ij(b, i, j, u);
    int64_t n = 0;
    const int64_t base = coffs[t];
    if (u.live) {
      const int gp = ((cd.group[i] * cd.G + cd.group[j]) * cd.S + cd.pclass[i]) * cd.S + cd.pclass[j];
      const int qe = cd.pair_off[gp + 1];
      for (int q0 = cd.pair_off[gp]; q0 < qe; q0 += 64) {  // two chunks of 32 candidates, evaluated before either is emitted
        const int qa = q0 + lane, qb = qa + 32;
        const int ka = qa < qe ? __ldg(cd.list + qa) : -1, kb = qb < qe ? __ldg(cd.list + qb) : -1;
        double va = 0.0, vb = 0.0;
        const bool keepa = ka >= 0 && entry(b, u, ka, va);
        const bool keepb = kb >= 0 && entry(b, u, kb, vb);
        const unsigned maska = __ballot_sync(0xffffffffu, keepa), maskb = __ballot_sync(0xffffffffu, keepb);
        if (keepa) {
          const int64_t at = base + n + __popc(maska & ((1u << lane) - 1));
          tk[at] = ka;
          tv[at] = va;
        }
        n += __popc(maska);
        if (keepb) {
          const int64_t at = base + n + __popc(maskb & ((1u << lane) - 1));
          tk[at] = kb;
          tv[at] = vb;
        }
        n += __popc(maskb);
      }
    }
    if (lane == 0) counts[t] = n;
  }
}

// one warp per pair (i, j): its kept entries from the scratch segment to offs[t] of the output arrays
__global__ void __launch_bounds__(256)
n_copy_kernel(int m, int r0, int64_t npairs, const int64_t* __restrict__ coffs, const int64_t* __restrict__ offs,
              const int32_t* __restrict__ tk, const double* __restrict__ tv, int64_t* __restrict__ oi,
              int64_t* __restrict__ oj, int64_t* __restrict__ ok, double* __restrict__ oval) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t t = warp; t < npairs; t += nwarps) {
    const int64_t dst = offs[t], n = offs[t + 1] - dst, src = coffs[t];
    const int64_t i1 = r0 + t / m + 1, j1 = t % m + 1;
    for (int64_t e = lane; e < n; e += 32) {
      oi[dst + e] = i1;
      oj[dst + e] = j1;
      ok[dst + e] = (int64_t)tk[src + e] + 1;
      oval[dst + e] = tv[src + e];
    }
  }
}

}  // namespace


// ================================================================================================
// Dense quadrature contraction ("K1a"): the same tensor by brute force on the FP64 tensor pipe.
//
//   N_ijk = - sum_q sum_c  W_q Phi_ic(q) * G_jkc(q),      G_jkc(q) = sum_a Phi_ja(q) d_a Phi_kc(q)
//
// on the exact tensor grid (uniform in x and z, Gauss-Legendre in y): a GEMM with M = m (i), N = m^2 (j,k) and
// K = 3 Nq whose B operand is generated on the fly in shared memory and never stored.  One CTA owns a
// 128 (i) x 128 (8 j x 16 k) tile of the output; per chunk of 16 grid points it stages the pre-tiled A block with
// cp.async (double buffered), generates the 48 x 128 B block from 8 x 16 x 3 values of Phi_j and 16 x 16 x 9 of
// grad Phi_k, and its 16 warps issue 16 DMMAs per k-step each (4 x 4 fragments, A and B fragments read
// conflict-free as [k-step][row][4]).  The dense result goes to a scratch slab and is compacted to the
// reference's COO order.  This path exists for the FP64 roofline and as an independent check of the separable
// path; it does ~2000x more arithmetic for the same numbers.
constexpr int kDI = 128, kDK = 16;
// Tile shape of the contraction: an output tile is 128 (i) x DJ * 16 (DJ values of j x 16 of k); a chunk is DQ grid
// points (K = 3 DQ); one thread per (column, quarter of the chunk's k range), i.e. 4 * DJ * 16 threads.
//   DenseWide   128 x 128 tile, 16 warps, one CTA per SM, chunks of 12 points
//   DenseTwin   128 x  64 tile,  8 warps, TWO CTAs per SM, chunks of 8 points: while one CTA drains into its per-chunk
//               barrier the other keeps the tensor pipe fed (the barrier tail was the last large stall of the wide shape)
template <int DJ, int DQ>
struct DenseCfgT {
  static constexpr int J = DJ, Q = DQ, Cols = DJ * kDK, Kc = 3 * DQ, Threads = 4 * Cols, Warps = Threads / 32;
  static constexpr int WC = Cols / 32;          // warps along the columns; Warps / WC = 4 along the rows
  static constexpr int Bld = Cols + 4;          // row stride of the generated B block: the 4 k-rows of a fragment land 8 banks apart
  static constexpr int KS = Kc / 4;             // k-steps per chunk = values of B each thread generates per chunk
  static constexpr int CTAs = Threads <= 256 ? 2 : 1;
  static constexpr size_t SmemDoubles = 2 * ((size_t)Kc * kDI + (size_t)Kc * Bld + DJ * DQ * 3 + kDK * DQ * 9);
  static_assert(KS % 3 == 0, "a thread's k range must start on a grid point");
};
using DenseWide = DenseCfgT<8, 12>;
using DenseTwin = DenseCfgT<4, 8>;
// DenseRing: 128 x 128 tile, 16 warps, one CTA per SM, chunks of 8 points in a THREE-stage ring with one mbarrier per
// stage instead of a CTA barrier per chunk (dense_contract_ring_kernel)
struct DenseRing : DenseCfgT<8, 8> {
  static constexpr int Stages = 3;
  static constexpr size_t StageDoubles = (size_t)Kc * kDI + (size_t)Kc * Bld + J * Q * 3 + kDK * Q * 9;
  static constexpr size_t SmemDoubles = Stages * StageDoubles + Stages;  // + the mbarriers
};

struct DenseTables {
  int m, mpad, nq, nchunks;
  const double* U;   // [mpad/8][nchunks][16 q][3 a][8 modes]        Phi_{n,a}(q): one contiguous block per (j-tile, chunk)
  const double* G9;  // [mpad/16][nchunks][16 q][9 (a*3+c)][16 modes]  d_a Phi_{n,c}(q): mode index fastest, so that the
                     //                                               B generation reads shared memory conflict-free
  const double* A;   // [mpad/128][nchunks][12][128][4]   W_q Phi_{i,c}(q), k = q_local*3 + c
};

// one thread per (mode, grid point): fills U, G9 and the tiled A
template <class Cfg>
__global__ void dense_tables_kernel(DevBasis b, int mpad, int nxg, int nyg, int nzg, int nq,
                                    const double* __restrict__ Ex, const double* __restrict__ Exd,
                                    const double* __restrict__ Ez, const double* __restrict__ Ezd,
                                    const double* __restrict__ F, int ny_tab, const double* __restrict__ wq,
                                    double* __restrict__ U, double* __restrict__ G9, double* __restrict__ A,
                                    int nchunks, double* __restrict__ raw, const double* __restrict__ mixed) {
  // raw != null: the 12 planes (u_c, d_a u_c) are written as raw[(plane * nq + q) * m + n] for the mixing GEMM and
  // nothing else; mixed != null: the planes are read from there (Phi = Psi C) instead of being evaluated
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)mpad * nq) return;
  const int n = (int)(t / nq), q = (int)(t % nq);
  double u[3] = {0, 0, 0}, g[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  double w = 0.0;
  const int nreal = nxg * nyg * nzg;
  if (mixed) {
    if (n < b.m && q < nreal) {
      w = wq[q];
      for (int c = 0; c < 3; ++c) u[c] = mixed[((int64_t)c * nq + q) * b.m + n];
      for (int e = 0; e < 9; ++e) g[e] = mixed[((int64_t)(3 + e) * nq + q) * b.m + n];
    }
  } else if (n < b.m && q < nreal) {
    const int ix = q / (nyg * nzg), iy = (q / nzg) % nyg, iz = q % nzg;
    w = wq[q];
    for (int c = 0; c < 3; ++c) {
      const double cf = b.coef[n * 3 + c];
      if (cf == 0.0) continue;
      const int jx = b.jx[n * 3 + c], kz = b.kz[n * 3 + c], pid = b.pid[n * 3 + c];
      const double ex = Ex[jx * nxg + ix], exd = Exd[jx * nxg + ix], ez = Ez[kz * nzg + iz], ezd = Ezd[kz * nzg + iz];
      const double f = F[pid * ny_tab + iy], fd = F[(pid + 1) * ny_tab + iy];
      u[c] = cf * ex * ez * f;
      g[0 * 3 + c] = cf * exd * ez * f;
      g[1 * 3 + c] = cf * ex * ez * fd;
      g[2 * 3 + c] = cf * ex * ezd * f;
    }
  }
  if (raw) {
    if (n < b.m) {
      for (int c = 0; c < 3; ++c) raw[((int64_t)c * nq + q) * b.m + n] = u[c];
      for (int e = 0; e < 9; ++e) raw[((int64_t)(3 + e) * nq + q) * b.m + n] = g[e];
    }
    return;
  }
  constexpr int kDJ = Cfg::J, kDQ = Cfg::Q, kDKc = Cfg::Kc;
  const int it = n / kDI, il = n % kDI, chunk = q / kDQ, ql = q % kDQ;
  for (int c = 0; c < 3; ++c)
    U[((((size_t)(n / kDJ) * nchunks + chunk) * kDQ + ql) * 3 + c) * kDJ + n % kDJ] = u[c];
  for (int e = 0; e < 9; ++e)
    G9[((((size_t)(n / kDK) * nchunks + chunk) * kDQ + ql) * 9 + e) * kDK + n % kDK] = g[e];
  for (int c = 0; c < 3; ++c) {
    const int k = ql * 3 + c;
    A[((((size_t)it * nchunks + chunk) * (kDKc / 4) + k / 4) * kDI + il) * 4 + k % 4] = w * u[c];
  }
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}

// Software pipeline (round 2; round 1 generated B and multiplied in two barrier-separated phases with a 4-way
// conflicted B store: DMMA pipe 74 % busy, 0.71 of peak).  Per chunk of kDQ = 12 grid points (K = 36, nine k-steps):
//   * A(c), B(c) are consumed by the DMMA loop while every thread generates ITS nine values of B(c+1), one per k-step,
//     from the tables of chunk c+1 -- generation (2.3 % of the flops) hides under the tensor work instead of stalling it;
//   * cp.async brings A(c+1) and the tables of chunk c+2 during the same iteration;
//   * ONE barrier per chunk publishes all three.
// B is stored [k][kDBld] (row of 128 columns padded to 132 doubles): the generating warp writes 32 consecutive doubles
// (conflict-free), and the four k-rows of a B fragment start 8 banks apart, so fragment loads are conflict-free too.
template <class Cfg>
__global__ void __launch_bounds__(Cfg::Threads, Cfg::CTAs)
dense_contract_kernel(DenseTables T, int it0, int njt, double* __restrict__ Nd /* [rows][m][m] */, int row0, int row1) {
  constexpr int kDJ = Cfg::J, kDQ = Cfg::Q, kDKc = Cfg::Kc, kDCols = Cfg::Cols, kDThreads = Cfg::Threads, kDBld = Cfg::Bld;
  constexpr int KS = Cfg::KS;
  extern __shared__ __align__(16) double sm[];
  double* Abuf = sm;                                  // [2][9 k-steps][128 rows][4]
  double* Bbuf = Abuf + 2 * (kDKc * kDI);             // [2][36 k][132]
  double* Utb = Bbuf + 2 * (kDKc * kDBld);            // [2][12 q][3 a][8 modes]
  double* Gtb = Utb + 2 * (kDJ * kDQ * 3);            // [2][12 q][9 (a*3+c)][16 modes]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, r = lane >> 2, c = lane & 3;
  const int it = it0 + blockIdx.y;
  const int jt = blockIdx.x % njt, kt = blockIdx.x / njt;
  const int j0 = jt * kDJ, k0 = kt * kDK;
  const int wi = warp / Cfg::WC, wc = warp % Cfg::WC;
  const int nchunks = T.nchunks;

  auto stage_A = [&](int chunk) {
    const double* src = T.A + ((size_t)it * nchunks + chunk) * (kDKc * kDI);
    double* dst = Abuf + (chunk & 1) * (kDKc * kDI);
    for (int e = tid; e < kDKc * kDI / 2; e += kDThreads) cp_async16(dst + 2 * e, src + 2 * e);
  };
  auto stage_tables = [&](int chunk) {
    // the (j-tile, chunk) block of U and the (k-tile, chunk) block of G9 are contiguous in global memory
    const double* usrc = T.U + ((size_t)jt * nchunks + chunk) * (kDJ * kDQ * 3);
    for (int e = tid; e < kDJ * kDQ * 3 / 2; e += kDThreads) cp_async16(Utb + (chunk & 1) * (kDJ * kDQ * 3) + 2 * e, usrc + 2 * e);
    const double* gsrc = T.G9 + ((size_t)kt * nchunks + chunk) * (kDK * kDQ * 9);
    for (int e = tid; e < kDK * kDQ * 9 / 2; e += kDThreads) cp_async16(Gtb + (chunk & 1) * (kDK * kDQ * 9) + 2 * e, gsrc + 2 * e);
  };
  // this thread's share of a B block: column col, k in [KS kg, KS kg + KS) = KS / 3 grid points, all three c
  const int col = tid % kDCols, kg = tid / kDCols;
  const int jj = col / kDK, kk = col % kDK;
  // value `step` of chunk `chunk`: B[k = KS kg + step][col] = sum_a U[q][a][jj] * G9[q][a*3+cc][kk]
  auto generate = [&](int chunk, int step, double& u0, double& u1, double& u2) {
    const int q = kg * (KS / 3) + step / 3, cc = step % 3;
    const double* u = Utb + (chunk & 1) * (kDJ * kDQ * 3) + jj;
    const double* g = Gtb + (chunk & 1) * (kDK * kDQ * 9) + kk;
    if (cc == 0) {
      u0 = u[(q * 3) * kDJ];
      u1 = u[(q * 3 + 1) * kDJ];
      u2 = u[(q * 3 + 2) * kDJ];
    }
    const double v = u0 * g[(q * 9 + cc) * kDK] + u1 * g[(q * 9 + 3 + cc) * kDK] + u2 * g[(q * 9 + 6 + cc) * kDK];
    Bbuf[(chunk & 1) * (kDKc * kDBld) + (kg * KS + step) * kDBld + col] = v;
  };

  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

  // prologue: tables(0), A(0); B(0) generated in full; tables(1) in flight
  stage_tables(0);
  stage_A(0);
  asm volatile("cp.async.commit_group;\n" ::: "memory");
  asm volatile("cp.async.wait_all;\n" ::: "memory");
  __syncthreads();
  if (nchunks > 1) stage_tables(1);
  asm volatile("cp.async.commit_group;\n" ::: "memory");
  {
    double u0 = 0, u1 = 0, u2 = 0;
#pragma unroll
    for (int step = 0; step < KS; ++step) generate(0, step, u0, u1, u2);
  }
  asm volatile("cp.async.wait_all;\n" ::: "memory");
  __syncthreads();

  for (int chunk = 0; chunk < nchunks; ++chunk) {
    // here: A(chunk), B(chunk) and the tables of chunk+1 are in shared memory and visible
    if (chunk + 1 < nchunks) stage_A(chunk + 1);       // into the buffer A(chunk-1) has left
    if (chunk + 2 < nchunks) stage_tables(chunk + 2);  // into the buffer the tables of `chunk` have left
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    const bool gen = chunk + 1 < nchunks;
    const double* Ab = Abuf + (chunk & 1) * (kDKc * kDI);
    const double* Bb = Bbuf + (chunk & 1) * (kDKc * kDBld);
    double u0 = 0, u1 = 0, u2 = 0;
#pragma unroll
    for (int ks = 0; ks < kDKc / 4; ++ks) {
      double af[4], bf[4];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) af[mt] = Ab[(ks * kDI + wi * 32 + mt * 8 + r) * 4 + c];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) bf[nt] = Bb[(ks * 4 + c) * kDBld + wc * 32 + nt * 8 + r];
      if (gen) generate(chunk + 1, ks, u0, u1, u2);
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                       : "+d"(acc[mt][nt][0]), "+d"(acc[mt][nt][1])
                       : "d"(af[mt]), "d"(bf[nt]));
        }
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    __syncthreads();  // A(chunk+1), tables(chunk+2), B(chunk+1) published; everyone is done with A(chunk), B(chunk)
  }
  // epilogue: N = -C into the dense slab
  const int m = T.m;
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
    const int i = it * kDI + wi * 32 + mt * 8 + r;
    if (i >= m || i < row0 || i >= row1) continue;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int ocol = wc * 32 + nt * 8 + 2 * c + e;
        const int j = j0 + ocol / kDK, k = k0 + ocol % kDK;
        if (j < m && k < m) Nd[((size_t)(i - row0) * m + j) * m + k] = -acc[mt][nt][e];
      }
  }
}

// The same contraction without a CTA barrier in the loop.  In the two-buffer kernel above every chunk ends in a
// __syncthreads (B(c+1) is written by all threads and read by all warps): ncu put 9 % of the warp time into that
// barrier, with the DMMA pipe 84.6 % busy.  Here the operands live in a ring of three stages with one mbarrier each:
//   during chunk c a thread  issues the cp.async of A(c+1) and of the tables of chunk c+2 (arrival on barrier (c+1)
//                            when those copies have landed: cp.async.mbarrier.arrive.noinc),
//                            generates its values of B(c+1) in the FIRST HALF of the chunk's k-steps and then arrives
//                            on barrier (c+1) (release), and multiplies A(c), B(c) in all k-steps;
//   before chunk c+1         it waits on barrier (c+1) -- which needs every thread past the MIDDLE of chunk c only, so
//                            warps may drift by half a chunk without anyone waiting.
// Stage (c+1) mod 3 held chunk c-2: a thread that has passed barrier (c) knows that everyone is past the middle of
// chunk c-1, hence done reading chunk c-2.
__device__ __forceinline__ void dense_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void dense_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void dense_mbar_arrive_on_copies(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void dense_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done)
                 : "r"(a), "r"(parity)
                 : "memory");
  } while (!done);
}

__global__ void __launch_bounds__(DenseRing::Threads, 1)
dense_contract_ring_kernel(DenseTables T, int it0, int njt, double* __restrict__ Nd /* [rows][m][m] */, int row0, int row1) {
  using Cfg = DenseRing;
  constexpr int kDJ = Cfg::J, kDQ = Cfg::Q, kDKc = Cfg::Kc, kDCols = Cfg::Cols, kDThreads = Cfg::Threads, kDBld = Cfg::Bld;
  constexpr int KS = Cfg::KS, NS = Cfg::Stages, kHalf = KS / 2;  // k-steps per chunk; generation happens in the first kHalf
  static_assert(KS % 2 == 0 && (KS / kHalf) * kHalf == KS, "values per generating k-step");
  constexpr int kPer = KS / kHalf;  // B values a thread generates per k-step of the first half
  extern __shared__ __align__(16) double sm[];
  double* Abuf = sm;                                 // [NS][KS][128 rows][4]
  double* Bbuf = Abuf + NS * (kDKc * kDI);           // [NS][Kc][Bld]
  double* Utb = Bbuf + NS * (kDKc * kDBld);          // [NS][Q][3 a][J modes]
  double* Gtb = Utb + NS * (kDJ * kDQ * 3);          // [NS][Q][9 (a*3+c)][16 modes]
  uint64_t* bar = reinterpret_cast<uint64_t*>(Gtb + NS * (kDK * kDQ * 9));  // [NS]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, r = lane >> 2, c = lane & 3;
  const int it = it0 + blockIdx.y;
  const int jt = blockIdx.x % njt, kt = blockIdx.x / njt;
  const int j0 = jt * kDJ, k0 = kt * kDK;
  const int wi = warp / Cfg::WC, wc = warp % Cfg::WC;
  const int nchunks = T.nchunks;

  auto stage_A = [&](int chunk) {
    const double* src = T.A + ((size_t)it * nchunks + chunk) * (kDKc * kDI);
    double* dst = Abuf + (chunk % NS) * (kDKc * kDI);
    for (int e = tid; e < kDKc * kDI / 2; e += kDThreads) cp_async16(dst + 2 * e, src + 2 * e);
  };
  auto stage_tables = [&](int chunk) {
    const double* usrc = T.U + ((size_t)jt * nchunks + chunk) * (kDJ * kDQ * 3);
    for (int e = tid; e < kDJ * kDQ * 3 / 2; e += kDThreads) cp_async16(Utb + (chunk % NS) * (kDJ * kDQ * 3) + 2 * e, usrc + 2 * e);
    const double* gsrc = T.G9 + ((size_t)kt * nchunks + chunk) * (kDK * kDQ * 9);
    for (int e = tid; e < kDK * kDQ * 9 / 2; e += kDThreads) cp_async16(Gtb + (chunk % NS) * (kDK * kDQ * 9) + 2 * e, gsrc + 2 * e);
  };
  const int col = tid % kDCols, kg = tid / kDCols;
  const int jj = col / kDK, kk = col % kDK;
  auto generate = [&](int stage, int step, double& u0, double& u1, double& u2) {
    const int q = kg * (KS / 3) + step / 3, cc = step % 3;
    const double* u = Utb + stage * (kDJ * kDQ * 3) + jj;
    const double* g = Gtb + stage * (kDK * kDQ * 9) + kk;
    if (cc == 0) {
      u0 = u[(q * 3) * kDJ];
      u1 = u[(q * 3 + 1) * kDJ];
      u2 = u[(q * 3 + 2) * kDJ];
    }
    const double v = u0 * g[(q * 9 + cc) * kDK] + u1 * g[(q * 9 + 3 + cc) * kDK] + u2 * g[(q * 9 + 6 + cc) * kDK];
    Bbuf[stage * (kDKc * kDBld) + (kg * KS + step) * kDBld + col] = v;
  };

  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

  // prologue: barriers; tables(0) and A(0); then tables(1) in flight while B(0) is generated in full; all of it
  // arrives on barrier 0
  if (tid == 0)
    for (int s = 0; s < NS; ++s) dense_mbar_init(&bar[s], 2 * kDThreads);  // per thread: its copies and its B values
  stage_tables(0);
  stage_A(0);
  asm volatile("cp.async.commit_group;\n" ::: "memory");
  asm volatile("cp.async.wait_all;\n" ::: "memory");
  __syncthreads();
  if (nchunks > 1) stage_tables(1);
  dense_mbar_arrive_on_copies(&bar[0]);
  {
    double u0 = 0, u1 = 0, u2 = 0;
#pragma unroll
    for (int step = 0; step < KS; ++step) generate(0, step, u0, u1, u2);
  }
  dense_mbar_arrive(&bar[0]);

  int stage = 0;        // chunk % NS
  uint32_t parity = 0;  // (chunk / NS) & 1
  for (int chunk = 0; chunk < nchunks; ++chunk) {
    dense_mbar_wait(&bar[stage], parity);  // A(chunk), B(chunk), tables(chunk+1) are in shared memory and visible
    const int nstage = stage + 1 == NS ? 0 : stage + 1;
    if (chunk + 1 < nchunks) stage_A(chunk + 1);       // into the stage chunk-2 has left
    if (chunk + 2 < nchunks) stage_tables(chunk + 2);  // into the stage whose tables (of chunk-1) were last read in chunk-2
    dense_mbar_arrive_on_copies(&bar[nstage]);
    const bool gen = chunk + 1 < nchunks;
    const double* Ab = Abuf + stage * (kDKc * kDI);
    const double* Bb = Bbuf + stage * (kDKc * kDBld);
    double u0 = 0, u1 = 0, u2 = 0;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      double af[4], bf[4];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) af[mt] = Ab[(ks * kDI + wi * 32 + mt * 8 + r) * 4 + c];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) bf[nt] = Bb[(ks * 4 + c) * kDBld + wc * 32 + nt * 8 + r];
      if (ks < kHalf && gen) {
#pragma unroll
        for (int e = 0; e < kPer; ++e) generate(nstage, ks * kPer + e, u0, u1, u2);
      }
      if (ks == kHalf) dense_mbar_arrive(&bar[nstage]);  // this thread's part of B(chunk+1) is written
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                       : "+d"(acc[mt][nt][0]), "+d"(acc[mt][nt][1])
                       : "d"(af[mt]), "d"(bf[nt]));
        }
    }
    stage = nstage;
    if (stage == 0) parity ^= 1u;
  }
  asm volatile("cp.async.wait_all;\n" ::: "memory");
  // epilogue: N = -C into the dense slab
  const int m = T.m;
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
    const int i = it * kDI + wi * 32 + mt * 8 + r;
    if (i >= m || i < row0 || i >= row1) continue;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int ocol = wc * 32 + nt * 8 + 2 * c + e;
        const int j = j0 + ocol / kDK, k = k0 + ocol % kDK;
        if (j < m && k < m) Nd[((size_t)(i - row0) * m + j) * m + k] = -acc[mt][nt][e];
      }
  }
}

__global__ void absmax_kernel(const double* __restrict__ v, int64_t n, double* __restrict__ out) {
  double mx = 0.0;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
    mx = fmax(mx, fabs(v[t]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > 0.0)
    atomicMax((unsigned long long*)out, (unsigned long long)__double_as_longlong(mx));  // non-negative doubles order as integers
}

// compaction of the dense slab to COO in (i,j,k) order: same two-pass structure as n_kernel
template <bool FILL>
__global__ void __launch_bounds__(256)
dense_compact_kernel(const double* __restrict__ Nd, int m, int r0, int r1, double thr, int64_t* __restrict__ counts,
                     const int64_t* __restrict__ offs, int64_t* __restrict__ oi, int64_t* __restrict__ oj,
                     int64_t* __restrict__ ok, double* __restrict__ oval) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t npairs = (int64_t)(r1 - r0) * m;
  for (int64_t t = warp; t < npairs; t += nwarps) {
    const int i = r0 + (int)(t / m), j = (int)(t % m);
    const double* row = Nd + (size_t)t * m;
    int64_t n = 0;
    const int64_t base = FILL ? offs[t] : 0;
    for (int k0 = 0; k0 < m; k0 += 32) {
      const int k = k0 + lane;
      const double v = k < m ? row[k] : 0.0;
      const bool keep = fabs(v) > thr;
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
    if (!FILL && lane == 0) counts[t] = n;
  }
}

int32_t ca::assemble_slab(double alpha, double gamma, const int64_t* ijkl, int64_t m64, int32_t normalize,
                          int64_t row_begin, int64_t row_end, const AssembleOpts& dopt, ca_assembly** out) {
  return guarded([&] {
    if (!out) fail(CA_ERR_INVALID, "ca_assemble: out is null");
    *out = nullptr;
    if (m64 <= 0 || !ijkl) fail(CA_ERR_INVALID, "ca_assemble: empty index set");
    if (row_begin < 0 || row_end > m64 || row_begin > row_end)
      fail(CA_ERR_INVALID, "ca_assemble: row range [%lld,%lld) outside 0:%lld", (long long)row_begin, (long long)row_end, (long long)m64);
    if (m64 > 46000) fail(CA_ERR_UNSUPPORTED, "ca_assemble: m too large");
    BasisTables T = build_tables(alpha, gamma, ijkl, m64, normalize != 0, dopt.dense ? dopt.ny : 0);
    if (dopt.mix && !dopt.dense)
      fail(CA_ERR_INVALID, "ca_assemble: a mixed basis (Phi = Psi C) has no factorised modes; use the dense quadrature path");
    if (dopt.mode_scale)
      for (int64_t n = 0; n < m64; ++n) {
        if (!(dopt.mode_scale[n] != 0.0) || !std::isfinite(dopt.mode_scale[n]))
          fail(CA_ERR_INVALID, "ca_assemble: mode_scale[%lld] must be finite and non-zero", (long long)n);
        for (int c = 0; c < 3; ++c) T.comp[(size_t)n * 3 + c].coef *= dopt.mode_scale[n];
      }
    require_device();
    std::unique_ptr<ca_assembly> h(new ca_assembly);
    CA_CUDA(cudaGetDevice(&h->device));
    const int m = (int)m64, r0 = (int)row_begin, r1 = (int)row_end, rows = r1 - r0;
    h->m = m;
    h->r0 = r0;
    h->r1 = r1;
    h->shear.assign((size_t)m, 0.0);
    for (int n = 0; n < m; ++n) {
      const Component& u = T.comp[(size_t)n * 3];  // the u component (mode_scale already applied)
      if (u.coef != 0.0 && u.jx == 0 && u.kz == 0) h->shear[(size_t)n] = u.coef * T.wall_avg[4 * u.l + u.d + 1];
    }
    if (dopt.mix) {  // Phi_n = sum_p Psi_p C[p,n]: the shear functional is linear in the mode
      std::vector<double> mixed((size_t)m, 0.0);
      for (int n = 0; n < m; ++n)
        for (int p2 = 0; p2 < m; ++p2) mixed[(size_t)n] += h->shear[(size_t)p2] * dopt.mix[p2 + (size_t)m * n];
      h->shear.swap(mixed);
    }

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
    std::vector<double> coefT((size_t)m * 3);
    std::vector<uint32_t> packT((size_t)m * 3);
    for (int n = 0; n < m; ++n)
      for (int c = 0; c < 3; ++c) {
        const size_t at = (size_t)n * 3 + c;
        if (jx[at] < 0 || jx[at] >= 1024 || kz[at] < 0 || kz[at] >= 1024 || pid[at] < 0 || pid[at] >= 4096)
          fail(CA_ERR_UNSUPPORTED, "ca_assemble: J, K < 512 and L < 1024 are required");
        coefT[(size_t)c * m + n] = coef[at];
        packT[(size_t)c * m + n] = (uint32_t)jx[at] | ((uint32_t)kz[at] << 10) | ((uint32_t)pid[at] << 20);
      }
    DeviceBuffer<double> dcoefT;
    DeviceBuffer<uint32_t> dpackT;
    dcoefT.upload(coefT.data(), coefT.size());
    dpackT.upload(packT.data(), packT.size());
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

    // ---- candidate ranges of the third factor per pair of wavenumber groups (see Candidates)
    const int ngz = T.K + 1, G = (T.J + 1) * ngz;
    std::vector<int32_t> group((size_t)m, 0);
    bool uniform = getenv("CA_ASSEMBLE_NOPRUNE") == nullptr;  // A/B hook: scan every k
    for (int n = 0; n < m && uniform; ++n) {
      int jm = -1, km = -1;
      for (int c = 0; c < 3; ++c) {
        const Component& cp = T.comp[(size_t)n * 3 + c];
        if (cp.coef == 0.0) continue;
        const int aj = std::abs(cp.jx), ak = std::abs(cp.kz);
        if (jm >= 0 && (aj != jm || ak != km)) uniform = false;  // not a single-wavenumber mode: no pruning
        jm = aj;
        km = ak;
      }
      if (jm > T.J || km > T.K) uniform = false;
      group[n] = jm < 0 ? 0 : jm * ngz + km;
    }
    // Wall-normal parity class of a mode.  In the divergence-free families u and w have one parity and v the other, so
    // a mode is "even" (u, v, w = +1, -1, +1) or "odd"; if some mode does not follow that pattern, the classes are the
    // distinct signatures (parity of the u, v, w polynomials, 0 = absent) as long as there are only a handful.
    std::vector<int32_t> pclass((size_t)m, 0);
    std::vector<std::array<int, 3>> sigs;
    bool by_parity = uniform && getenv("CA_ASSEMBLE_NOPARITY") == nullptr;
    auto signature = [&](int n) {
      std::array<int, 3> sg{0, 0, 0};
      for (int c = 0; c < 3; ++c) {
        const Component& cp = T.comp[(size_t)n * 3 + c];
        if (cp.coef != 0.0) sg[c] = T.parity[4 * cp.l + cp.d];
      }
      return sg;
    };
    bool binary = by_parity;
    for (int n = 0; n < m && binary; ++n) {
      const std::array<int, 3> sg = signature(n);
      const int P = sg[0] != 0 ? sg[0] : (sg[2] != 0 ? sg[2] : -sg[1]);
      if (P == 0 || (sg[0] != 0 && sg[0] != P) || (sg[1] != 0 && sg[1] != -P) || (sg[2] != 0 && sg[2] != P)) binary = false;
      pclass[n] = P > 0 ? 0 : 1;
    }
    if (binary) {
      sigs = {std::array<int, 3>{1, -1, 1}, std::array<int, 3>{-1, 1, -1}};
    } else {
      for (int n = 0; n < m && by_parity; ++n) {
        const std::array<int, 3> sg = signature(n);
        size_t at = std::find(sigs.begin(), sigs.end(), sg) - sigs.begin();
        if (at == sigs.size()) sigs.push_back(sg);
        pclass[n] = (int32_t)at;
        if (sigs.size() > 6) by_parity = false;  // the table grows with S^2: not worth it beyond a handful of classes
      }
    }
    if (!by_parity) {
      sigs.assign(1, std::array<int, 3>{0, 0, 0});
      std::fill(pclass.begin(), pclass.end(), 0);
    }
    const int S = (int)sigs.size();
    // can <Psi_i^c, (Psi_j^a d_a) Psi_k^c> survive the y integral for some live (c, a)?  d/dy flips the parity of k's factor
    auto parity_allows = [&](int si, int sj, int sk) {
      if (!by_parity) return true;
      for (int c = 0; c < 3; ++c)
        for (int a = 0; a < 3; ++a)
          if (sigs[si][c] != 0 && sigs[sj][a] != 0 && sigs[sk][c] != 0 &&
              sigs[si][c] * sigs[sj][a] * sigs[sk][c] * (a == 1 ? -1 : 1) == 1)
            return true;
      return false;
    };
    std::vector<int32_t> pair_off, list;
    int Geff = G;
    if (!uniform) {
      Geff = 1;
      std::fill(group.begin(), group.end(), 0);
      pair_off = {0, m};
      list.resize((size_t)m);
      std::iota(list.begin(), list.end(), 0);
    } else {
      std::vector<std::vector<int2>> of_group((size_t)G);  // contiguous index ranges of every group
      for (int n = 0; n < m;) {
        int e = n;
        while (e < m && group[e] == group[n]) ++e;
        of_group[group[n]].push_back(make_int2(n, e));
        n = e;
      }
      pair_off.assign((size_t)G * G * S * S + 1, 0);
      std::vector<int32_t> cands;
      for (int ga = 0; ga < G; ++ga)
        for (int gb = 0; gb < G; ++gb) {
          const int ja = ga / ngz, ka = ga % ngz, jb = gb / ngz, kb = gb % ngz;
          std::vector<int> gs;
          for (int jm : {ja + jb, std::abs(ja - jb)})
            for (int km : {ka + kb, std::abs(ka - kb)})
              if (jm <= T.J && km <= T.K) gs.push_back(jm * ngz + km);
          std::sort(gs.begin(), gs.end());
          gs.erase(std::unique(gs.begin(), gs.end()), gs.end());
          std::vector<int2> rs;
          for (int g : gs) rs.insert(rs.end(), of_group[g].begin(), of_group[g].end());
          std::sort(rs.begin(), rs.end(), [](const int2& x, const int2& y) { return x.x < y.x; });
          cands.clear();
          for (const int2& r : rs)
            for (int k = r.x; k < r.y; ++k) cands.push_back(k);
          for (int si = 0; si < S; ++si)
            for (int sj = 0; sj < S; ++sj) {
              for (int32_t k : cands)
                if (parity_allows(si, sj, pclass[k])) list.push_back(k);
              pair_off[(((size_t)ga * G + gb) * S + si) * S + sj + 1] = (int32_t)list.size();
            }
        }
    }
    DeviceBuffer<int32_t> dgroup, dpair_off, dlist, dpclass;
    dgroup.upload(group.data(), group.size());
    dpair_off.upload(pair_off.data(), pair_off.size());
    dlist.upload(list.data(), list.size());
    dpclass.upload(pclass.data(), pclass.size());
    const Candidates cand{dgroup.ptr, dpair_off.ptr, dlist.ptr, dpclass.ptr, Geff, uniform ? S : 1};

    // Device time = the kernels of the two passes (tables + count + scan, then fill).  Buffers whose size is known are
    // allocated before the first event and the output arrays between the two timed spans: cudaMalloc is synchronous
    // and its cost varies by milliseconds from call to call.
    const int64_t nij_pre = (int64_t)rows * m;
    DeviceBuffer<int64_t> counts, offs;
    ScanScratch scan_tmp;
    if (nij_pre > 0) {
      h->B.alloc((size_t)nij_pre);
      h->A1.alloc((size_t)nij_pre);
      h->A2.alloc((size_t)nij_pre);
      h->D.alloc((size_t)nij_pre);
      counts.alloc((size_t)nij_pre + 1);
      offs.alloc((size_t)nij_pre + 1);
      scan_tmp.sums.alloc((size_t)(nij_pre + 1) / kScanTile + kScanTile + 16);  // allocated outside the timed spans
    }
    // separable path: scratch for the single evaluating pass -- one slot per candidate, counted on the host from the
    // list lengths (modes per (group, parity class) x list length of the class pair)
    bool single_pass = false;
    DeviceBuffer<int64_t> cnum, coffs;
    DeviceBuffer<int32_t> tk;
    DeviceBuffer<double> tv;
    if (nij_pre > 0 && !dopt.dense) {
      const int Sx = uniform ? S : 1;
      std::vector<int64_t> members((size_t)Geff * Sx, 0);
      for (int n = 0; n < m; ++n) members[(size_t)group[n] * Sx + pclass[n]]++;
      int64_t ncand = 0;
      for (int i = r0; i < r1; ++i) {
        const size_t ci = (size_t)group[i] * Sx + pclass[i];
        for (int gb = 0; gb < Geff; ++gb)
          for (int sj = 0; sj < Sx; ++sj) {
            const size_t gp = (((size_t)group[i] * Geff + gb) * Sx + pclass[i]) * Sx + sj;
            ncand += members[(size_t)gb * Sx + sj] * (int64_t)(pair_off[gp + 1] - pair_off[gp]);
          }
        (void)ci;
      }
      single_pass = ncand * 12 <= ((int64_t)16 << 30);  // otherwise: two evaluating passes, no scratch
      if (single_pass) {
        cnum.alloc((size_t)nij_pre + 1);
        coffs.alloc((size_t)nij_pre + 1);
        tk.alloc((size_t)std::max<int64_t>(ncand, 1));
        tv.alloc((size_t)std::max<int64_t>(ncand, 1));
      }
    }
    cudaEvent_t e0, e1, e2, e3;
    CA_CUDA(cudaEventCreate(&e0));
    CA_CUDA(cudaEventCreate(&e1));
    CA_CUDA(cudaEventCreate(&e2));
    CA_CUDA(cudaEventCreate(&e3));
    bool fill_timed = false;
    CA_CUDA(cudaEventRecord(e0));
    y3_kernel<<<(unsigned)((ny3 + 255) / 256), 256>>>(dF.ptr, dFlo.ptr, dw.ptr, dwlo.ptr, dpar.ptr, T.npid, T.ny, dY3.ptr);
    CA_CUDA(cudaGetLastError());
    count_launch();

    DevBasis b{m, T.J, T.K, T.npid, 2 * T.J + 1, 2 * T.K + 1, T.alpha, T.gamma, dcoef.ptr, djx.ptr, dkz.ptr, dpid.ptr, dk2.ptr,
               dX2.ptr, dZ2.ptr, dX3.ptr, dZ3.ptr, dY2.ptr, dY2y.ptr, dY3.ptr, dcoefT.ptr, dpackT.ptr};
    const int64_t nij = (int64_t)rows * m;
    if (nij > 0) {
      DeviceBuffer<double> dmix;
      if (dopt.mix) {
        // B' = C' B C (likewise A1, A2, D): the four matrices of the plain basis in full, then two small GEMMs each
        dmix.upload(dopt.mix, (size_t)m * m);
        DeviceBuffer<double> full[4], t1;
        for (auto& f : full) f.alloc((size_t)m * m);
        t1.alloc((size_t)m * m);
        linear_kernel<<<(unsigned)(((int64_t)m * m + 255) / 256), 256>>>(b, 0, m, full[0].ptr, full[1].ptr, full[2].ptr, full[3].ptr);
        CA_CUDA(cudaGetLastError());
        count_launch();
        double* dst[4] = {h->B.ptr, h->A1.ptr, h->A2.ptr, h->D.ptr};
        for (int q = 0; q < 4; ++q) {
          launch_dgemm(m, m, m, 1.0, rowmajor(full[q].ptr, m), colmajor(dmix.ptr, m), 0.0, t1.ptr, m, 1, nullptr);
          launch_dgemm(rows, m, m, 1.0, MatView{dmix.ptr + (size_t)m * r0, m, 1}, rowmajor(t1.ptr, m), 0.0, dst[q], m, 1, nullptr);
        }
        CA_CUDA(cudaStreamSynchronize(nullptr));
      } else {
        linear_kernel<<<(unsigned)((nij + 255) / 256), 256>>>(b, r0, r1, h->B.ptr, h->A1.ptr, h->A2.ptr, h->D.ptr);
        CA_CUDA(cudaGetLastError());
        count_launch();
      }

      CA_CUDA(cudaMemsetAsync(counts.ptr, 0, ((size_t)nij + 1) * sizeof(int64_t)));
      int sms = 0;
      CA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
      const unsigned grid = (unsigned)std::min<int64_t>((nij + 7) / 8, (int64_t)sms * 16);
      // separable path: 128-thread CTAs (92 registers: five CTAs = 20 warps per SM instead of two x eight)
      const unsigned ngrid = (unsigned)std::min<int64_t>((nij + 3) / 4, (int64_t)sms * 40);
      DeviceBuffer<double> Nd;  // dense path: [rows][m][m] scratch slab
      double thr = 0.0;
      if (dopt.dense) {
        // ---- tensor grid: uniform in x, z (exact for trigonometric degree 3J, 3K), Gauss-Legendre in y
        const int nxg = dopt.nx > 0 ? dopt.nx : 3 * T.J + 1, nzg = dopt.nz > 0 ? dopt.nz : 3 * T.K + 1, nyg = T.ny;
        if (nxg <= 3 * T.J || nzg <= 3 * T.K)
          fail(CA_ERR_INVALID, "grid %d x %d x %d is not exact for J, K = %d, %d (need Nx > 3J, Nz > 3K)", nxg, nyg, nzg, T.J, T.K);
        // tile shape of the contraction: R (default) the three-stage ring, T the twin two-CTA shape, W the wide one
        const char shape = getenv("CA_DENSE_CFG") ? getenv("CA_DENSE_CFG")[0] : 'R';
        auto with_cfg = [&](auto&& f) {
          if (shape == 'T') f(DenseTwin{});
          else if (shape == 'W') f(DenseWide{});
          else f(DenseRing{});
        };
        int kDQ = 0, kDJ = 0;
        with_cfg([&](auto cfg) { kDQ = decltype(cfg)::Q; kDJ = decltype(cfg)::J; });
        const int nreal = nxg * nyg * nzg, nq = (nreal + kDQ - 1) / kDQ * kDQ, nchunks = nq / kDQ;
        const int mpad = (m + kDI - 1) / kDI * kDI;
        if ((double)rows * m * m * 8.0 > 100e9)
          fail(CA_ERR_UNSUPPORTED, "dense scratch for %d rows of a %d-mode basis exceeds 100 GB: assemble in row slabs", rows, m);
        h->grid[0] = nxg; h->grid[1] = nyg; h->grid[2] = nzg;
        const double two_pi = 6.283185307179586476925286766559;
        auto trig_table = [&](int Jmax, int n, double wn, std::vector<double>& e, std::vector<double>& ed) {
          e.assign((size_t)(2 * Jmax + 1) * n, 0.0);
          ed.assign((size_t)(2 * Jmax + 1) * n, 0.0);
          for (int js = 0; js <= 2 * Jmax; ++js)
            for (int ix = 0; ix < n; ++ix) {
              const int j = js - Jmax;
              const double th = two_pi * ix / n;
              double v = 1.0, d = 0.0;
              if (j < 0) { v = std::cos(-j * th); d = -wn * (-j) * std::sin(-j * th); }
              if (j > 0) { v = std::sin(j * th); d = wn * j * std::cos(j * th); }
              e[(size_t)js * n + ix] = v;
              ed[(size_t)js * n + ix] = d;
            }
        };
        std::vector<double> ex, exd, ez, ezd, wq((size_t)nq, 0.0);
        trig_table(T.J, nxg, T.alpha, ex, exd);
        trig_table(T.K, nzg, T.gamma, ez, ezd);
        for (int q = 0; q < nreal; ++q) wq[q] = T.weights[(q / nzg) % nyg] / ((double)nxg * nzg);
        DeviceBuffer<double> dex, dexd, dez, dezd, dwq, dU, dG9, dA, dmax;
        dex.upload(ex.data(), ex.size());
        dexd.upload(exd.data(), exd.size());
        dez.upload(ez.data(), ez.size());
        dezd.upload(ezd.data(), ezd.size());
        dwq.upload(wq.data(), wq.size());
        dU.alloc((size_t)mpad * nq * 3);
        dG9.alloc((size_t)mpad * nq * 9);
        dA.alloc((size_t)mpad * nq * 3);
        dmax.alloc(1);
        Nd.alloc((size_t)rows * m * m);
        CA_CUDA(cudaMemsetAsync(dmax.ptr, 0, sizeof(double)));
        const int64_t ntab = (int64_t)mpad * nq;
        auto tables = [&](double* raw, const double* mixed) {
          with_cfg([&](auto cfg) {
            dense_tables_kernel<decltype(cfg)><<<(unsigned)((ntab + 255) / 256), 256>>>(b, mpad, nxg, nyg, nzg, nq, dex.ptr, dexd.ptr, dez.ptr,
                                                                                      dezd.ptr, dF.ptr, T.ny, dwq.ptr, dU.ptr, dG9.ptr, dA.ptr,
                                                                                      nchunks, raw, mixed);
          });
          CA_CUDA(cudaGetLastError());
          count_launch();
        };
        if (dopt.mix) {
          // Phi = Psi C on the grid: 12 planes (u_c, d_a u_c) of the plain basis, one GEMM [12 nq x m] . [m x m], then the
          // tiled operand layouts from the mixed planes.  ~0.5 % of the contraction's arithmetic.
          DeviceBuffer<double> raw, mixed;
          raw.alloc((size_t)12 * nq * m);
          mixed.alloc((size_t)12 * nq * m);
          tables(raw.ptr, nullptr);
          launch_dgemm(12 * nq, m, m, 1.0, rowmajor(raw.ptr, m), colmajor(dmix.ptr, m), 0.0, mixed.ptr, m, 1, nullptr);
          tables(nullptr, mixed.ptr);
          CA_CUDA(cudaStreamSynchronize(nullptr));  // raw / mixed are released here
        } else {
          tables(nullptr, nullptr);
        }
        DenseTables DT{m, mpad, nq, nchunks, dU.ptr, dG9.ptr, dA.ptr};
        const int it0 = r0 / kDI, nit = (r1 + kDI - 1) / kDI - it0;
        const int njt = (m + kDJ - 1) / kDJ, nkt = (m + kDK - 1) / kDK;
        cudaEvent_t c0, c1;
        CA_CUDA(cudaEventCreate(&c0));
        CA_CUDA(cudaEventCreate(&c1));
        const dim3 cgrid((unsigned)(njt * nkt), (unsigned)nit);
        if (shape == 'T' || shape == 'W') {
          with_cfg([&](auto cfg) {
            using Cfg = decltype(cfg);
            if constexpr (!std::is_same<Cfg, DenseRing>::value) {
              const size_t smem = Cfg::SmemDoubles * sizeof(double);
              CA_CUDA(cudaFuncSetAttribute(dense_contract_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
              CA_CUDA(cudaEventRecord(c0));
              dense_contract_kernel<Cfg><<<cgrid, Cfg::Threads, smem>>>(DT, it0, njt, Nd.ptr, r0, r1);
            }
          });
        } else {
          const size_t smem = DenseRing::SmemDoubles * sizeof(double);
          CA_CUDA(cudaFuncSetAttribute(dense_contract_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          CA_CUDA(cudaEventRecord(c0));
          dense_contract_ring_kernel<<<cgrid, DenseRing::Threads, smem>>>(DT, it0, njt, Nd.ptr, r0, r1);
        }
        CA_CUDA(cudaGetLastError());
        count_launch();
        CA_CUDA(cudaEventRecord(c1));
        CA_CUDA(cudaEventSynchronize(c1));
        float cms = 0;
        CA_CUDA(cudaEventElapsedTime(&cms, c0, c1));
        cudaEventDestroy(c0);
        cudaEventDestroy(c1);
        h->contract_seconds = cms * 1e-3;
        h->contract_flops = 2.0 * rows * (double)m * m * 3.0 * nreal;
        absmax_kernel<<<sms * 8, 256>>>(Nd.ptr, (int64_t)rows * m * m, dmax.ptr);
        CA_CUDA(cudaGetLastError());
        count_launch();
        double gmax = 0.0;
        CA_CUDA(cudaMemcpy(&gmax, dmax.ptr, sizeof(double), cudaMemcpyDeviceToHost));
        thr = std::max(kAbsZero, kDenseRelZero * gmax);
        dense_compact_kernel<false><<<grid, 256>>>(Nd.ptr, m, r0, r1, thr, counts.ptr, nullptr, nullptr, nullptr, nullptr, nullptr);
      } else if (!single_pass) {
        n_kernel<false><<<ngrid, 128>>>(b, cand, r0, r1, counts.ptr, nullptr, nullptr, nullptr, nullptr, nullptr);
      } else {
        CA_CUDA(cudaMemsetAsync(cnum.ptr + nij, 0, sizeof(int64_t)));
        cand_count_kernel<<<(unsigned)((nij + 255) / 256), 256>>>(cand, m, r0, nij, cnum.ptr);
        CA_CUDA(cudaGetLastError());
        count_launch();
        exclusive_scan(cnum.ptr, coffs.ptr, nij + 1, scan_tmp, nullptr);
        n_eval_kernel<<<ngrid, 128>>>(b, cand, r0, r1, coffs.ptr, counts.ptr, tk.ptr, tv.ptr);
      }
      CA_CUDA(cudaGetLastError());
      count_launch();
      exclusive_scan(counts.ptr, offs.ptr, nij + 1, scan_tmp, nullptr);
      CA_CUDA(cudaEventRecord(e1));
      int64_t nnz = 0;
      CA_CUDA(cudaMemcpy(&nnz, offs.ptr + nij, sizeof(int64_t), cudaMemcpyDeviceToHost));
      h->nnz = nnz;
      if (nnz > 0) {
        h->oi.alloc((size_t)nnz);
        h->oj.alloc((size_t)nnz);
        h->ok.alloc((size_t)nnz);
        h->oval.alloc((size_t)nnz);
        CA_CUDA(cudaEventRecord(e2));
        if (dopt.dense)
          dense_compact_kernel<true><<<grid, 256>>>(Nd.ptr, m, r0, r1, thr, nullptr, offs.ptr, h->oi.ptr, h->oj.ptr, h->ok.ptr, h->oval.ptr);
        else if (!single_pass)
          n_kernel<true><<<ngrid, 128>>>(b, cand, r0, r1, nullptr, offs.ptr, h->oi.ptr, h->oj.ptr, h->ok.ptr, h->oval.ptr);
        else
          n_copy_kernel<<<(unsigned)std::min<int64_t>((nij + 7) / 8, (int64_t)sms * 32), 256>>>(
              m, r0, nij, coffs.ptr, offs.ptr, tk.ptr, tv.ptr, h->oi.ptr, h->oj.ptr, h->ok.ptr, h->oval.ptr);
        CA_CUDA(cudaGetLastError());
        count_launch();
        CA_CUDA(cudaEventRecord(e3));
        fill_timed = true;
      }
    } else {
      CA_CUDA(cudaEventRecord(e1));
    }
    CA_CUDA(cudaDeviceSynchronize());
    float ms = 0, ms2 = 0;
    CA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (fill_timed) CA_CUDA(cudaEventElapsedTime(&ms2, e2, e3));
    h->seconds = (ms + ms2) * 1e-3;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaEventDestroy(e2);
    cudaEventDestroy(e3);
    *out = h.release();
  });
}

extern "C" {

int32_t ca_assemble(double alpha, double gamma, const int64_t* ijkl, int64_t m, int32_t normalize,
                    int64_t row_begin, int64_t row_end, ca_assembly** out) {
  return assemble_slab(alpha, gamma, ijkl, m, normalize, row_begin, row_end, AssembleOpts{}, out);
}

int32_t ca_assemble_dense(double alpha, double gamma, const int64_t* ijkl, int64_t m, int32_t normalize,
                          int64_t row_begin, int64_t row_end, int32_t nx, int32_t ny, int32_t nz, ca_assembly** out) {
  AssembleOpts d;
  d.dense = true;
  d.nx = nx;
  d.ny = ny;
  d.nz = nz;
  return assemble_slab(alpha, gamma, ijkl, m, normalize, row_begin, row_end, d, out);
}

int32_t ca_assemble_basis(double alpha, double gamma, const int64_t* ijkl, int64_t m, int32_t normalize,
                          const double* mode_scale, const double* mix, int32_t dense, int32_t nx, int32_t ny, int32_t nz,
                          int64_t row_begin, int64_t row_end, ca_assembly** out) {
  AssembleOpts d;
  d.dense = dense != 0;
  d.nx = nx;
  d.ny = ny;
  d.nz = nz;
  d.mode_scale = mode_scale;
  d.mix = mix;
  return assemble_slab(alpha, gamma, ijkl, m, normalize, row_begin, row_end, d, out);
}

int32_t ca_assembly_dense_info(const ca_assembly* h, int32_t* grid3, double* contract_seconds, double* contract_flops) {
  return guarded([&] {
    if (!h) fail(CA_ERR_INVALID, "null handle");
    if (grid3) { grid3[0] = h->grid[0]; grid3[1] = h->grid[1]; grid3[2] = h->grid[2]; }
    if (contract_seconds) *contract_seconds = h->contract_seconds;
    if (contract_flops) *contract_flops = h->contract_flops;
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
