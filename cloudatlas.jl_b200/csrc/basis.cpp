#include "basis.h"

#include <cmath>

#include "common.h"

namespace ca {

namespace {

// The six mode families: which (j,k) carry them and their lowest wall-normal index.
struct Family {
  bool need_j0, need_jnz, need_k0, need_knz;
  int lmin;
};
constexpr Family kFamilies[6] = {
    {true, false, false, false, 0},   // 1: [E_k S_l'; 0; 0]                      j = 0
    {true, false, false, true, 1},    // 2: [0; gk E_k S_l; E_-k S_l']            j = 0, k != 0
    {false, false, true, false, 0},   // 3: [0; 0; E_j S_l']                      k = 0
    {false, true, true, false, 1},    // 4: [E_-j S_l'; aj E_j S_l; 0]            k = 0, j != 0
    {false, true, false, true, 0},    // 5: [gk E_-j E_k S_l'; 0; -aj E_j E_-k S_l']
    {false, true, false, true, 1},    // 6: [gk E_-j E_k S_l'; 2agjk E_j E_k S_l; aj E_j E_-k S_l']
};

inline bool family_present(const Family& f, int j, int k) {
  if (f.need_j0 && j != 0) return false;
  if (f.need_jnz && j == 0) return false;
  if (f.need_k0 && k != 0) return false;
  if (f.need_knz && k == 0) return false;
  return true;
}

// wave index sequence 0, -1, 1, -2, 2, ...
inline int wave_at(int n) { return n == 0 ? 0 : ((n & 1) ? -(n + 1) / 2 : n / 2); }

inline int even_sign(int n) { return (n % 2 == 0) ? 1 : -1; }  // C++ % truncates like Julia's

int reflection_sign(int axis, int fam, int j, int k, int l) {
  switch (axis) {
    case 0:  // x
      if (fam == 1) return -1;
      if (fam == 2) return 1;
      return j <= 0 ? 1 : -1;
    case 1:  // y
      return even_sign(l);
    default:  // z
      if (fam == 3) return -1;
      if (fam == 4) return 1;
      return k <= 0 ? 1 : -1;
  }
}

bool invariant_under(const ca_symmetry& g, int fam, int j, int k, int l) {
  int r = g.s;
  if (g.sx == -1) r *= reflection_sign(0, fam, j, k, l);
  if (g.sy == -1) r *= reflection_sign(1, fam, j, k, l);
  if (g.sz == -1) r *= reflection_sign(2, fam, j, k, l);
  if (g.ax2 == 1) r *= even_sign(j);
  else if (g.ax2 != 0) fail(CA_ERR_INVALID, "symmetries only implemented for half-box shifts");
  if (g.az2 == 1) r *= even_sign(k);
  else if (g.az2 != 0) fail(CA_ERR_INVALID, "symmetries only implemented for half-box shifts");
  return r == 1;
}

// <prod of up to three real Fourier modes> as a box average, by exact uniform quadrature; the values are
// multiples of 1/4, so rounding removes the quadrature's rounding noise entirely.
long double trig(int j, long double th) {
  if (j < 0) return cosl(-j * th);
  if (j > 0) return sinl(j * th);
  return 1.0L;
}

// trig(j, theta_n) for j = -J..J at the npts quadrature nodes, evaluated once (sinl / cosl dominated the table build:
// 0.5 s at J = K = 8 when every average recomputed them)
struct TrigNodes {
  int J, npts;
  std::vector<long double> v;  // [2J+1][npts]
  TrigNodes(int J_, int npts_) : J(J_), npts(npts_), v((size_t)(2 * J_ + 1) * npts_) {
    const long double two_pi = 6.283185307179586476925286766559L;
    for (int j = -J; j <= J; ++j)
      for (int n = 0; n < npts; ++n) v[(size_t)(j + J) * npts + n] = trig(j, two_pi * (n + 0.25L) / npts);
  }
  const long double* of(int j) const { return v.data() + (size_t)(j + J) * npts; }
};

double trig_average(const TrigNodes& t, int a, int b, int c) {
  const long double *ta = t.of(a), *tb = t.of(b), *tc = t.of(c);
  long double s = 0;
  for (int n = 0; n < t.npts; ++n) s += ta[n] * tb[n] * tc[n];
  return (double)(roundl(4 * s / t.npts) / 4);
}

void fourier_tables(double wn, int J, std::vector<double>& pair2, std::vector<double>& triple3) {
  const int n = 2 * J + 1, npts = 6 * J + 8;
  const TrigNodes t(J, npts);
  pair2.assign((size_t)n * n * 2, 0.0);
  triple3.assign((size_t)n * n * n * 2, 0.0);
  // d/dx E_c = wn * c * E_{-c}
  for (int a = -J; a <= J; ++a)
    for (int b = -J; b <= J; ++b) {
      pair2[((size_t)(a + J) * n + (b + J)) * 2 + 0] = trig_average(t, a, b, 0);
      pair2[((size_t)(a + J) * n + (b + J)) * 2 + 1] = wn * b * trig_average(t, a, -b, 0);
      for (int c = -J; c <= J; ++c) {
        const size_t at = (((size_t)(a + J) * n + (b + J)) * n + (c + J)) * 2;
        triple3[at + 0] = trig_average(t, a, b, c);
        triple3[at + 1] = wn * c * trig_average(t, a, b, -c);
      }
    }
}

void gauss_legendre(int n, std::vector<long double>& x, std::vector<long double>& w) {
  x.resize(n);
  w.resize(n);
  const long double pi = 3.141592653589793238462643383279503L;
  for (int i = 0; i < (n + 1) / 2; ++i) {
    long double z = cosl(pi * (i + 0.75L) / (n + 0.5L)), pp = 0;
    for (int it = 0; it < 100; ++it) {
      long double p1 = 1, p2 = 0;
      for (int j = 1; j <= n; ++j) {
        const long double p3 = p2;
        p2 = p1;
        p1 = ((2 * j - 1) * z * p2 - (j - 1) * p3) / j;
      }
      pp = n * (z * p1 - p2) / (z * z - 1);
      const long double dz = p1 / pp;
      z -= dz;
      if (fabsl(dz) < 1e-19L) break;
    }
    x[i] = -z;
    x[n - 1 - i] = z;
    w[i] = w[n - 1 - i] = 2 / ((1 - z * z) * pp * pp);
  }
  if (n & 1) x[n / 2] = 0.0L;
}

}  // namespace

bool invariant_under_row(const ca_symmetry& g, int fam, int j, int k, int l) { return invariant_under(g, fam, j, k, l); }

std::vector<int64_t> basis_indices(int J, int K, int L, const ca_symmetry* H, int nH) {
  if (J < 0 || K < 0 || L < 0) fail(CA_ERR_INVALID, "basisIndices: J, K, L must be non-negative");
  std::vector<int64_t> out;
  for (int jn = 0; jn <= 2 * J; ++jn)
    for (int kn = 0; kn <= 2 * K; ++kn) {
      const int j = wave_at(jn), k = wave_at(kn);
      for (int f = 0; f < 6; ++f) {
        if (!family_present(kFamilies[f], j, k)) continue;
        for (int l = kFamilies[f].lmin; l <= L; ++l) {
          bool keep = true;
          for (int g = 0; g < nH && keep; ++g) keep = invariant_under(H[g], f + 1, j, k, l);
          if (keep) {
            out.push_back(f + 1);
            out.push_back(j);
            out.push_back(k);
            out.push_back(l);
          }
        }
      }
    }
  return out;
}

BasisTables build_tables(double alpha, double gamma, const int64_t* ijkl, int64_t m64, bool normalize,
                         int ny_override) {
  if (m64 <= 0 || !ijkl) fail(CA_ERR_INVALID, "basis: empty index set");
  BasisTables T;
  const int m = T.m = (int)m64;
  T.alpha = alpha;
  T.gamma = gamma;
  for (int n = 0; n < m; ++n) {
    const int64_t f = ijkl[n], j = ijkl[n + m64], k = ijkl[n + 2 * m64], l = ijkl[n + 3 * m64];
    if (f < 1 || f > 6) fail(CA_ERR_INVALID, "invalid index i = %lld not in 1:6", (long long)f);
    if (l < 0) fail(CA_ERR_INVALID, "negative wall-normal index in row %d", n + 1);
    T.J = std::max<int>(T.J, (int)std::llabs(j));
    T.K = std::max<int>(T.K, (int)std::llabs(k));
    T.L = std::max<int>(T.L, (int)l);
  }
  T.comp.assign((size_t)m * 3, Component{0.0, 0, 0, 0, 0});
  for (int n = 0; n < m; ++n) {
    const int f = (int)ijkl[n], j = (int)ijkl[n + m64], k = (int)ijkl[n + 2 * m64], l = (int)ijkl[n + 3 * m64];
    Component* c = &T.comp[(size_t)n * 3];
    const double gk = gamma * k, aj = alpha * j;
    switch (f) {
      case 1: c[0] = {1.0, 0, k, l, 1}; break;
      case 2: c[1] = {gk, 0, k, l, 0}; c[2] = {1.0, 0, -k, l, 1}; break;
      case 3: c[2] = {1.0, j, 0, l, 1}; break;
      case 4: c[0] = {1.0, -j, 0, l, 1}; c[1] = {aj, j, 0, l, 0}; break;
      case 5: c[0] = {gk, -j, k, l, 1}; c[2] = {-aj, j, -k, l, 1}; break;
      case 6: c[0] = {gk, -j, k, l, 1}; c[1] = {2 * alpha * gamma * j * k, j, k, l, 0}; c[2] = {aj, j, -k, l, 1}; break;
    }
  }

  // ---- wall-normal node tables: S_l^(d), d = 0..3, by Legendre recurrences (no monomial coefficients)
  const int L = T.L;
  T.npid = 4 * (L + 1);
  T.ny = (3 * (L + 3)) / 2 + 2;  // exact for degree 3(L+3)+1 <= 2 ny - 1
  if (ny_override > 0) {
    if (ny_override < T.ny) fail(CA_ERR_INVALID, "ny = %d is not exact for L = %d (need >= %d)", ny_override, L, T.ny);
    T.ny = ny_override;
  }
  std::vector<long double> yl, wl;
  gauss_legendre(T.ny, yl, wl);
  yl.push_back(1.0L);   // two extra evaluation points (not quadrature nodes): the walls y = +1, -1
  yl.push_back(-1.0L);
  for (auto& w : wl) w *= 0.5L;  // inner products are (1/2) int_{-1}^{1}   (BasisFunctions.jl:284)
  const int nyw = T.ny + 2;  // nodes + the two walls
  std::vector<long double> Fw((size_t)T.npid * 2, 0.0L);
  std::vector<long double> Fl((size_t)T.npid * T.ny, 0.0L);
  T.parity.assign(T.npid, 0);
  for (int l = 0; l <= L; ++l)
    for (int d = 0; d < 4; ++d) T.parity[4 * l + d] = (int8_t)(((l + 1 + d) % 2 == 0) ? 1 : -1);
  for (int qq = 0; qq < nyw; ++qq) {
    const long double y = yl[qq];
    const bool wall = qq >= T.ny;
    const int q = wall ? qq - T.ny : qq;
    long double* dst = wall ? Fw.data() : Fl.data();
    const int stride = wall ? 2 : T.ny;
    // Legendre values and first three derivatives: P'_n = P'_{n-2} + (2n-1) P_{n-1}, same one order up
    std::vector<long double> P(L + 1, 0), P1(L + 1, 0), P2(L + 1, 0), P3(L + 1, 0);
    P[0] = 1;
    if (L >= 1) { P[1] = y; P1[1] = 1; }
    for (int n = 2; n <= L; ++n) {
      P[n] = ((2 * n - 1) * y * P[n - 1] - (n - 1) * P[n - 2]) / n;
      P1[n] = P1[n - 2] + (2 * n - 1) * P[n - 1];
      P2[n] = P2[n - 2] + (2 * n - 1) * P1[n - 1];
      P3[n] = P3[n - 2] + (2 * n - 1) * P2[n - 1];
    }
    const long double u0 = (1 - y * y) * (1 - y * y), u1 = -4 * y + 4 * y * y * y, u2 = -4 + 12 * y * y, u3 = 24 * y;
    dst[(size_t)0 * stride + q] = y - y * y * y / 3;
    dst[(size_t)1 * stride + q] = 1 - y * y;
    dst[(size_t)2 * stride + q] = -2 * y;
    dst[(size_t)3 * stride + q] = -2.0L;
    for (int l = 1; l <= L; ++l) {
      const long double p = P[l - 1], p1 = P1[l - 1], p2 = P2[l - 1], p3 = P3[l - 1];
      dst[(size_t)(4 * l + 0) * stride + q] = u0 * p;
      dst[(size_t)(4 * l + 1) * stride + q] = u1 * p + u0 * p1;
      dst[(size_t)(4 * l + 2) * stride + q] = u2 * p + 2 * u1 * p1 + u0 * p2;
      dst[(size_t)(4 * l + 3) * stride + q] = u3 * p + 3 * u2 * p1 + 3 * u1 * p2 + u0 * p3;
    }
  }
  T.wall_avg.resize(T.npid);
  for (int p = 0; p < T.npid; ++p) T.wall_avg[p] = (double)((Fw[(size_t)p * 2] + Fw[(size_t)p * 2 + 1]) / 2);
  T.nodes.resize(T.ny);
  T.weights.resize(T.ny);
  T.wlo.resize(T.ny);
  for (int q = 0; q < T.ny; ++q) {
    T.nodes[q] = (double)yl[q];
    T.weights[q] = (double)wl[q];
    T.wlo[q] = (double)(wl[q] - (long double)T.weights[q]);
  }
  T.F.resize(Fl.size());
  T.Flo.resize(Fl.size());
  for (size_t e = 0; e < Fl.size(); ++e) {
    T.F[e] = (double)Fl[e];
    T.Flo[e] = (double)(Fl[e] - (long double)T.F[e]);
  }
  T.Y2.assign((size_t)T.npid * T.npid, 0.0);
  T.Y2y.assign((size_t)T.npid * T.npid, 0.0);
  for (int p = 0; p < T.npid; ++p)
    for (int q = 0; q < T.npid; ++q) {
      long double s = 0, sy = 0, mag = 0, magy = 0;
      for (int n = 0; n < T.ny; ++n) {
        const long double t = wl[n] * Fl[(size_t)p * T.ny + n] * Fl[(size_t)q * T.ny + n];
        s += t;
        sy += t * yl[n];
        mag += fabsl(t);
        magy += fabsl(t * yl[n]);
      }
      // analytic zeros that are not parity zeros come out as rounding noise of the quadrature: flush them
      if (fabsl(s) <= 1e-16L * mag) s = 0;
      if (fabsl(sy) <= 1e-16L * magy) sy = 0;
      const int par = T.parity[p] * T.parity[q];
      T.Y2[(size_t)p * T.npid + q] = par == 1 ? (double)s : 0.0;    // odd integrand: exactly zero
      T.Y2y[(size_t)p * T.npid + q] = par == -1 ? (double)sy : 0.0;
    }
  fourier_tables(alpha, T.J, T.X2, T.X3);
  fourier_tables(gamma, T.K, T.Z2, T.Z3);

  if (normalize) {
    const int nx = 2 * T.J + 1, nz = 2 * T.K + 1;
    for (int n = 0; n < m; ++n) {
      Component* c = &T.comp[(size_t)n * 3];
      double nn = 0;
      for (int a = 0; a < 3; ++a) {
        if (c[a].coef == 0) continue;
        const int pid = 4 * c[a].l + c[a].d;
        nn += c[a].coef * c[a].coef * T.X2[((size_t)(c[a].jx + T.J) * nx + (c[a].jx + T.J)) * 2] *
              T.Z2[((size_t)(c[a].kz + T.K) * nz + (c[a].kz + T.K)) * 2] * T.Y2[(size_t)pid * T.npid + pid];
      }
      const double s = 1 / std::sqrt(nn);
      for (int a = 0; a < 3; ++a) c[a].coef = s * c[a].coef;
    }
  }
  return T;
}

}  // namespace ca

extern "C" {

int32_t ca_basis_indices(int32_t J, int32_t K, int32_t L, const ca_symmetry* H, int32_t nH,
                         int64_t* ijkl_out, int64_t* m_inout) {
  return ca::guarded([&] {
    if (!m_inout) ca::fail(CA_ERR_INVALID, "ca_basis_indices: m_inout is null");
    if (nH < 0 || (nH > 0 && !H)) ca::fail(CA_ERR_INVALID, "ca_basis_indices: bad symmetry list");
    std::vector<int64_t> rows = ca::basis_indices(J, K, L, H, nH);
    const int64_t m = (int64_t)rows.size() / 4;
    if (ijkl_out) {
      if (*m_inout < m) ca::fail(CA_ERR_INVALID, "ca_basis_indices: capacity %lld < %lld rows", (long long)*m_inout, (long long)m);
      for (int64_t n = 0; n < m; ++n)
        for (int c = 0; c < 4; ++c) ijkl_out[n + c * m] = rows[(size_t)n * 4 + c];
    }
    *m_inout = m;
  });
}

int32_t ca_basis_shear(double alpha, double gamma, const int64_t* ijkl, int64_t m, int32_t normalize,
                       double* shear_out) {
  return ca::guarded([&] {
    if (!shear_out) ca::fail(CA_ERR_INVALID, "ca_basis_shear: null output");
    ca::BasisTables T = ca::build_tables(alpha, gamma, ijkl, m, normalize != 0);
    for (int64_t n = 0; n < m; ++n) {
      const ca::Component& u = T.comp[(size_t)n * 3];  // the u component
      shear_out[n] = (u.coef != 0.0 && u.jx == 0 && u.kz == 0) ? u.coef * T.wall_avg[4 * u.l + u.d + 1] : 0.0;
    }
  });
}

int32_t ca_symmetric(const int64_t* ijkl4, const ca_symmetry* H, int32_t nH, int32_t* out) {
  return ca::guarded([&] {
    if (!ijkl4 || !out || nH < 0 || (nH > 0 && !H)) ca::fail(CA_ERR_INVALID, "ca_symmetric: bad arguments");
    const int fam = (int)ijkl4[0], j = (int)ijkl4[1], k = (int)ijkl4[2], l = (int)ijkl4[3];
    *out = 1;
    for (int n = 0; n < nH; ++n) {
      // the reference raises for i outside 1:6 only where a parity table is consulted (x / z reflections)
      if ((fam < 1 || fam > 6) && (H[n].sx == -1 || H[n].sz == -1))
        ca::fail(CA_ERR_INVALID, "invalid index i = %d not in 1:6", fam);
      if (!ca::invariant_under_row(H[n], fam, j, k, l)) {
        *out = 0;
        return;
      }
    }
  });
}

int32_t ca_basis_components(double alpha, double gamma, const int64_t* ijkl, int64_t m, int32_t normalize,
                            double* coef, int32_t* jx, int32_t* kz, int32_t* l, int32_t* d) {
  return ca::guarded([&] {
    ca::BasisTables T = ca::build_tables(alpha, gamma, ijkl, m, normalize != 0);
    for (int64_t n = 0; n < m; ++n)
      for (int c = 0; c < 3; ++c) {
        const ca::Component& q = T.comp[(size_t)n * 3 + c];
        if (coef) coef[n + c * m] = q.coef;
        if (jx) jx[n + c * m] = q.jx;
        if (kz) kz[n + c * m] = q.kz;
        if (l) l[n + c * m] = q.l;
        if (d) d[n + c * m] = q.d;
      }
  });
}

}  // extern "C"
