// Tabular description of the divergence-free plane-Couette basis (host side).
//
// The reference builds symbolic objects (FourierMode / BasisComponent / BasisFunction,
// src/BasisFunctions.jl:39-51,222-252,378-380); everything basisSet ever produces (:630-702) is a
// product  coeff * E_jx(x) * E_kz(z) * S_l^(d)(y),  so here a basis is three small arrays per mode and
// everything downstream is table lookups:
//   E_j(x) = cos(a|j|x) (j<0), 1 (j=0), sin(a j x) (j>0)          (BasisFunctions.jl:30-42)
//   S_0 = y - y^3/3,  S_l = (1-y^2)^2 P_{l-1}(y)  (l>=1),  d = number of y-derivatives  (:636-642)
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/cloudatlas_b200.h"

namespace ca {

// (i,j,k,l) rows in the reference's order (BasisFunctions.jl:471-476, 489-564), optionally filtered by
// symmetry generators (Symmetries.jl:25-60, parity tables BasisFunctions.jl:847-907).
std::vector<int64_t> basis_indices(int J, int K, int L, const ca_symmetry* H, int nH);  // row-major m x 4

// sigma Psi_ijkl == Psi_ijkl for one generator (Symmetries.jl:25-49); fam = i in 1..6
bool invariant_under_row(const ca_symmetry& g, int fam, int j, int k, int l);

struct Component {
  double coef;  // 0 = component absent
  int32_t jx, kz;
  int32_t l, d;  // wall-normal polynomial S_l differentiated d times
};

struct BasisTables {
  int m = 0, J = 0, K = 0, L = 0;
  double alpha = 0, gamma = 0;
  std::vector<Component> comp;  // [m][3]  (u, v, w)
  // wall-normal tables over pid = 4*l + d, d = 0..3
  int npid = 0, ny = 0;
  std::vector<double> nodes, weights;  // Gauss-Legendre on [-1,1]; weights include the 1/2 of the average
  std::vector<double> F;               // [npid][ny] values of S_l^(d) at the nodes (hi part)
  std::vector<double> Flo, wlo;        // low parts: F + Flo and weights + wlo carry the 64-bit mantissa
                                       // of the host's long-double evaluation (double-double on the GPU)
  std::vector<int8_t> parity;          // [npid] +1 even / -1 odd
  std::vector<double> Y2, Y2y;         // [npid][npid]: <F_p F_q>, <y F_p F_q>  (exact zeros by parity)
  // Fourier tables, index j+J (k+K): pair  <E_a, d^d E_b>  [n][n][2],  triple  <E_a, E_b d^d E_c>  [n][n][n][2]
  std::vector<double> X2, Z2, X3, Z3;
  std::vector<double> wall_avg;        // [npid] (F_p(1) + F_p(-1)) / 2, for the wall shear rate
};

// normalize: scale each mode by 1/sqrt(<Psi,Psi>)  (BasisFunctions.jl:697-699)
// ny_override > 0: number of Gauss-Legendre nodes (must be at least the default exact count)
BasisTables build_tables(double alpha, double gamma, const int64_t* ijkl /*m x 4 column-major*/, int64_t m,
                         bool normalize, int ny_override = 0);

}  // namespace ca
