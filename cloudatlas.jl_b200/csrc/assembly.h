// Result of a Galerkin assembly on the device (shared by assemble.cu, comm.cu and the model build).
#pragma once
#include <vector>

#include "common.h"

struct ca_assembly {
  int device = 0;
  int64_t m = 0, r0 = 0, r1 = 0, nnz = 0;
  double seconds = 0;
  double contract_seconds = 0, contract_flops = 0;  // dense path: the GEMM kernel alone, 2 * rows * m^2 * 3 Nq
  int32_t grid[3] = {0, 0, 0};
  ca::DeviceBuffer<double> B, A1, A2, D;  // (r1-r0) x m row-major; D = <curl, curl> (dissipation)
  ca::DeviceBuffer<int64_t> oi, oj, ok;
  ca::DeviceBuffer<double> oval;
  std::vector<double> shear;              // Psi_shear (ODEModels.jl:65-78), all m modes, with mode_scale applied
  // sharded assembly (comm.cu): the slabs of all ranks gathered; seconds of the collective and bytes received
  double gather_seconds = 0;
  int64_t gather_bytes = 0;
  int32_t world = 1;
  void bind() const { CA_CUDA(cudaSetDevice(device)); }
};

namespace ca {

struct AssembleOpts {
  bool dense = false;
  int nx = 0, ny = 0, nz = 0;
  const double* mode_scale = nullptr;  // m: Psi_n -> s_n Psi_n (the reference's c * Psi, BasisFunctions.jl:417)
  const double* mix = nullptr;         // m x m column-major C: Phi_n = sum_p Psi_p C[p,n] (dense path only)
};

// rows [row_begin, row_end) of B, A1, A2, D and the N entries with i in that range, on the current device
int32_t assemble_slab(double alpha, double gamma, const int64_t* ijkl, int64_t m, int32_t normalize, int64_t row_begin,
                      int64_t row_end, const AssembleOpts& opts, ca_assembly** out);

}  // namespace ca
