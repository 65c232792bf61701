// Model-specialised ensemble kernel for small m (<= kJitMaxM): one thread per state, the state's
// coefficients streamed through a per-thread cp.async ring into registers, the operator baked into the
// instruction stream (NVRTC, sm_100a).
//
// For m = 17 the ensemble RHS is HBM-bound (272 B and ~1.3 kFLOP per state); the DMMA row-tile
// kernel wastes most of its tensor work on padding at such sizes (fill 18 %), whereas straight-line code
// with x in registers runs at the DFMA rate with no index traffic at all.  The generated source is plain
// CUDA C++ (kept in `source` for inspection); it is compiled once per operator.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "common.h"
#include "pack.h"

namespace ca {

constexpr int kJitMaxM = 48;
constexpr int64_t kJitMinStates = 4096;

struct SmallKernel {
  int m = 0;
  bool ready = false, failed = false;
  std::string source, log;
  cudaLibrary_t lib = nullptr;
  cudaKernel_t kern = nullptr;
  double* dLC = nullptr;       // device address of the module's __constant__ linear coefficients
  std::vector<int32_t> lin_i, lin_j;  // pattern of the linear part (row, column), order of LC
  int threads = 0, stages = 0, ctas_per_sm = 2;  // CTA shape of the generated kernel (jit_small.cu: small_shape)
  size_t smem_bytes = 0;
  size_t cubin_bytes = 0;      // > 0 once NVRTC has compiled the generated source
  double cur_invR = 0.0;
  bool lc_set = false;
  ~SmallKernel();
  // quad: symmetrised terms (a <= b < m); linear pattern may be empty.  Returns false (and sets `log`)
  // when NVRTC is not available or compilation fails -- callers then use the generic kernels.
  bool build(const std::vector<Term>& quad, const std::vector<int32_t>& li, const std::vector<int32_t>& lj, int m);
  // coefficients of the linear part in pattern order (host pointer, nlin doubles)
  void set_linear(const double* lc, cudaStream_t st);
  void launch(const double* dX, int64_t nstates, int64_t ldx, double* dOUT, int64_t ldo, cudaStream_t st) const;
};

}  // namespace ca
