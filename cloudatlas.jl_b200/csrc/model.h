// The device side of an ODEModel (shared by api_model.cu and hookstep.cu).
#pragma once
#include <cmath>

#include "bilinear.h"
#include "jit_small.h"

struct ca_model {
  int device = 0;
  int64_t m = 0;
  int64_t nblocks = 0, nnz_lin = 0;
  ca::DevicePack op;           // symmetric pack of Q plus the linear slots
  ca::JacobianPack jac;        // of Q, built on first use (Df, device-resident solver)
  bool jac_built = false;
  std::vector<ca::Term> qterms;  // folded Q = B^-1 N (ordered j,k), kept on the host for the lazy builds
  ca::SmallKernel small;         // model-specialised ensemble kernel for small m, compiled on first use
  bool small_tried = false;
  std::vector<double> hL1, hL2;  // host copies (column-major) for re-targeting R
  // raw (unfolded) operators for the time stepper cnab2, which needs N(x) and B separately
  std::vector<double> hB, hA1, hA2;   // column-major host copies
  std::vector<ca::Term> ncoo;         // N as given
  ca::DevicePack nraw;                // symmetric pack of N, built on first use
  ca::StreamPack nraw_stream;
  bool nraw_built = false, nraw_stream_built = false;
  ca::DevicePack cn_lin;              // dense [M1 M2] acting on [x; g] (2m inputs), per (R, dt)
  ca::StreamPack cn_lin_stream;
  double cn_R = 0.0, cn_dt = 0.0;
  bool cn_built = false;
  ca::DeviceBuffer<double> cn_buf[2], cn_n[2];
  ca::DevicePack obs;                 // two output rows: shear and dissipation (ODEModels.jl:65-137)
  ca::StreamPack obs_stream;
  bool obs_set = false;
  ca::StreamPack stream;         // few-state streaming form of Q, built on first use
  bool stream_built = false;
  ca::DeviceBuffer<double> L1r, L2r;  // row-major copies of L1, L2 for the streaming kernel's dense part
  ca::StreamPack& streaming() {
    if (!stream_built) {
      stream.build(qterms, (int32_t)m, true);
      stream_built = true;
    }
    return stream;
  }
  // f (mode QUAD) or Df action (mode JVP) for an ensemble on the device: DMMA kernel for ensembles,
  // streaming kernel for a few states or when the state tile does not fit
  void eval_dev(int mode, const double* dX, const double* dV, int64_t nstates, int64_t ldx, double R,
                double* dOUT, int64_t ldo, cudaStream_t st);
  ca::JacobianPack& jacobian() {
    if (!jac_built) {
      jac.build(qterms, (int32_t)m);
      jac_built = true;
    }
    return jac;
  }
  ca::DeviceBuffer<double> L1, L2;            // dense m x m column-major (for Df)
  ca::DeviceBuffer<int32_t> q_rowptr;         // CSR of the symmetrised folded Q (a <= b), by output row
  ca::DeviceBuffer<uint32_t> q_ab;            // a | b << 16
  ca::DeviceBuffer<double> q_val;
  ca::DeviceBuffer<int64_t> lin_pos;          // linear slots inside op.vals
  ca::DeviceBuffer<double> lin_l1, lin_l2;
  double cur_invR = std::nan("");
  cudaStream_t streams[ca::kStage] = {nullptr, nullptr, nullptr};
  ca::DeviceBuffer<double> stage_in[ca::kStage], stage_out[ca::kStage], small_x, small_out, jac_out;
  ~ca_model() {
    for (auto& s : streams)
      if (s) cudaStreamDestroy(s);
  }
  void bind() const { CA_CUDA(cudaSetDevice(device)); }
  void need_streams() {
    for (auto& s : streams)
      if (!s) CA_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  }
  // The packed values are shared by all streams: re-targeting R is ordered after everything queued.
  void set_R(double R, cudaStream_t stream);
};

