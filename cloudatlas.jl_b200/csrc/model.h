// The device side of an ODEModel (shared by api_model.cu and hookstep.cu).
#pragma once
#include <cmath>

#include "bilinear.h"
#include "jit_small.h"

namespace ca {

// Seconds per phase of the model build (ca_model_build_info); on_device = 1 for the device build (model_build.cu).
struct ModelBuildStats {
  double blocks = 0, fold = 0, symmetrise = 0, cluster = 0, tile_unions = 0, layout = 0, place = 0, stream = 0, total = 0;
  int32_t on_device = 0;
  int64_t nclusters = 0, ntiles = 0, nterms = 0;
};

// The folded quadratic operator Q = B^-1 N of a device-built model, block-dense and device-resident: block b holds
// g_b rows x |U_b| ordered pairs (j * m + k, ascending) of values, zeros included.  Expanded into host terms only when
// a lazily built structure needs them (Df pattern, generated small kernel, in-kernel solver CSR).
struct FoldedDevice {
  int64_t total_u = 0;
  std::vector<int64_t> u_off, dense_off;  // [nb+1] offsets into ukeys / R
  std::vector<int32_t> row_off, rows;     // members of the blocks
  DeviceBuffer<uint32_t> ukeys;
  DeviceBuffer<double> R;
  void to_host_terms(std::vector<Term>& out, int32_t m) const;
};

}  // namespace ca

struct ca_model {
  int device = 0;
  int64_t m = 0;
  int64_t nblocks = 0, nnz_lin = 0;
  ca::DevicePack op;           // symmetric pack of Q plus the linear slots
  ca::JacobianPack jac;        // of Q, built on first use (Df, device-resident solver)
  bool jac_built = false;
  std::vector<ca::Term> qterms;  // folded Q = B^-1 N (ordered j,k), kept on the host for the lazy builds
  // device build: the folded operator, N as given and A1, A2 stay on the device; host copies are fetched on first use
  ca::FoldedDevice folded;
  bool qterms_on_device = false, ncoo_on_device = false, linear_on_device = false;
  ca::DeviceBuffer<ca::Term> dncoo;
  ca::DeviceBuffer<double> dA1, dA2;
  ca::ModelBuildStats build_stats;
  const std::vector<ca::Term>& host_qterms() {
    if (qterms_on_device) {
      folded.to_host_terms(qterms, (int32_t)m);
      qterms_on_device = false;
    }
    return qterms;
  }
  const std::vector<ca::Term>& host_ncoo() {
    if (ncoo_on_device) {
      ncoo.resize(dncoo.count);
      if (dncoo.count) CA_CUDA(cudaMemcpy(ncoo.data(), dncoo.ptr, dncoo.count * sizeof(ca::Term), cudaMemcpyDeviceToHost));
      ncoo_on_device = false;
    }
    return ncoo;
  }
  // B (unfolded, column-major) on the host: the device build keeps it on the device until cnab2 asks for it
  ca::DeviceBuffer<double> dB;
  bool B_on_device = false;
  const std::vector<double>& host_B() {
    if (B_on_device) {
      hB.resize((size_t)m * m);
      CA_CUDA(cudaMemcpy(hB.data(), dB.ptr, hB.size() * sizeof(double), cudaMemcpyDeviceToHost));
      B_on_device = false;
    }
    return hB;
  }
  // fills hA1, hA2, hL1, hL2 (column-major)
  void host_linear() {
    if (!linear_on_device) return;
    const size_t mm = (size_t)m * m;
    hA1.resize(mm);
    hA2.resize(mm);
    hL1.resize(mm);
    hL2.resize(mm);
    CA_CUDA(cudaMemcpy(hA1.data(), dA1.ptr, mm * sizeof(double), cudaMemcpyDeviceToHost));
    CA_CUDA(cudaMemcpy(hA2.data(), dA2.ptr, mm * sizeof(double), cudaMemcpyDeviceToHost));
    CA_CUDA(cudaMemcpy(hL1.data(), L1.ptr, mm * sizeof(double), cudaMemcpyDeviceToHost));
    CA_CUDA(cudaMemcpy(hL2.data(), L2.ptr, mm * sizeof(double), cudaMemcpyDeviceToHost));
    linear_on_device = false;
  }
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
  bool cn_built = false, cn_dense = false;
  ca::DeviceBuffer<double> cn_M1, cn_M2;   // M1 = LHS^-1 RHS, M2 = LHS^-1 (column-major), formed on the device per (R, dt)
  ca::DeviceBuffer<double> cn_buf[2], cn_n[2];
  ca::DevicePack obs;                 // two output rows: shear and dissipation (ODEModels.jl:65-137)
  ca::StreamPack obs_stream;
  bool obs_set = false, obs_built = false;
  std::vector<double> obs_shear, obs_D;  // as given to ca_model_set_observables; packed on first use
  ca::StreamPack stream;         // few-state streaming form of Q, built on first use
  bool stream_built = false;
  ca::DeviceBuffer<double> L1r, L2r;  // row-major copies of L1, L2 for the streaming kernel's dense part
  ca::StreamPack& streaming() {
    if (!stream_built) {
      stream.build(host_qterms(), (int32_t)m, true);
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
      jac.build(host_qterms(), (int32_t)m);
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

