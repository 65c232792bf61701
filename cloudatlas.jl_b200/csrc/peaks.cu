// FP64 peak measurement (roofline denominators): DMMA and DFMA register-resident loops.
#include "bilinear_kernels.cuh"
#include "common.h"

using namespace ca;

namespace {

__global__ void __launch_bounds__(256) peak_dmma_kernel(double* out, int iters, double a, double b) {
  double c0[8], c1[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) { c0[t] = threadIdx.x * 1e-3; c1[t] = t; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int t = 0; t < 8; ++t) dmma884(c0[t], c1[t], a, b);
  }
  double s = 0;
#pragma unroll
  for (int t = 0; t < 8; ++t) s += c0[t] + c1[t];
  if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) peak_dfma_kernel(double* out, int iters, double a, double b) {
  double acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = threadIdx.x * 1e-3 + c;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = fma(acc[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) s += acc[c];
  if (s == 123.456) out[0] = s;
}

template <typename L>
double best_ms(L launch) {
  cudaEvent_t e0, e1;
  CA_CUDA(cudaEventCreate(&e0));
  CA_CUDA(cudaEventCreate(&e1));
  launch();
  CA_CUDA(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < 5; ++r) {
    CA_CUDA(cudaEventRecord(e0));
    launch();
    CA_CUDA(cudaEventRecord(e1));
    CA_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    CA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    best = ms < best ? ms : best;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return best;
}

}  // namespace

extern "C" int32_t ca_measure_fp64_peak(double* dmma_tflops, double* dfma_tflops) {
  return guarded([&] {
    require_device();
    int dev = 0, sms = 0;
    CA_CUDA(cudaGetDevice(&dev));
    CA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    DeviceBuffer<double> out;
    out.alloc(8);
    const int iters = 8000, grid = sms * 4;
    if (dmma_tflops) {
      const double ms = best_ms([&] { peak_dmma_kernel<<<grid, 256>>>(out.ptr, iters, 1.0000001, 1e-9); count_launch(); });
      *dmma_tflops = 2.0 * 256 * 8 * iters * 8.0 * grid / (ms * 1e-3) / 1e12;
    }
    if (dfma_tflops) {
      const double ms = best_ms([&] { peak_dfma_kernel<<<grid, 256>>>(out.ptr, iters, 1.0000001, 1e-9); count_launch(); });
      *dfma_tflops = 2.0 * 8 * iters * 256.0 * grid / (ms * 1e-3) / 1e12;
    }
    CA_CUDA(cudaGetLastError());
  });
}
