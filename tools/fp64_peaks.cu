// FP64 denominators for the roofline: register-resident DFMA loop and DMMA (mma.sync f64) loop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peaks tools/fp64_peaks.cu
// Prints one JSON object. MEASURED_PEAKS.json has no FP64 entry, so the
// bench measures its own; this tool is the stand-alone version of that measurement.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)

template <int CHAINS>
__global__ void __launch_bounds__(256) dfma_loop(double* out, int iters, double a, double b) {
  double acc[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) acc[c] = threadIdx.x * 1e-3 + c;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = fma(acc[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += acc[c];
  if (s == 123.456) out[0] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int TILES>
__global__ void __launch_bounds__(256) dmma_loop(double* out, int iters, double a, double b) {
  double c0[TILES], c1[TILES];
#pragma unroll
  for (int t = 0; t < TILES; ++t) { c0[t] = threadIdx.x * 1e-3; c1[t] = t; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int t = 0; t < TILES; ++t) dmma884(c0[t], c1[t], a, b);
  }
  double s = 0;
#pragma unroll
  for (int t = 0; t < TILES; ++t) s += c0[t] + c1[t];
  if (s == 123.456) out[0] = s;
}

// m16n8k8 f64 (sm_90+ shape): A 4 regs, B 2 regs, C 4 regs
template <int TILES>
__global__ void __launch_bounds__(256) dmma1688_loop(double* out, int iters, double a, double b) {
  double c[TILES][4];
#pragma unroll
  for (int t = 0; t < TILES; ++t) { c[t][0] = threadIdx.x * 1e-3; c[t][1] = t; c[t][2] = 1; c[t][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int t = 0; t < TILES; ++t)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+d"(c[t][0]), "+d"(c[t][1]), "+d"(c[t][2]), "+d"(c[t][3])
                   : "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a));
  }
  double s = 0;
#pragma unroll
  for (int t = 0; t < TILES; ++t) s += c[t][0] + c[t][1] + c[t][2] + c[t][3];
  if (s == 123.456) out[0] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch(); CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, 64));
  const int iters = 20000;
  printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
  for (int cps = 1; cps <= 8; cps *= 2) {
    int grid = sms * cps;
    double ms = time_ms([&] { dfma_loop<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
    double tf = 2.0 * 8 * iters * 256.0 * grid / (ms * 1e-3) / 1e12;
    printf(", \"dfma_tflops_%dcta\": %.3f", cps, tf);
  }
  for (int cps = 1; cps <= 8; cps *= 2) {
    int grid = sms * cps;
    double ms = time_ms([&] { dmma_loop<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
    double tf = 2.0 * 256 * 8 * iters * 8.0 * grid / (ms * 1e-3) / 1e12;
    printf(", \"dmma884_tflops_%dcta\": %.3f", cps, tf);
  }
  for (int cps = 1; cps <= 4; cps *= 2) {
    int grid = sms * cps;
    double ms = time_ms([&] { dmma1688_loop<4><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
    double tf = 2.0 * 16 * 8 * 8 * 4 * iters * 8.0 * grid / (ms * 1e-3) / 1e12;
    printf(", \"dmma1688_tflops_%dcta\": %.3f", cps, tf);
  }
  // latency-ish: one tile chain per warp, 1 warp per SM sub-partition
  {
    double ms = time_ms([&] { dmma_loop<1><<<sms, 128>>>(out, iters, 1.0000001, 1e-9); }, 3);
    printf(", \"dmma884_dependent_ns\": %.2f", ms * 1e6 / iters);
    ms = time_ms([&] { dfma_loop<1><<<sms, 128>>>(out, iters, 1.0000001, 1e-9); }, 3);
    printf(", \"dfma_dependent_ns\": %.2f", ms * 1e6 / iters);
  }
  printf("}\n");
  return 0;
}
