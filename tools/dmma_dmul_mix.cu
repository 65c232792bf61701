// How much FP64-pipe time does a DMUL cost when interleaved with DMMA.8x8x4?  (design probe for K2)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NMMA, int NMUL, bool DEP>
__global__ void __launch_bounds__(256) k(double* out, int iters, double a, double b) {
  double c0[8], c1[8], p[4];
  for (int t = 0; t < 8; ++t) { c0[t] = threadIdx.x * 1e-3; c1[t] = t; }
  for (int t = 0; t < 4; ++t) p[t] = 1.0 + t * 1e-9 + threadIdx.x * 1e-12;
  double x = a;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int t = 0; t < NMUL; ++t) asm volatile("mul.f64 %0, %0, %1;" : "+d"(p[t]) : "d"(x));
#pragma unroll
    for (int t = 0; t < NMMA; ++t) dmma(c0[t], c1[t], DEP ? p[t % (NMUL ? NMUL : 1)] : a, b);
  }
  double s = 0;
  for (int t = 0; t < 8; ++t) s += c0[t] + c1[t];
  for (int t = 0; t < 4; ++t) s += p[t];
  if (s == 123.456) out[0] = s;
}
template <int NMMA, int NMUL, bool DEP>
void run(const char* name, double* out) {
  const int iters = 20000, grid = 148 * 2;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NMMA, NMUL, DEP><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<NMMA, NMUL, DEP><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  // cycles per iteration per SM sub-partition: 4 warps per scheduler (2 CTAs x 8 warps / 4)
  double cyc = ms * 1e-3 * 1.965e9 / iters / 4.0;
  printf("%-28s %8.3f ms  %6.1f cycles/iter/warp-slot  (DMMA-only ideal %d)\n", name, ms, cyc, NMMA * 16);
}
int main() {
  double* out; cudaMalloc(&out, 64);
  run<8, 0, false>("8 DMMA", out);
  run<8, 4, false>("8 DMMA + 4 DMUL indep", out);
  run<8, 4, true>("8 DMMA + 4 DMUL feeding", out);
  run<4, 4, false>("4 DMMA + 4 DMUL indep", out);
  run<4, 4, true>("4 DMMA + 4 DMUL feeding", out);
  run<8, 1, false>("8 DMMA + 1 DMUL indep", out);
  run<0, 4, false>("4 DMUL", out);
  return 0;
}
