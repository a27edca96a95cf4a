// dev: FP64 issue rate per SM sub-partition as a function of resident warps and of the independent chains
// per thread (explains why the level kernel's pipe utilisation follows the share of warps inside their
// FP64-dense phase).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/fp64_issue_probe tools/fp64_issue_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void probe(double* out, double a, double b, int iters) {
  double x[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) x[i] = a + i + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < CH; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;
}
template <int CH>
static void run(int warps_per_smsp, int sms) {
  double* d; cudaMalloc(&d, 8);
  const int threads = 128 * warps_per_smsp > 1024 ? 1024 : 128 * warps_per_smsp;
  const int blocks_per_sm = (128 * warps_per_smsp + threads - 1) / threads;
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<CH><<<sms * blocks_per_sm, threads>>>(d, 1.0000001, 1e-9, 100);
  cudaEventRecord(e0);
  probe<CH><<<sms * blocks_per_sm, threads>>>(d, 1.0000001, 1e-9, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double instr = (double)sms * blocks_per_sm * threads * (double)iters * 8 * CH;   // thread-level
  const double per_clk_sm = instr / (ms * 1e-3) / 1.965e9 / sms;
  printf("chains %d  warps/SMSP %d  %.3f ms  %.1f FP64 lanes/clk/SM (peak 64)  = %.0f%%\n", CH, warps_per_smsp, ms, per_clk_sm, per_clk_sm / 64 * 100);
  cudaFree(d);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  for (int w : {1, 2, 3, 4, 5, 6, 8}) run<1>(w, p.multiProcessorCount);
  for (int w : {1, 2, 3, 4, 5, 6, 8}) run<2>(w, p.multiProcessorCount);
  for (int w : {1, 2, 3, 4, 5, 6, 8}) run<4>(w, p.multiProcessorCount);
  for (int w : {1, 2, 3, 4, 5, 6, 8}) run<8>(w, p.multiProcessorCount);
  return 0;
}
