// ref_shim_gpu.cu -- the reference's own .cu files (compiled unmodified, where they lie under
// /root/reference) behind a C ABI: (1) cpu_perman64, the OpenMP all-double range kernel that the
// reference's hybrid chunk queue runs on host cores (gpu_exact_dense.cu:6-69) -- bench.py times it
// on a slice of the n=36 workload as the CPU baseline; (2) the reference's GPU wrappers, so that
// the kernels-to-beat can be timed on the same B200.
// TEST / BENCH INFRASTRUCTURE ONLY; output goes to oracle/_ref/ (git-ignored).
#include <cstdint>
#include <iostream>
#include <algorithm>
#include "util.h"
#include "gpu_exact_dense.cu"
#include "gpu_exact_sparse.cu"
#include "gpu_approximation_sparse.cu"

extern "C" {

// gpu_exact_dense.cu:6 -- partial sum over Gray indices [start, end), X in double, `threads` OpenMP threads
double ref_cpu_perman64(const double* mat, int nov, long long start, long long end, int threads) {
  double x[64];
  double* mat_t = new double[nov * nov];
  for (int j = 0; j < nov; j++) {
    double rs = .0f;
    for (int k = 0; k < nov; k++) rs += mat[(j * nov) + k];
    x[j] = mat[(j * nov) + (nov - 1)] - rs / 2;
  }
  for (int i = 0; i < nov; i++)
    for (int j = 0; j < nov; j++) mat_t[(i * nov) + j] = mat[(j * nov) + i];
  double r = cpu_perman64(mat_t, x, nov, start, end, threads);
  delete[] mat_t;
  return r;
}

// gpu_exact_dense.cu:701 with gpu_num devices (gpu_num = 1 runs on device 0; the single-GPU
// wrapper :640 hard-codes device 1).  Launch geometry as RunAlgo: 2048 x 128 for double files.
double ref_gpu_dense_multigpu(const double* mat, int nov, int gpu_num, int grid_dim, int block_dim) {
  return gpu_perman64_xshared_coalescing_mshared_multigpu((double*)mat, nov, gpu_num, grid_dim, block_dim);
}
double ref_gpu_dense_chunks(const double* mat, int nov, int gpu_num, int grid_dim, int block_dim) {
  return gpu_perman64_xshared_coalescing_mshared_multigpucpu_chunks((double*)mat, nov, gpu_num, false, 1, grid_dim, block_dim);
}
// gpu_exact_sparse.cu:916
double ref_gpu_sparse_multigpu(const double* mat, const int* cptrs, const int* rows, const double* cvals,
                               int nov, int gpu_num, int grid_dim, int block_dim) {
  return gpu_perman64_xshared_coalescing_mshared_multigpu_sparse((double*)mat, (int*)cptrs, (int*)rows,
                                                                 (double*)cvals, nov, gpu_num, grid_dim, block_dim);
}
// gpu_approximation_sparse.cu:497 / 663 with gpu_num devices: every launch runs 1024 x 1024 trials
// (the reference's own granularity), so number_of_times is effectively rounded up to 2^20 per launch
double ref_gpu_rasmussen_chunks_sparse(const int* cptrs, const int* rows, const int* rptrs, const int* cols,
                                       int nov, int nnz, int number_of_times, int gpu_num) {
  return gpu_perman64_rasmussen_multigpucpu_chunks_sparse((int*)cptrs, (int*)rows, (int*)rptrs, (int*)cols, nov, nnz,
                                                          number_of_times, gpu_num, false, 1, true);
}
double ref_gpu_scaling_chunks_sparse(const int* cptrs, const int* rows, const int* rptrs, const int* cols,
                                     int nov, int nnz, int number_of_times, int gpu_num, int y, int z) {
  return gpu_perman64_approximation_multigpucpu_chunks_sparse((int*)cptrs, (int*)rows, (int*)rptrs, (int*)cols, nov, nnz,
                                                              number_of_times, gpu_num, false, y, z, 1, true);
}
int ref_gpu_max_threads() { return omp_get_max_threads(); }

}  // extern "C"
