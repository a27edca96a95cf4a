// Dense Ryser in double-double arithmetic (~106 bits): what the revised front-end's -q flag
// (flags.calculation_quad, revised_perman/flags.h:61-64, main.cpp:1298-1325: "quad" calculation precision)
// becomes on a GPU that has no FP128.  X, the running products and the accumulator are unevaluated sums of
// two doubles; every operation is an error-free transformation (Knuth two-sum, FMA two-product) followed by
// a renormalisation, so a term carries ~1e-31 instead of 1e-16 of relative error and the signed sum over
// 2^(n-1) terms that cancel to a permanent many orders of magnitude below them keeps its digits (chesapeake,
// 39 x 39: FP64 Ryser is 2e-6 off).
//
// Cost per Gray index: N x (X update 8 + product 7) + accumulate 11 = 15 N + 11 FP64 instructions against
// 2 N for the FP64 kernel (7.7 x at N = 36); X lives in shared memory as in ryser_smem_kernel (two words per
// row and thread).  A precision mode, not the headline path.
#pragma once
#include <cuda_runtime.h>

namespace spb {

struct dd_t { double h, l; };

__device__ __forceinline__ dd_t dd_quick(double s, double e) {       // |s| >= |e|
  const double h = __dadd_rn(s, e);
  return dd_t{h, __dadd_rn(e, -__dadd_rn(h, -s))};
}
__device__ __forceinline__ dd_t dd_add_d(dd_t a, double b) {         // a + b
  const double s = __dadd_rn(a.h, b);
  const double bb = __dadd_rn(s, -a.h);
  const double e = __dadd_rn(__dadd_rn(a.h, -__dadd_rn(s, -bb)), __dadd_rn(b, -bb));
  return dd_quick(s, __dadd_rn(e, a.l));
}
__device__ __forceinline__ dd_t dd_add(dd_t a, dd_t b) {
  const double s = __dadd_rn(a.h, b.h);
  const double bb = __dadd_rn(s, -a.h);
  const double e = __dadd_rn(__dadd_rn(a.h, -__dadd_rn(s, -bb)), __dadd_rn(b.h, -bb));
  return dd_quick(s, __dadd_rn(e, __dadd_rn(a.l, b.l)));
}
__device__ __forceinline__ dd_t dd_mul(dd_t a, dd_t b) {
  const double p = __dmul_rn(a.h, b.h);
  double e = __fma_rn(a.h, b.h, -p);
  e = __fma_rn(a.h, b.l, e);
  e = __fma_rn(a.l, b.h, e);
  return dd_quick(p, e);
}

#define DDK_THREADS 128

// Same contract as ryser_smem_kernel: thread t of the grid owns [lo + t*per_thread, ...) n [lo, hi), initialises X
// explicitly at its start (gpu_exact_dense.cu:363-371).  xbase_lo[j] is the low word of the NW start vector
// (computed on the host in long double).  partials[b] / partials[gridDim.x + b] receive the block's sum (high
// and low word).
__global__ void __launch_bounds__(DDK_THREADS)
ryser_dd_kernel(const double* __restrict__ mat_t, const double* __restrict__ xbase, const double* __restrict__ xbase_lo,
                int n, unsigned long long lo, unsigned long long hi, unsigned long long per_thread,
                double* __restrict__ partials) {
  extern __shared__ __align__(16) double dsm[];
  double* colT = dsm;                                  // colT[k*n + j] = A[j][k]
  double* Xh = dsm + n * n + threadIdx.x;              // Xh[j*T], Xl[j*T]
  double* Xl = Xh + (size_t)n * DDK_THREADS;
  __shared__ double wh[DDK_THREADS / 32], wl[DDK_THREADS / 32];
  for (int e = threadIdx.x; e < n * n; e += DDK_THREADS) colT[e] = mat_t[e];
  for (int j = 0; j < n; ++j) { Xh[j * DDK_THREADS] = xbase[j]; Xl[j * DDK_THREADS] = xbase_lo[j]; }
  __syncthreads();

  const unsigned long long gid = (unsigned long long)blockIdx.x * DDK_THREADS + threadIdx.x;
  unsigned long long i = lo + gid * per_thread;
  unsigned long long end = i + per_thread;
  if (end > hi) end = hi;
  dd_t acc{0.0, 0.0};
  if (i < end) {
    unsigned long long g = 0;
    if (i == 0) {
      dd_t p{1.0, 0.0};
      for (int j = 0; j < n; ++j) p = dd_mul(p, dd_t{Xh[j * DDK_THREADS], Xl[j * DDK_THREADS]});
      acc = p;                                          // NW base term, index 0
      i = 1;
    } else {
      g = (i - 1) ^ ((i - 1) >> 1);
      for (int k = 0; k < n - 1; ++k) {
        if ((g >> k) & 1ull) {
          const double* col = colT + k * n;
          for (int j = 0; j < n; ++j) {
            const dd_t x = dd_add_d(dd_t{Xh[j * DDK_THREADS], Xl[j * DDK_THREADS]}, col[j]);
            Xh[j * DDK_THREADS] = x.h; Xl[j * DDK_THREADS] = x.l;
          }
        }
      }
    }
    for (; i < end; ++i) {
      const int k = __ffsll((long long)i) - 1;
      g ^= (1ull << k);
      const bool add = (g >> k) & 1ull;
      const double* col = colT + k * n;
      dd_t p0{1.0, 0.0}, p1{1.0, 0.0};
      int j = 0;
      for (; j + 1 < n; j += 2) {
        const double c0 = add ? col[j] : -col[j], c1 = add ? col[j + 1] : -col[j + 1];
        const dd_t x0 = dd_add_d(dd_t{Xh[j * DDK_THREADS], Xl[j * DDK_THREADS]}, c0);
        const dd_t x1 = dd_add_d(dd_t{Xh[(j + 1) * DDK_THREADS], Xl[(j + 1) * DDK_THREADS]}, c1);
        Xh[j * DDK_THREADS] = x0.h; Xl[j * DDK_THREADS] = x0.l;
        Xh[(j + 1) * DDK_THREADS] = x1.h; Xl[(j + 1) * DDK_THREADS] = x1.l;
        p0 = dd_mul(p0, x0); p1 = dd_mul(p1, x1);
      }
      if (j < n) {
        const dd_t x0 = dd_add_d(dd_t{Xh[j * DDK_THREADS], Xl[j * DDK_THREADS]}, add ? col[j] : -col[j]);
        Xh[j * DDK_THREADS] = x0.h; Xl[j * DDK_THREADS] = x0.l;
        p0 = dd_mul(p0, x0);
      }
      dd_t prod = dd_mul(p0, p1);
      if (i & 1ull) { prod.h = -prod.h; prod.l = -prod.l; }
      acc = dd_add(acc, prod);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const dd_t other{__shfl_down_sync(0xffffffffu, acc.h, o), __shfl_down_sync(0xffffffffu, acc.l, o)};
    acc = dd_add(acc, other);
  }
  if ((threadIdx.x & 31) == 0) { wh[threadIdx.x >> 5] = acc.h; wl[threadIdx.x >> 5] = acc.l; }
  __syncthreads();
  if (threadIdx.x == 0) {
    dd_t v{0.0, 0.0};
    for (int w = 0; w < DDK_THREADS / 32; ++w) v = dd_add(v, dd_t{wh[w], wl[w]});
    partials[blockIdx.x] = v.h;
    partials[gridDim.x + blockIdx.x] = v.l;
  }
}

}  // namespace spb
