// Dense Ryser in double-double arithmetic (~106 bits): what the revised front-end's -q flag
// (flags.calculation_quad, revised_perman/flags.h:61-64, main.cpp:1298-1325: "quad" calculation precision)
// becomes on a GPU that has no FP128.  X, the running products and the accumulator are unevaluated sums of
// two doubles; every operation is an error-free transformation (Knuth two-sum, FMA two-product) followed by
// a renormalisation, so a term carries ~1e-31 instead of 1e-16 of relative error and the signed sum over
// 2^(n-1) terms that cancel to a permanent many orders of magnitude below them keeps its digits (chesapeake,
// 39 x 39: FP64 Ryser is 2e-6 off).
//
// Cost per Gray index: N x (X update 8 + product 7) + accumulate 11 + 14 = 15 N + 25 FP64 instructions against
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
      // four independent product chains (a double-double multiply is six dependent FP64 instructions deep)
      dd_t p[4] = {{1.0, 0.0}, {1.0, 0.0}, {1.0, 0.0}, {1.0, 0.0}};
      int j = 0;
      for (; j + 3 < n; j += 4) {
        dd_t x[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const double c = add ? col[j + t] : -col[j + t];
          x[t] = dd_add_d(dd_t{Xh[(j + t) * DDK_THREADS], Xl[(j + t) * DDK_THREADS]}, c);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          Xh[(j + t) * DDK_THREADS] = x[t].h; Xl[(j + t) * DDK_THREADS] = x[t].l;
          p[t] = dd_mul(p[t], x[t]);
        }
      }
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        if (j + t < n) {
          const dd_t x0 = dd_add_d(dd_t{Xh[(j + t) * DDK_THREADS], Xl[(j + t) * DDK_THREADS]}, add ? col[j + t] : -col[j + t]);
          Xh[(j + t) * DDK_THREADS] = x0.h; Xl[(j + t) * DDK_THREADS] = x0.l;
          p[t] = dd_mul(p[t], x0);
        }
      }
      dd_t prod = dd_mul(dd_mul(p[0], p[1]), dd_mul(p[2], p[3]));
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

// The same sum with the block structure of ryser_reg.cuh: a thread walks its range in aligned blocks of 16 Gray
// indices, ROWS OUTSIDE, STEPS INSIDE -- a row's X pair is read from shared memory once per block, taken through the
// block's 15 low-column flips in registers and multiplied into 16 running double-double products (16 independent
// chains: a double-double multiply is six dependent FP64 instructions deep and the loop kernel above has only four
// chains to hide that with), then moved on by the high column that opens the next block and written back once.
// Requirements: lo and per_thread multiples of 16, hi - lo a multiple of 16, n >= 6 (the host sends the ragged
// ends of a range through ryser_dd_kernel).
__global__ void __launch_bounds__(DDK_THREADS, 3)      // shared memory (2n doubles per thread) admits three blocks at n = 30 ... 44
ryser_dd_blk_kernel(const double* __restrict__ mat_t, const double* __restrict__ xbase, const double* __restrict__ xbase_lo,
                    int n, unsigned long long lo, unsigned long long hi, unsigned long long per_thread,
                    double* __restrict__ partials) {
  extern __shared__ __align__(16) double dsm[];
  double* colT = dsm;                                  // colT[k*n + j] = A[j][k]
  double* Xh = dsm + n * n + threadIdx.x;              // Xh[j*T], Xl[j*T]
  double* Xl = Xh + (size_t)n * DDK_THREADS;
  __shared__ double wh[DDK_THREADS / 32], wl[DDK_THREADS / 32];
  for (int e = threadIdx.x; e < n * n; e += DDK_THREADS) colT[e] = mat_t[e];
  __syncthreads();

  const unsigned long long gid = (unsigned long long)blockIdx.x * DDK_THREADS + threadIdx.x;
  unsigned long long s = lo + gid * per_thread;
  unsigned long long end = s + per_thread;
  if (end > hi) end = hi;
  dd_t acc{0.0, 0.0};
  if (s < end) {
    // explicit X at the first index of the range (gpu_exact_dense.cu:363-371)
    const unsigned long long g0 = s ^ (s >> 1);
    for (int j = 0; j < n; ++j) {
      dd_t x{xbase[j], xbase_lo[j]};
      for (int k = 0; k < n - 1; ++k)
        if ((g0 >> k) & 1ull) x = dd_add_d(x, colT[k * n + j]);
      Xh[j * DDK_THREADS] = x.h; Xl[j * DDK_THREADS] = x.l;
    }
    for (; s < end; s += 16) {
      // inside the block column q < 3 is added when bit q+1 of u is clear; column 3 flips at u = 8 towards
      // bit 4 of the index; the next block opens with column k = ctz(s + 16), added when bit k+1 of s + 16 is clear
      const double s3 = ((s >> 4) & 1ull) ? -1.0 : 1.0;
      const unsigned long long sn = s + 16;
      const int k = __ffsll((long long)sn) - 1;
      const double sk = ((sn >> (k + 1)) & 1ull) ? -1.0 : 1.0;
      const double* colk = colT + (k < n - 1 ? k : 0) * n;       // (k = n-1 only after the very last block: unused)
      dd_t P[16];
      // two rows per trip: the walk of a row is one chain of 16 dependent double-double additions, and two of them
      // side by side (plus the 16 independent products) keep the FP64 pipe fed
      int j = 0;
      for (; j + 1 < n; j += 2) {
        dd_t x{Xh[j * DDK_THREADS], Xl[j * DDK_THREADS]};
        dd_t y{Xh[(j + 1) * DDK_THREADS], Xl[(j + 1) * DDK_THREADS]};
        const double m0 = colT[j], m1 = colT[n + j], m2 = colT[2 * n + j], m3 = s3 * colT[3 * n + j];
        const double q0 = colT[j + 1], q1 = colT[n + j + 1], q2 = colT[2 * n + j + 1], q3 = s3 * colT[3 * n + j + 1];
        const double ck = sk * colk[j], dk = sk * colk[j + 1];
        if (j == 0) P[0] = dd_mul(x, y); else P[0] = dd_mul(dd_mul(P[0], x), y);
#pragma unroll
        for (int u = 1; u < 16; ++u) {
          const int K = (u & 1) ? 0 : (u & 2) ? 1 : (u & 4) ? 2 : 3;
          const double m = K == 0 ? m0 : K == 1 ? m1 : K == 2 ? m2 : m3;
          const double q = K == 0 ? q0 : K == 1 ? q1 : K == 2 ? q2 : q3;
          const bool add = (K == 3 || ((u >> (K + 1)) & 1) == 0);
          x = dd_add_d(x, add ? m : -m);
          y = dd_add_d(y, add ? q : -q);
          if (j == 0) P[u] = dd_mul(x, y); else P[u] = dd_mul(dd_mul(P[u], x), y);
        }
        x = dd_add_d(x, ck);
        y = dd_add_d(y, dk);
        Xh[j * DDK_THREADS] = x.h; Xl[j * DDK_THREADS] = x.l;
        Xh[(j + 1) * DDK_THREADS] = y.h; Xl[(j + 1) * DDK_THREADS] = y.l;
      }
      if (j < n) {                                                 // odd n: the last row on its own (n >= 6: P is set)
        dd_t x{Xh[j * DDK_THREADS], Xl[j * DDK_THREADS]};
        const double m0 = colT[j], m1 = colT[n + j], m2 = colT[2 * n + j], m3 = s3 * colT[3 * n + j];
        const double ck = sk * colk[j];
        P[0] = dd_mul(P[0], x);
#pragma unroll
        for (int u = 1; u < 16; ++u) {
          const int K = (u & 1) ? 0 : (u & 2) ? 1 : (u & 4) ? 2 : 3;
          const double m = K == 0 ? m0 : K == 1 ? m1 : K == 2 ? m2 : m3;
          x = dd_add_d(x, (K == 3 || ((u >> (K + 1)) & 1) == 0) ? m : -m);
          P[u] = dd_mul(P[u], x);
        }
        x = dd_add_d(x, ck);
        Xh[j * DDK_THREADS] = x.h; Xl[j * DDK_THREADS] = x.l;
      }
#pragma unroll
      for (int u = 0; u < 16; u += 2) {                           // s is even: index s + u carries the sign (-1)^u
        acc = dd_add(acc, P[u]);
        acc = dd_add(acc, dd_t{-P[u + 1].h, -P[u + 1].l});
      }
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
