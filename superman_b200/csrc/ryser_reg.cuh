// Dense Nijenhuis-Wilf / Ryser Gray-code kernel for sm_100a, X vector in registers.
//
// Replaces kernel_xshared_coalescing_mshared (reference gpu_exact_dense.cu:329-399), which
// keeps X as float in shared memory, finds the flipped column with __ffsll every step and
// walks "for every Gray index: for every row".  Here every thread owns one tile of 2^c
// consecutive Gray indices (aligned to 2^c) and the loop nest is turned inside out:
//
//   for each aligned block of 2^B indices of the tile:            (runtime loop)
//     for each row j:                                             (unrolled, X[j] in a register)
//       v = X[j] +/- A[j][k_hi]                                   (the one "high" column flipped
//                                                                  at the block start, k_hi >= B,
//                                                                  identical for the whole grid:
//                                                                  shared-memory broadcast)
//       P[0] *= v
//       for u = 1 .. 2^B-1:  v +/-= A[j][ctz(u)];  P[u] *= v      (columns < B, direction known
//                                                                  at compile time)
//       X[j] = v
//     acc += P[0] - P[1] + P[2] - ...
//
// so the B most frequently flipped columns of a row are loaded once per block (2 LDS.128 per
// row for B = 4) instead of once per Gray index, every X[j] is consumed the moment it is
// produced (no second copy of X alive while a product is still pending), and the 2^B running
// products P[u] are independent DMUL chains that cover the FP64 latency inside one warp.
// Signed terms are accumulated per thread and closed with a warp-shuffle + block reduction
// into one double per block (no per-thread partial array, no host sum over 2^18 doubles).
//
// FP64 instructions per Gray index: N DADD/DFMA (x update) + (N-1) DMUL + 1 DADD = 2N.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace spb {

__device__ __forceinline__ void lds_f64x2(uint32_t addr, double& a, double& b) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}
__device__ __forceinline__ void lds_f64(uint32_t addr, double& a) {
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a) : "r"(addr));
}

__host__ __device__ constexpr int ctz_c(int u) { int k = 0; while (((u >> k) & 1) == 0) ++k; return k; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// Shared-memory image used by the register kernel (doubles):
//   colT[k * NP + j] = A[j][k]   k in [0, N-1), NP = N rounded up to even   (column-major)
//   lowR[j * LB + k] = A[j][k]   k in [0, B),   LB = B rounded up to even   (row-major, low columns)
template <int N, int B>
struct RegLayout {
  static constexpr int NP = N + (N & 1);
  static constexpr int LB = B + (B & 1);
  static constexpr int COLT = 0;
  static constexpr int LOWR = COLT + N * NP;
  static constexpr int TOTAL = LOWR + N * LB;
};

// Work layout: a "group" is SPB-THREADS (=128) consecutive tiles of 2^c indices, aligned to
// 128 tiles; thread t of a block owns tile t of each of the block's `groups_per_block` consecutive
// groups.  Inside a group the Gray codes of the 128 tile starts agree in every bit >= c+7, so the
// block computes that common part of X once (X_blk, shared memory) and a thread only adds its own
// 8 columns c-1 .. c+6 (masked by its Gray bits): the explicit X start (cf. gpu_exact_dense.cu:
// 363-371) costs 8n FMAs per tile instead of (n-c)n.  That makes SHORT tiles affordable, and short
// tiles are what keeps the result accurate: X is updated in place 2^c times per tile and its
// rounding drift grows with the length of that chain (measured on double/30_0.20_0: 3e-11
// relative error at c=13, 1e-12 at c=8; the reference's chains are 2^13..2^17 long).
//
// partials[blockIdx.x] receives the block's signed sum.  Index 0 (the NW base term that the
// reference adds on the host, gpu_exact_dense.cu:653,691) is an ordinary tile start here, so a
// launch over [0, 2^(n-1)) yields the complete sum.
// Requires B + 1 <= c, c + 7 <= N - 1, group_first * 128 tiles aligned (it is a group index).
template <int N, int B, int THREADS, int MINBLOCKS>
__global__ void __launch_bounds__(THREADS, MINBLOCKS)
ryser_reg_kernel(const double* __restrict__ mat_t,   // mat_t[k*N + j] = A[j][k]
                 const double* __restrict__ xbase,   // NW start vector (gpu_exact_dense.cu:647-654)
                 double* __restrict__ partials, unsigned long long group_first,
                 unsigned long long n_groups, int groups_per_block, int c) {
  static_assert(THREADS == 128, "a group is 128 tiles: 7 thread-specific Gray bits");
  using L = RegLayout<N, B>;
  constexpr int NP = L::NP, LB = L::LB, NB = 1 << B;
  __shared__ __align__(16) double sm[L::TOTAL];
  __shared__ __align__(16) double x_blk[NP];
  __shared__ double warp_part[THREADS / 32];
  // staging: consecutive threads write consecutive shared-memory words in both images (no bank
  // conflicts); the low-column image is gathered from global memory instead
  for (int e = threadIdx.x; e < N * N; e += THREADS) sm[L::COLT + (e / N) * NP + (e % N)] = mat_t[e];
  for (int e = threadIdx.x; e < N * LB; e += THREADS) {
    const int j = e / LB, k = e % LB;
    sm[L::LOWR + e] = (k < B) ? mat_t[k * N + j] : 0.0;
  }
  const uint32_t sm_colT = (uint32_t)__cvta_generic_to_shared(sm + L::COLT);
  const uint32_t sm_lowR = (uint32_t)__cvta_generic_to_shared(sm + L::LOWR);

  double acc = 0.0;
  const unsigned long long g0 = (unsigned long long)blockIdx.x * (unsigned)groups_per_block;
#pragma unroll 1
  for (int gi = 0; gi < groups_per_block; ++gi) {
    const unsigned long long grp = g0 + gi;
    if (grp >= n_groups) break;                                            // block-uniform
    const unsigned long long tile0 = (group_first + grp) * THREADS;        // first tile of the group
    __syncthreads();                                                       // x_blk free / sm staged
    if (threadIdx.x < N) {
      // common part: Gray bits >= c+7 of the group's tile starts
      const unsigned long long s0 = tile0 << c;
      const unsigned long long gc = s0 ^ (s0 >> 1);
      double v = xbase[threadIdx.x];
      for (int k = c + 7; k < N - 1; ++k)
        if ((gc >> k) & 1ull) v += sm[L::COLT + k * NP + threadIdx.x];
      x_blk[threadIdx.x] = v;
    }
    __syncthreads();
    const unsigned long long s = (tile0 + threadIdx.x) << c;               // first index of my tile
    const unsigned long long g = s ^ (s >> 1);                             // Gray code at the tile start
    double x[N];
#pragma unroll
    for (int j = 0; j < N; ++j) x[j] = x_blk[j];
    // my own bits c-1 .. c+6: every lane reads the same column (broadcast) and masks it
#pragma unroll 1
    for (int k = c - 1; k < c + 7; ++k) {
      const double f = (double)((g >> k) & 1ull);
      const double* col = sm + L::COLT + k * NP;
#pragma unroll
      for (int j = 0; j < N; ++j) x[j] = fma(f, col[j], x[j]);
    }

    const int nblk = 1 << (c - B);
    const int tile_odd = threadIdx.x & 1;        // bit c of the tile's first index (tile0 is even)
#pragma unroll 1
    for (int blk = 0; blk < nblk; ++blk) {
      // high column flipped at the block start i0 = s + blk*2^B: k = ctz(i0) = B + ctz(blk), the
      // same for all threads.  Gray bit k after the flip is 1 ^ bit(k+1) of i0 -> add (+1) when
      // that bit is clear; bit k+1 of i0 is a bit of blk, or the tile's parity when k+1 == c.
      // blk == 0 is the tile start, where X is already explicit: weight 0 leaves it unchanged.
      const int k = (blk != 0) ? (B + __ffs(blk) - 1) : B;
      const int up = (k + 1 < c) ? ((blk >> (k + 1 - B)) & 1) : tile_odd;
      const double sg = (blk != 0) ? (up ? -1.0 : 1.0) : 0.0;
      // column B-1 flips in the middle of the block; its direction is bit B of i0
      const double sg_top = (blk & 1) ? -1.0 : 1.0;
      const uint32_t hi_addr = sm_colT + (uint32_t)(k * NP * 8);

      double P[NB];
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double m[LB];
#pragma unroll
        for (int q = 0; q < LB; q += 2) lds_f64x2(sm_lowR + (uint32_t)((j * LB + q) * 8), m[q], m[q + 1]);
        double d;
        lds_f64(hi_addr + (uint32_t)(j * 8), d);
        double v = fma(sg, d, x[j]);
        P[0] = (j == 0) ? v : P[0] * v;
#pragma unroll
        for (int u = 1; u < NB; ++u) {
          const int K = ctz_c(u);
          if (K == B - 1) {
            v = fma(sg_top, m[K], v);
          } else if (((u >> (K + 1)) & 1) == 0) {
            v += m[K];
          } else {
            v -= m[K];
          }
          P[u] = (j == 0) ? v : P[u] * v;
        }
        x[j] = v;
      }
      // term sign (-1)^i: block start is even
      double blk_sum = 0.0;
#pragma unroll
      for (int u = 0; u < NB; u += 2) blk_sum += (P[u] - P[u + 1]);
      acc += blk_sum;
    }
  }

  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) v += warp_part[w];
    partials[blockIdx.x] = v;
  }
}

}  // namespace spb
