// Dense Nijenhuis-Wilf / Ryser Gray-code kernel for sm_100a, X vector in registers.
//
// Replaces kernel_xshared_coalescing_mshared (reference gpu_exact_dense.cu:329-399), which
// keeps X as float in shared memory, finds the flipped column with __ffsll every step and
// walks "for every Gray index: for every row".  Here every thread owns one tile of 2^c
// consecutive Gray indices (aligned to 2^c) and the loop nest is turned inside out:
//
//   for each aligned block of 2^B indices of the tile:            (runtime loop)
//     for each row j:                                             (unrolled, X[j] in a register)
//       v = X[j] + (+/-A[j][k_hi])                                (the one "high" column flipped
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
//
// Register budget (the kernel runs 4 blocks of 128 threads per SM = 128 registers per thread up
// to n = 40): X takes 2N registers and P 2^(B+1); everything else in the block loop has to fit in
// what is left, so nothing else is a double that lives across the row loop:
//   * the direction of the high-column update is not a +/-1.0 factor but a choice between two
//     shared-memory images of the matrix (A and -A; a block of zeros for the tile's first block,
//     where X is already explicit), i.e. one address register and a plain DADD;
//   * the direction of column B-1 (flipped in the middle of every block) selects one of two
//     low-column images the same way;
//   * the running sum of the tile lives in shared memory (one LDS/STS pair per 2^B indices).
//
// Work distribution: a persistent grid (resident blocks only) pulls "groups" of 128 tiles from an
// atomic counter, so the launch has no wave-quantisation tail; every group writes its own partial
// sum, so the result does not depend on which block took which group (bit-reproducible).
//
// FP64 instructions per Gray index: N DADD (x update) + (N-1) DMUL + 1 DADD = 2N.
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
//   POS [k * NP + j] =  A[j][k]   k in [0, N-1), NP = N rounded up to even   (column-major)
//   NEG [k * NP + j] = -A[j][k]
//   ZERO[j]          =  0
//   LOW0[j * LB + k] =  A[j][k]   k in [0, B),   LB = B rounded up to even    (row-major, low columns)
//   LOW1             =  LOW0 with column B-1 negated
// The images are static shared memory while they fit its 48 KiB (N <= 50: addresses are compile-time
// constants); larger orders take them from opted-in dynamic shared memory (DYN; 70 KiB at N = 64).
// image of the sparse hot/cold kernel (sparse_reg.cuh): colT[k * NP + j] = A[j][k], lowR[j * LB + k]
template <int N, int B>
struct RegLayout {
  static constexpr int NP = N + (N & 1);
  static constexpr int LB = B + (B & 1);
  static constexpr int COLT = 0;
  static constexpr int LOWR = COLT + N * NP;
  static constexpr int TOTAL = LOWR + N * LB;
};

template <int N, int B>
struct DenseLayout {
  // column pitch: N rounded up to even (16-byte loads).  Neither the pitch nor the distance between the
  // two low-column images may be a power of two: ptxas then computes the block loop's (uniform)
  // addresses with shifts in vector registers, and LDS [R + UR + imm] instead of LDS [UR + imm]
  // measured 3 % slower on the whole kernel (n = 32: 7.88 -> 7.68 ms, n = 64: 0.952 -> 0.973 of peak)
  static constexpr int NE = N + (N & 1);
  static constexpr int NP = ((NE & (NE - 1)) == 0) ? NE + 2 : NE;
  static constexpr int LB = B + (B & 1);
  static constexpr bool DYN = (N > 50);
  static constexpr int POS = 0;
  static constexpr int NEG = POS + N * NP;
  static constexpr int ZERO = NEG + N * NP;
  static constexpr int LOW0 = ZERO + NP;
  static constexpr int LOWSZ = ((N * LB) & (N * LB - 1)) == 0 ? N * LB + 2 : N * LB;   // same reason
  static constexpr int LOW1 = LOW0 + LOWSZ;
  static constexpr int TOTAL = LOW1 + N * LB;
};

// Work layout: a "group" is SPB-THREADS (=128) consecutive tiles of 2^c indices, aligned to
// 128 tiles; thread t of a block owns tile t of every group the block pulls from the queue.
// Inside a group the Gray codes of the 128 tile starts agree in every bit >= c+7, so the
// block computes that common part of X once (X_blk, shared memory) and a thread only adds its own
// 8 columns c-1 .. c+6 (masked by its Gray bits): the explicit X start (cf. gpu_exact_dense.cu:
// 363-371) costs 8n FMAs per tile instead of (n-c)n.  That makes SHORT tiles affordable, and short
// tiles are what keeps the result accurate: X is updated in place 2^c times per tile and its
// rounding drift grows with the length of that chain (measured on double/30_0.20_0: 3e-11
// relative error at c=13, 1e-12 at c=8; the reference's chains are 2^13..2^17 long).
//
// partials[g] receives the signed sum of group group_first + g, g in [0, n_groups).  Index 0 (the NW
// base term that the reference adds on the host, gpu_exact_dense.cu:653,691) is an ordinary tile
// start here, so a launch over [0, 2^(n-1)) yields the complete sum.
// queue[0] is the next group to hand out and queue[1] the number of blocks that have left the loop;
// both are zero between launches (the last block to leave resets them).
// Requires B + 1 <= c, c + 7 <= N - 1.
template <int N, int B, int THREADS, int MINBLOCKS>
__global__ void __launch_bounds__(THREADS, MINBLOCKS)
ryser_reg_kernel(const double* __restrict__ mat_t,   // mat_t[k*N + j] = A[j][k]
                 const double* __restrict__ xbase,   // NW start vector (gpu_exact_dense.cu:647-654)
                 double* __restrict__ partials, unsigned long long group_first,
                 unsigned int n_groups, unsigned int* __restrict__ queue, int c) {
  static_assert(THREADS == 128, "a group is 128 tiles: 7 thread-specific Gray bits");
  using L = DenseLayout<N, B>;
  constexpr int NP = L::NP, LB = L::LB, NB = 1 << B;
  extern __shared__ __align__(16) double sm_dyn[];
  __shared__ __align__(16) double sm_static[L::DYN ? 2 : L::TOTAL];
  double* const sm = L::DYN ? sm_dyn : sm_static;
  __shared__ __align__(16) double x_blk[NP];
  __shared__ double acc_sm[THREADS];
  __shared__ double warp_part[THREADS / 32];
  __shared__ unsigned int next_grp[2];
  // staging: consecutive threads write consecutive shared-memory words in every image (no bank
  // conflicts); the low-column images are gathered from global memory instead
  for (int e = threadIdx.x; e < N * N; e += THREADS) {
    const double a = mat_t[e];
    sm[L::POS + (e / N) * NP + (e % N)] = a;
    sm[L::NEG + (e / N) * NP + (e % N)] = -a;
  }
  if (threadIdx.x < NP) sm[L::ZERO + threadIdx.x] = 0.0;
  for (int e = threadIdx.x; e < N * LB; e += THREADS) {
    const int j = e / LB, k = e % LB;
    const double a = (k < B) ? mat_t[k * N + j] : 0.0;
    sm[L::LOW0 + e] = a;
    sm[L::LOW1 + e] = (k == B - 1) ? -a : a;
  }
  const uint32_t sm_pos = (uint32_t)__cvta_generic_to_shared(sm + L::POS);
  const uint32_t sm_low = (uint32_t)__cvta_generic_to_shared(sm + L::LOW0);

  if (threadIdx.x == 0) next_grp[0] = atomicAdd(&queue[0], 1u);
  __syncthreads();                                                         // images staged, first group known
  const int nblk = 1 << (c - B);
#pragma unroll 1
  for (int it = 0;; ++it) {
    const unsigned int grp = next_grp[it & 1];
    if (grp >= n_groups) break;                                            // block-uniform
    const unsigned long long tile0 = (group_first + grp) * THREADS;        // first tile of the group
    // the next group is fetched now and published after the X start below: the round trip of the
    // atomic is hidden behind that work instead of holding the whole block at the barrier
    unsigned int nxt = 0;
    if (threadIdx.x == 0) nxt = atomicAdd(&queue[0], 1u);
    if (threadIdx.x < N) {
      // common part: Gray bits >= c+7 of the group's tile starts
      const unsigned long long s0 = tile0 << c;
      const unsigned long long gc = s0 ^ (s0 >> 1);
      double v = xbase[threadIdx.x];
      for (int k = c + 7; k < N - 1; ++k)
        if ((gc >> k) & 1ull) v += sm[L::POS + k * NP + threadIdx.x];
      x_blk[threadIdx.x] = v;
    }
    acc_sm[threadIdx.x] = 0.0;
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
      const double* col = sm + L::POS + k * NP;
#pragma unroll
      for (int j = 0; j < N; ++j) x[j] = fma(f, col[j], x[j]);
    }
    if (threadIdx.x == 0) next_grp[(it & 1) ^ 1] = nxt;                    // read after the next barrier

#pragma unroll 1
    for (int blk = 0; blk < nblk; ++blk) {
      // high column flipped at the block start i0 = s + blk*2^B: k = ctz(i0) = B + ctz(blk), the
      // same for all threads.  Gray bit k after the flip is 1 ^ bit(k+1) of i0 -> add (+1) when
      // that bit is clear; bit k+1 of i0 is a bit of blk, or the tile's parity when k+1 == c.
      // blk == 0 is the tile start, where X is already explicit: nothing is added there.
      // Everything here is block-uniform (addresses live in uniform registers) except that one
      // parity case, the middle block of the tile: it adds column c-1 like an even tile, and odd
      // tiles take the column out twice beforehand (N extra FMAs per tile of 2^c indices).
      const int k = (blk != 0) ? (B + __ffs(blk) - 1) : B;
      const int up = (k + 1 < c) ? ((blk >> (k + 1 - B)) & 1) : 0;
      if (blk == (nblk >> 1)) {
        const double f = -2.0 * (double)(threadIdx.x & 1);   // bit c of the tile's first index (a group starts even)
        const double* col = sm + L::POS + (c - 1) * NP;
#pragma unroll
        for (int j = 0; j < N; ++j) x[j] = fma(f, col[j], x[j]);
      }
      const int img = (blk != 0) ? (up ? (L::NEG - L::POS) + k * NP : k * NP) : (L::ZERO - L::POS);
      const uint32_t hi_addr = sm_pos + (uint32_t)(img * 8);
      // column B-1 flips in the middle of the block; its direction is bit B of i0
      const uint32_t low_addr = sm_low + (uint32_t)((blk & 1) * (L::LOW1 - L::LOW0) * 8);

      double P[NB];
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double m[LB];
#pragma unroll
        for (int q = 0; q < LB; q += 2) lds_f64x2(low_addr + (uint32_t)((j * LB + q) * 8), m[q], m[q + 1]);
        double d;
        lds_f64(hi_addr + (uint32_t)(j * 8), d);
        double v = x[j] + d;
        P[0] = (j == 0) ? v : P[0] * v;
#pragma unroll
        for (int u = 1; u < NB; ++u) {
          const int K = ctz_c(u);
          // column B-1 takes its sign from the image; the others alternate with bit K+1 of u
          if (K == B - 1 || ((u >> (K + 1)) & 1) == 0) v += m[K];
          else v -= m[K];
          P[u] = (j == 0) ? v : P[u] * v;
        }
        x[j] = v;
      }
      // term sign (-1)^i: block start is even
      double blk_sum = 0.0;
#pragma unroll
      for (int u = 0; u < NB; u += 2) blk_sum += (P[u] - P[u + 1]);
      acc_sm[threadIdx.x] += blk_sum;
    }

    const double acc = warp_sum(acc_sm[threadIdx.x]);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = acc;
    __syncthreads();                               // also: everybody is done with x_blk / next_grp[it&1]
    if (threadIdx.x == 0) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < THREADS / 32; ++w) v += warp_part[w];
      partials[next_grp[it & 1]] = v;              // = grp, re-read: one register less across the block loop
    }
  }
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&queue[1], 1u) == gridDim.x - 1) {   // nobody will touch the counters any more
      queue[0] = 0u;
      queue[1] = 0u;
      __threadfence();
    }
  }
}

}  // namespace spb
