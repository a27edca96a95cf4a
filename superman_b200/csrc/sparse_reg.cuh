// SpaRyser / SkipPer kernel for sm_100a: X in registers, warp-uniform control flow.
//
// Replaces kernel_xshared_coalescing_mshared_sparse (gpu_exact_sparse.cu:455-552: per step walk
// CCS column k, keep the product incrementally with an FP64 DIVIDE per touched row and a
// data-dependent x==0 branch per row) and kernel_xshared_coalescing_mshared_skipper
// (gpu_exact_sparse.cu:555-670: per-thread skip lengths, so the lanes of a warp run apart).
//
// Same tile / block decomposition as the dense kernel (ryser_reg.cuh): a thread owns a tile of 2^c
// Gray indices and walks it in aligned blocks of 2^B.  Sparsity is used through a ROW order chosen
// on the host: rows are sorted by the lowest column that touches them, so that
//   rows [0, H)      "hot"  : have an entry in one of the B low columns -> full treatment,
//                             2^B running products, as in the dense kernel;
//   rows [H, N)      "cold" : constant inside a block -> ONE update and ONE multiply per block
//                             into a common factor Q (instead of 2^B of each);
//   rows [N-TC, N)   "tile-cold": touched only by columns >= c -> constant over the whole tile.
// A Ryser term is prod_hot * Q, so a block whose Q is exactly zero contributes exactly zero: the
// SkipPer variant skips its hot-row work when EVERY lane of the warp has Q == 0 (one __all_sync,
// no divergence), and before that it drops whole tiles whose tile-cold rows contain a zero: the
// warp tests 32 candidate tiles at a time, ballots, and compacts the survivors into a
// shared-memory queue so that the expensive part always runs with full warps.  Skipped terms are
// exact zeros, so the sum is the one SpaRyser computes (up to the order of additions).
//
// H, TC, c are grid-uniform kernel arguments; every branch on them is a uniform branch.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "ryser_reg.cuh"

namespace spb {

struct SparseArgs {
  const double* mat_t;        // mat_t[k*N + j] = D[j][k], rows already in hot-first order
  const double* xbase;        // NW start vector in the same row order
  double* partials;           // one signed sum per block
  unsigned long long* visited;// one count of evaluated 2^B-blocks per block
  unsigned long long tile_first, n_tiles;
  int c;                      // log2 tile
  int H;                      // hot rows
  int TC;                     // tile-cold rows (SKIP only)
  int tiles_per_warp;         // candidates handed to one warp (multiple of 32)
};

template <int N, int B, int THREADS, int MINBLOCKS, bool SKIP>
__global__ void __launch_bounds__(THREADS, MINBLOCKS)
sparse_reg_kernel(const SparseArgs a) {
  using L = RegLayout<N, B>;
  constexpr int NP = L::NP, LB = L::LB, NB = 1 << B, WARPS = THREADS / 32;
  __shared__ __align__(16) double sm[L::TOTAL];
  __shared__ double warp_part[WARPS];
  __shared__ unsigned long long warp_vis[WARPS];
  __shared__ unsigned long long queue[WARPS][64];
  __shared__ unsigned long long s_cand[WARPS][2];   // the warp's next candidate tile and the end of its range
  __shared__ double s_acc[THREADS];            // per-thread running sum and evaluated-block count: kept out of
  __shared__ unsigned int s_vis[THREADS];      // the register file, which X and the running products fill
  // staging: consecutive threads write consecutive shared-memory words in both images (no bank
  // conflicts); the low-column image is gathered from global memory instead
  for (int e = threadIdx.x; e < N * N; e += THREADS) sm[L::COLT + (e / N) * NP + (e % N)] = a.mat_t[e];
  for (int e = threadIdx.x; e < N * LB; e += THREADS) {
    const int j = e / LB, k = e % LB;
    sm[L::LOWR + e] = (k < B) ? a.mat_t[k * N + j] : 0.0;
  }
  __syncthreads();
  const uint32_t sm_colT = (uint32_t)__cvta_generic_to_shared(sm + L::COLT);
  const uint32_t sm_lowR = (uint32_t)__cvta_generic_to_shared(sm + L::LOWR);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  constexpr int G = 4;
  const int c = a.c;
  const int Hg = ((a.H + G - 1) / G) * G;

  if (lane == 0) {
    const unsigned long long wg = (unsigned long long)blockIdx.x * WARPS + wib;
    const unsigned long long c0 = wg * (unsigned long long)a.tiles_per_warp;
    const unsigned long long c1 = c0 + (unsigned long long)a.tiles_per_warp;
    s_cand[wib][0] = c0;
    s_cand[wib][1] = c1 > a.n_tiles ? a.n_tiles : c1;
  }
  __syncwarp();

  s_acc[threadIdx.x] = 0.0;
  s_vis[threadIdx.x] = 0u;
  int q = 0;   // live tiles waiting in queue[wib]
  for (;;) {
    // ---- refill: test 32 candidate tiles per trip until a full warp of survivors is queued ----
    unsigned long long cand = s_cand[wib][0];
    const unsigned long long cand_hi = s_cand[wib][1];
    while (q < 32 && cand < cand_hi) {
      const unsigned long long t = cand + lane;
      bool alive = t < cand_hi;
      if (SKIP && a.TC > 0) {
        const unsigned long long s = (a.tile_first + (alive ? t : cand)) << c;
        const unsigned long long g = s ^ (s >> 1);
        for (int r = N - a.TC; r < N; ++r) {
          double xr = a.xbase[r];
          for (int k = c; k < N - 1; ++k)
            xr = fma((double)((g >> k) & 1ull), sm[L::COLT + k * NP + r], xr);
          alive = alive && (xr != 0.0);
        }
      }
      const unsigned m = __ballot_sync(0xffffffffu, alive);
      if (alive) queue[wib][q + __popc(m & ((1u << lane) - 1u))] = t;
      q += __popc(m);
      cand += 32;
    }
    __syncwarp();
    if (lane == 0) s_cand[wib][0] = cand;
    if (q == 0) break;
    const int take = q < 32 ? q : 32;
    const bool active = lane < take;
    const unsigned long long my_tile = queue[wib][active ? lane : 0];
    __syncwarp();
    if (lane + 32 < q) queue[wib][lane] = queue[wib][lane + 32];
    q -= take;
    __syncwarp();

    // ---- one tile per lane -----------------------------------------------------------------------
    const unsigned long long s = (a.tile_first + my_tile) << c;
    const unsigned long long g = s ^ (s >> 1);
    double x[N];
#pragma unroll
    for (int j = 0; j < N; ++j) x[j] = a.xbase[j];
    for (int k = c - 1; k < N - 1; ++k) {
      const double f = (double)((g >> k) & 1ull);
      const double* col = sm + L::COLT + k * NP;
#pragma unroll
      for (int j = 0; j < N; ++j) x[j] = fma(f, col[j], x[j]);
    }
    const int nblk = 1 << (c - B);
    unsigned long long i0 = s;
#pragma unroll 1
    for (int blk = 0; blk < nblk; ++blk, i0 += (unsigned long long)NB) {
      const int k = (blk != 0) ? (B + __ffs(blk) - 1) : B;
      const double sg = (blk != 0) ? (((i0 >> (k + 1)) & 1ull) ? -1.0 : 1.0) : 0.0;
      const double sg_top = (blk & 1) ? -1.0 : 1.0;
      const uint32_t hi_addr = sm_colT + (uint32_t)(k * NP * 8);

      // Rows are handled in groups of G: one uniform branch per group instead of one per row, so
      // the G rows of a group are straight-line code whose chains interleave.  Hg = H rounded up
      // to a whole group (a cold row treated as hot just adds zeros).
      // cold groups: one update, one multiply per row and block.  The cold rows are the contiguous
      // range [Hg, N): a single uniform jump into an unrolled sequence (switch with fall-through)
      // gives the scheduler ONE basic block for all of them, so the independent updates and the
      // four product chains overlap instead of waiting on each other group by group.
      double q0 = 1.0, q1 = 1.0, q2 = 1.0, q3 = 1.0;
#define SPB_COLD_GROUP(gidx)                                                   \
  case gidx:                                                                   \
    if constexpr (G * (gidx) < N) {                                            \
      _Pragma("unroll") for (int j = G * (gidx); j < G * (gidx) + G && j < N; ++j) { \
        double d;                                                              \
        lds_f64(hi_addr + (uint32_t)(j * 8), d);                               \
        x[j] = fma(sg, d, x[j]);                                               \
        if ((j & 3) == 0) q0 *= x[j];                                          \
        else if ((j & 3) == 1) q1 *= x[j];                                     \
        else if ((j & 3) == 2) q2 *= x[j];                                     \
        else q3 *= x[j];                                                       \
      }                                                                        \
    }                                                                          \
    [[fallthrough]];
      switch (Hg / G) {
        SPB_COLD_GROUP(0) SPB_COLD_GROUP(1) SPB_COLD_GROUP(2) SPB_COLD_GROUP(3)
        SPB_COLD_GROUP(4) SPB_COLD_GROUP(5) SPB_COLD_GROUP(6) SPB_COLD_GROUP(7)
        SPB_COLD_GROUP(8) SPB_COLD_GROUP(9) SPB_COLD_GROUP(10) SPB_COLD_GROUP(11)
        default: break;
      }
#undef SPB_COLD_GROUP
      const double Q = (q0 * q1) * (q2 * q3);
      const bool skip_blk = SKIP && __all_sync(0xffffffffu, !active || Q == 0.0);
      if (skip_blk) {
        // whole warp: the block contributes exact zeros; apply its net effect on the hot rows
        // (the high column, and column B-1 which is the only low column left flipped)
#pragma unroll
        for (int j0 = 0; j0 < N; j0 += G) {
          if (j0 < Hg) {
#pragma unroll
            for (int j = j0; j < j0 + G && j < N; ++j) {
              double d, mt;
              lds_f64(hi_addr + (uint32_t)(j * 8), d);
              lds_f64(sm_lowR + (uint32_t)((j * LB + (B - 1)) * 8), mt);
              x[j] = fma(sg_top, mt, fma(sg, d, x[j]));
            }
          }
        }
      } else {
        double P[NB];
#pragma unroll
        for (int u = 0; u < NB; ++u) P[u] = 1.0;
#pragma unroll
        for (int j0 = 0; j0 < N; j0 += G) {
          if (j0 < Hg) {
#pragma unroll
            for (int j = j0; j < j0 + G && j < N; ++j) {
              double m[LB];
#pragma unroll
              for (int qq = 0; qq < LB; qq += 2)
                lds_f64x2(sm_lowR + (uint32_t)((j * LB + qq) * 8), m[qq], m[qq + 1]);
              double d;
              lds_f64(hi_addr + (uint32_t)(j * 8), d);
              double v = fma(sg, d, x[j]);
              P[0] *= v;
#pragma unroll
              for (int u = 1; u < NB; ++u) {
                const int K = ctz_c(u);
                if (K == B - 1) v = fma(sg_top, m[K], v);
                else if (((u >> (K + 1)) & 1) == 0) v += m[K];
                else v -= m[K];
                P[u] *= v;
              }
              x[j] = v;
            }
          }
        }
        double blk_sum = 0.0;
#pragma unroll
        for (int u = 0; u < NB; u += 2) blk_sum += (P[u] - P[u + 1]);
        if (active) {                                   // inactive lanes run a placeholder tile
          s_acc[threadIdx.x] = fma(blk_sum, Q, s_acc[threadIdx.x]);
          s_vis[threadIdx.x] += 1u;
        }
      }
    }
  }

  const double acc = warp_sum(s_acc[threadIdx.x]);
  unsigned long long vis = s_vis[threadIdx.x];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) vis += __shfl_down_sync(0xffffffffu, vis, o);
  if (lane == 0) { warp_part[wib] = acc; warp_vis[wib] = vis; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    unsigned long long n = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) { v += warp_part[w]; n += warp_vis[w]; }
    a.partials[blockIdx.x] = v;
    a.visited[blockIdx.x] = n;
  }
}

}  // namespace spb
