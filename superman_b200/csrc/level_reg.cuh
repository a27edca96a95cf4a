// LevelRyser: SpaRyser / SkipPer kernel for sparse matrices, sm_100a.
//
// Same tile / block decomposition as ryser_reg.cuh (a thread owns a tile of 2^c Gray indices and
// walks it in aligned blocks of 2^B), but the work per index follows the STRUCTURE of the matrix:
// a row whose lowest non-zero column is L ("level" L) can only change when a column >= L flips.
//
//   hot rows (level < B): X in registers, in B*S fixed slots, S per level (the host packs the rows
//     into slots; free slots hold the neutral row x = 1, entries 0).  A level-L slot takes only
//     2^(B-L) distinct values inside a block, so it costs 2^(B-L) updates and multiplies into its
//     level's 2^(B-L) running products PL[L][.] instead of 2^B of each.  The 2^B terms of a block,
//     term_u = PL[0][u] * PL[1][u>>1] * ... * PL[B-1][u>>(B-1)] * Q, are summed with their signs by
//     pairing bottom-up: E_0[w] = PL0[2w] - PL0[2w+1], E_L[w] = PL_L[2w] E_{L-1}[2w] + PL_L[2w+1]
//     E_{L-1}[2w+1], sum = Q * E_{B-1}[0] -- 2^B + B - 1 instructions instead of 3 * 2^B - 2.
//     Everything here is compile-time structured: straight-line code, no runtime branch.
//   register-cold rows: the R cold rows of lowest level (the ones refreshed most often) also stay
//     in registers: one update and one multiply per block, in straight-line code.
//   cold rows (level >= B): X in shared memory (X[row][thread], conflict-free), sorted by level.
//     Their product Q is kept as suffix products SP[i] = prod(rows of level >= B+i): the block that
//     flips high column k only touches the rows of level <= k, refreshes SP[k-B .. 0] and reuses
//     SP[k-B+1].  Half of the blocks flip column B, a quarter column B+1, ...: the expected number
//     of cold rows touched per block is small.  Levels without rows share one SP slot with the
//     next level that has some (s_grp), and every row knows the slot it closes (s_slot, a dummy
//     slot for most rows), so the refresh is ONE flat loop over the touched rows.
//
// SkipPer (SKIP = true): terms with a zero cold row are exact zeros (Q == 0).  Tiles whose
// tile-constant rows (level >= c) contain a zero are dropped by a warp-wide filter (ballot +
// shared-memory queue compaction, full warps only); blocks with Q == 0 in every lane skip the hot
// work (__all_sync) and apply only its net effect.  All control flow is warp-uniform.
//
// Replaces kernel_xshared_coalescing_mshared_sparse / _skipper (gpu_exact_sparse.cu:455-670).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "ryser_reg.cuh"


namespace spb {

struct LevelArgs {
  // packed device image (doubles unless noted), see sp_sparse.cu: level_pack()
  const double* colT_hot;    // [(n-1) * HSP]   colT_hot[k*HSP + s]  = D[slot s][k]   (hot slots, then R register-cold)
  const double* lowR;        // [HS * LB]       lowR[s*LB + q]       = D[slot s][q], q < B
  const double* dcold;       // [(n-1) * NCP]   dcold[k*NCP + jc]    = D[cold row jc][k]
  const double* xb_hot;      // [HS + R]
  const double* xb_cold;     // [NC]
  const int* cold_start;     // [n - B + 2]     first cold row of level >= B+i
  double* partials;
  unsigned long long* visited;
  unsigned long long tile_first, n_tiles;
  int n, NC, NCP, HSP;
  int c;
  int tiles_per_warp;
};

template <int B, int S, int R>
struct LevelLayout {
  static constexpr int HS = B * S;           // level slots
  static constexpr int HT = HS + R;          // + register-cold rows
  static constexpr int HSP = HT + (HT & 1);
  static constexpr int LB = B + (B & 1);
};

// dynamic shared memory (doubles):  colT_hot | lowR | dcold | xb_hot | xb_cold | Xc[NC][T] | SP[c-B+2][T]
// then ints: cold_start[n-B+2] | grp[c-B+1] | slot[NC]
template <int B, int S, int R, int THREADS, int MINBLOCKS, bool SKIP>
__global__ void __launch_bounds__(THREADS, MINBLOCKS)
level_reg_kernel(const LevelArgs a) {
  using LL = LevelLayout<B, S, R>;
  constexpr int HS = LL::HS, HT = LL::HT, HSP = LL::HSP, LB = LL::LB, NB = 1 << B, WARPS = THREADS / 32;
  extern __shared__ __align__(16) double dsm[];
  const int n = a.n, NC = a.NC, NCP = a.NCP, c = a.c;
  const int nseg = c - B;                       // SP[0 .. nseg]
  double* s_colT = dsm;
  double* s_lowR = s_colT + (size_t)(n - 1) * HSP;
  double* s_dcold = s_lowR + HS * LB;
  double* s_xbh = s_dcold + (size_t)(n - 1) * NCP;
  double* s_xbc = s_xbh + HSP;
  double* s_X = s_xbc + NCP;                    // [NC][THREADS]
  double* s_SP = s_X + (size_t)NC * THREADS;    // [nseg + 2][THREADS]: one per group of levels + a dummy
  int* s_cs = reinterpret_cast<int*>(s_SP + (size_t)(nseg + 2) * THREADS);
  int* s_grp = s_cs + (n - B + 2);              // [nseg + 1]  SP slot of segment seg
  int* s_slot = s_grp + (nseg + 1);             // [NC]        SP slot closed by cold row jc (or the dummy)
  __shared__ double warp_part[WARPS];
  __shared__ unsigned long long warp_vis[WARPS];
  __shared__ unsigned long long queue[WARPS][64];

  for (int e = threadIdx.x; e < (n - 1) * HSP; e += THREADS) s_colT[e] = a.colT_hot[e];
  for (int e = threadIdx.x; e < HS * LB; e += THREADS) s_lowR[e] = a.lowR[e];
  for (int e = threadIdx.x; e < (n - 1) * NCP; e += THREADS) s_dcold[e] = a.dcold[e];
  for (int e = threadIdx.x; e < HT; e += THREADS) s_xbh[e] = a.xb_hot[e];
  for (int e = threadIdx.x; e < NC; e += THREADS) s_xbc[e] = a.xb_cold[e];
  for (int e = threadIdx.x; e < n - B + 2; e += THREADS) s_cs[e] = a.cold_start[e];
  __syncthreads();
  if (threadIdx.x == 0) {
    // segments with the same first row (levels without rows) share a slot
    int g = 0;
    s_grp[0] = 0;
    for (int seg = 1; seg <= nseg; ++seg) {
      if (s_cs[seg] != s_cs[seg - 1]) ++g;
      s_grp[seg] = g;
    }
    const int dummy = g + 1;
    for (int jc = 0; jc < NC; ++jc) s_slot[jc] = dummy;
    for (int seg = 0; seg <= nseg; ++seg)
      if (s_cs[seg] < NC) s_slot[s_cs[seg]] = s_grp[seg];
  }
  __syncthreads();

  const uint32_t sm_colT = (uint32_t)__cvta_generic_to_shared(s_colT);
  const uint32_t sm_lowR = (uint32_t)__cvta_generic_to_shared(s_lowR);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* myX = s_X + threadIdx.x;
  double* mySP = s_SP + threadIdx.x;
  const int tc_first = s_cs[nseg];              // cold rows >= this index are constant over a tile

  const unsigned long long wg = (unsigned long long)blockIdx.x * WARPS + wib;
  unsigned long long cand = wg * (unsigned long long)a.tiles_per_warp;
  unsigned long long cand_hi = cand + (unsigned long long)a.tiles_per_warp;
  if (cand_hi > a.n_tiles) cand_hi = a.n_tiles;

  double acc = 0.0;
  unsigned long long vis = 0;
  int q = 0;
  for (;;) {
    // ---- refill: test 32 candidate tiles per trip until a full warp of survivors is queued ----
    while (q < 32 && cand < cand_hi) {
      const unsigned long long t = cand + lane;
      bool alive = t < cand_hi;
      if (SKIP && tc_first < NC) {
        const unsigned long long s = (a.tile_first + (alive ? t : cand)) << c;
        const unsigned long long g = s ^ (s >> 1);
        for (int jc = tc_first; jc < NC; ++jc) {
          double xr = s_xbc[jc];
          for (int k = c; k < n - 1; ++k)
            xr = fma((double)((g >> k) & 1ull), s_dcold[k * NCP + jc], xr);
          alive = alive && (xr != 0.0);
        }
      }
      const unsigned m = __ballot_sync(0xffffffffu, alive);
      if (alive) queue[wib][q + __popc(m & ((1u << lane) - 1u))] = t;
      q += __popc(m);
      cand += 32;
    }
    __syncwarp();
    if (q == 0) break;
    const int take = q < 32 ? q : 32;
    const bool active = lane < take;
    const unsigned long long my_tile = queue[wib][active ? lane : 0];
    __syncwarp();
    if (lane + 32 < q) queue[wib][lane] = queue[wib][lane + 32];
    q -= take;
    __syncwarp();

    // ---- explicit X at the tile start (cf. gpu_exact_sparse.cu:497-503) -------------------------
    const unsigned long long s = (a.tile_first + my_tile) << c;
    const unsigned long long g = s ^ (s >> 1);
    double xh[HT];
#pragma unroll
    for (int i = 0; i < HT; ++i) xh[i] = s_xbh[i];
    for (int k = c - 1; k < n - 1; ++k) {
      const double f = (double)((g >> k) & 1ull);
      const double* col = s_colT + k * HSP;
#pragma unroll
      for (int i = 0; i < HT; ++i) xh[i] = fma(f, col[i], xh[i]);
    }
    {
      // cold rows, from the last (highest level) to the first, building the suffix products
      double run = 1.0;
      if (tc_first == NC) mySP[s_grp[nseg] * THREADS] = 1.0;     // no tile-constant rows: empty product
      for (int jc = NC - 1; jc >= 0; --jc) {
        double x = s_xbc[jc];
        for (int k = c - 1; k < n - 1; ++k)
          x = fma((double)((g >> k) & 1ull), s_dcold[k * NCP + jc], x);
        myX[jc * THREADS] = x;
        run *= x;
        mySP[s_slot[jc] * THREADS] = run;
      }
    }

    double tile_acc = 0.0;
    unsigned long long tile_vis = 0;
    const int nblk = 1 << (c - B);
    const int tile_odd = (int)((a.tile_first + my_tile) & 1ull);
#pragma unroll 1
    for (int blk = 0; blk < nblk; ++blk) {
      const int z = (blk != 0) ? (__ffs(blk) - 1) : 0;     // k - B
      const int k = B + z;
      const int up = (k + 1 < c) ? ((blk >> (z + 1)) & 1) : tile_odd;
      const double sg = (blk != 0) ? (up ? -1.0 : 1.0) : 0.0;
      const double sg_top = (blk & 1) ? -1.0 : 1.0;

      const uint32_t hi_addr = sm_colT + (uint32_t)(k * HSP * 8);
      // ---- cold rows of level <= k: update, refresh SP[z .. 0] ----
      double Q;
      if (blk != 0) {
        double run = mySP[s_grp[z + 1] * THREADS];
        const double* dk = s_dcold + k * NCP;
        double* px = myX + (size_t)s_cs[z + 1] * THREADS;
        int jc = s_cs[z + 1] - 1;
        for (; jc >= 0; --jc) {
          px -= THREADS;
          const double x = fma(sg, dk[jc], *px);
          *px = x;
          run *= x;
          mySP[s_slot[jc] * THREADS] = run;
        }
        Q = run;
      } else {
        Q = mySP[0];
      }

      // ---- register-cold rows: one update, one multiply ----
      if (R > 0) {
        double r0 = 1.0, r1 = 1.0;
#pragma unroll
        for (int i = HS; i < HT; ++i) {
          double d;
          lds_f64(hi_addr + (uint32_t)(i * 8), d);
          xh[i] = fma(sg, d, xh[i]);
          if (i & 1) r1 *= xh[i]; else r0 *= xh[i];
        }
        Q *= r0 * r1;
      }
      const bool skip_blk = SKIP && __all_sync(0xffffffffu, !active || Q == 0.0);
      if (skip_blk) {
        // exact zeros: only the block's net effect on the hot slots (high column + column B-1)
#pragma unroll
        for (int i = 0; i < HS; ++i) {
          double mt;
          lds_f64(sm_lowR + (uint32_t)((i * LB + (B - 1)) * 8), mt);
          double d;
          lds_f64(hi_addr + (uint32_t)(i * 8), d);
          xh[i] = fma(sg_top, mt, fma(sg, d, xh[i]));
        }
      } else {
        // ---- hot slots: level L takes 2^(B-L) values; the slots of a level advance together, CH
        // independent chains at a time, and multiply into the level's running products PL[L][w] ----
        constexpr int CH = (B * S + R > 20) ? 1 : (S % 3 == 0) ? 3 : (S % 2 == 0) ? 2 : 1;   // register budget
        double PL[2 * NB];
        // layout: level L occupies indices [off(L), off(L) + 2^(B-L)), off(L) = 2*NB - 2*(NB >> L)
#pragma unroll
        for (int L = 0; L < B; ++L) {
          const int off = 2 * NB - 2 * (NB >> L);
          const int cnt = NB >> L;
#pragma unroll
          for (int t0 = 0; t0 < S; t0 += CH) {
            double v[CH], m[CH][LB];
#pragma unroll
            for (int t = 0; t < CH; ++t) {
              const int i = L * S + t0 + t;
#pragma unroll
              for (int qq = (L & ~1); qq < LB; qq += 2)       // only columns >= L flip inside this level
                lds_f64x2(sm_lowR + (uint32_t)((i * LB + qq) * 8), m[t][qq], m[t][qq + 1]);
              double d;
              lds_f64(hi_addr + (uint32_t)(i * 8), d);
              v[t] = fma(sg, d, xh[i]);
            }
            {
              double pr = v[0];
#pragma unroll
              for (int t = 1; t < CH; ++t) pr *= v[t];
              PL[off] = (t0 == 0) ? pr : PL[off] * pr;
            }
#pragma unroll
            for (int w = 1; w < cnt; ++w) {
              const int u = w << L;
              const int K = ctz_c(u);
#pragma unroll
              for (int t = 0; t < CH; ++t) {
                if (K == B - 1) v[t] = fma(sg_top, m[t][K], v[t]);
                else if (((u >> (K + 1)) & 1) == 0) v[t] += m[t][K];
                else v[t] -= m[t][K];
              }
              double pr = v[0];
#pragma unroll
              for (int t = 1; t < CH; ++t) pr *= v[t];
              PL[off + w] = (t0 == 0) ? pr : PL[off + w] * pr;
            }
#pragma unroll
            for (int t = 0; t < CH; ++t) xh[L * S + t0 + t] = v[t];
          }
        }
        // signed pair sums bottom-up: E_0[w] = PL0[2w] - PL0[2w+1],
        // E_L[w] = PL_L[2w] E_{L-1}[2w] + PL_L[2w+1] E_{L-1}[2w+1]
        double E[NB / 2];
#pragma unroll
        for (int w = 0; w < NB / 2; ++w) E[w] = PL[2 * w] - PL[2 * w + 1];          // off(0) = 0
#pragma unroll
        for (int L = 1; L < B; ++L) {
          const int off = 2 * NB - 2 * (NB >> L);
#pragma unroll
          for (int w = 0; w < (NB >> (L + 1)); ++w)
            E[w] = fma(PL[off + 2 * w], E[2 * w], PL[off + 2 * w + 1] * E[2 * w + 1]);
        }
        tile_acc = fma(Q, E[0], tile_acc);
        tile_vis += 1;
      }
    }
    if (active) { acc += tile_acc; vis += tile_vis; }
    __syncwarp();
  }

  acc = warp_sum(acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) vis += __shfl_down_sync(0xffffffffu, vis, o);
  if (lane == 0) { warp_part[wib] = acc; warp_vis[wib] = vis; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    unsigned long long cnt = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) { v += warp_part[w]; cnt += warp_vis[w]; }
    a.partials[blockIdx.x] = v;
    a.visited[blockIdx.x] = cnt;
  }
}

}  // namespace spb
