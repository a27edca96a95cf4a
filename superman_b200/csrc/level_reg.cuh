// LevelRyser: SpaRyser / SkipPer kernel for sparse matrices, sm_100a.
//
// Same tile / block decomposition as ryser_reg.cuh (a thread owns a tile of 2^c Gray indices and
// walks it in aligned blocks of 2^B), but the work per index follows the STRUCTURE of the matrix:
// a row whose lowest non-zero column is L ("level" L) can only change when a column >= L flips.
//
//   hot rows (level < B): X in registers, in fixed slots: S0 for level 0 (the most expensive level:
//     2^B values per block) and S for each other level (the host packs the rows into slots; a free
//     slot takes a cold row in SpaRyser plans, else the neutral row x = 1, entries 0).  A level-L slot takes only
//     2^(B-L) distinct values inside a block, so it costs 2^(B-L) updates and multiplies into its
//     level's 2^(B-L) products PL_L[.] instead of 2^B of each.  The 2^B terms of a block,
//     term_u = PL_0[u] * PL_1[u>>1] * ... * PL_{B-1}[u>>(B-1)] * Q, are summed with their signs by
//     pairing bottom-up: T_0[w] = PL_0[2w] - PL_0[2w+1], T_L[w] = PL_L[2w] T_{L-1}[2w] + PL_L[2w+1]
//     T_{L-1}[2w+1], sum = Q * T_{B-1}[0] -- 2^B + B - 1 instructions instead of 3 * 2^B - 2.  Every
//     product is folded into T the moment it exists, so at most 2^(B-1) partial sums are alive and
//     the register file has room for the register-cold rows below.
//     Everything here is compile-time structured: straight-line code, no runtime branch.
//   register-cold rows: the R cold rows of lowest level (the ones refreshed most often) also stay
//     in registers: one update and one multiply per block, in straight-line code.
//   cold rows (level >= B): X in shared memory (X[row][thread], conflict-free), sorted by level.
//     Their product Q is kept as suffix products SP[j] = prod(rows j .. NC-1): the block that flips high
//     column k only touches the rows of level <= k (rows [0, top), top from the constant bank), refreshes
//     SP[top-1 .. 0] and reuses SP[top].  Half of the blocks flip column B, a quarter column B+1, ...: the
//     expected number of cold rows touched per block is small.
//
// As in the dense kernel, the direction of a register row's update is not a +/-1.0 factor but a
// choice between shared-memory images (D, -D, zeros; two low-column images), so the block loop's
// addresses are block-uniform; the one update whose direction depends on the tile's parity (the
// middle block of a tile) adds the column like an even tile, and odd tiles take it out twice first.
// Configurations with 13-20 hot slots (spl_uniform_low) have no low-column images: the entries are kernel
// parameters and reach the FP64 instructions as uniform-register operands (LDCU.128, DADD R, R, UR); column
// B-1's direction is then the factor sB of one FMA per slot and block.
//
// Work distribution: a persistent grid; every warp pulls chunks of tiles from an atomic counter and
// leaves one partial sum per chunk (bit-reproducible whichever warp took which chunk).
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
#include <type_traits>
#include "ryser_reg.cuh"
#include "superman_b200_level.h"

#define SPB_LV_MAXSEG 64    // n - B + 2 <= 63
#define SPB_LV_MAXLOW 128   // hot slots (<= 32) x low columns (<= 4)

namespace spb {

// compile-time loop: f(std::integral_constant<int, I>) for I in [I0, N)
template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

struct LevelArgs {
  // packed device image (doubles unless noted), see sp_sparse.cu: level_pack()
  const double* colT_hot;    // [(n-1) * HSP]   colT_hot[k*HSP + s]  = D[slot s][k]   (hot slots, then R register-cold)
  const double* lowR;        // [HS * LB]       lowR[s*LB + q]       = D[slot s][q], q < B
  double low[SPB_LV_MAXLOW]; //                 the same image by value, for the configurations that read it as uniform-register
                             //                 operands (spl_uniform_low)
  const double* dcold;       // [(n-1) * NCP]   dcold[k*NCP + jc]    = D[cold row jc][k]
  const double* xb_hot;      // [HS + R]
  const double* xb_cold;     // [NC]
  int cold_start[SPB_LV_MAXSEG];   // [n - B + 2]  first cold row of level >= B+i (read through the constant bank)
  double* partials;          // [n_chunks]
  unsigned long long* visited;   // [n_chunks]  blocks of 2^B indices evaluated
  unsigned int* queue;       // {next chunk, warps that have left}; zero between launches
  // two phases of tiles: chunks [0, n_chunks_big) hold tiles of 2^c indices starting at tile tile_first, the
  // chunks after them tiles of 2^c_small starting at tile tile_first_small (in units of 2^c_small): the last
  // part of a range goes out in small pieces so that the persistent warps finish together
  unsigned long long tile_first, n_tiles;
  unsigned long long tile_first_small, n_tiles_small;
  unsigned int n_chunks;     // chunk = tiles_per_warp consecutive tiles (both phases)
  unsigned int n_chunks_big;
  int n, NC, NCP, HSP;
  int c, c_small;
  int tiles_per_warp;
};

template <int B, int S0, int S, int R>
struct LevelLayout {
  static constexpr int HS = S0 + (B - 1) * S;   // level slots: [0, S0) level 0, then S per level
  static constexpr int HT = HS + R;             // + register-cold rows
  static constexpr int HSP = HT + (HT & 1);
  static constexpr int LB = B + (B & 1);
  __host__ __device__ static constexpr int base(int L) { return L == 0 ? 0 : S0 + (L - 1) * S; }
  __host__ __device__ static constexpr int count(int L) { return L == 0 ? S0 : S; }
};

// register-cold rows and resident blocks per SM that go with a slot configuration: include/superman_b200_level.h
// (spl_regcold, spl_minblocks), shared with the C host code that packs the matrix

// bytes of dynamic shared memory the kernel needs (host and device agree through this one function)
__host__ __device__ inline size_t level_smem_bytes(int n, int B, int HS, int HSP, int LB, int NC, int NCP, int threads) {
  const size_t dbl = 2 * (size_t)(n - 1) * HSP + HSP + (spl_uniform_low(HS) ? 0 : 2 * (size_t)HS * LB) + (size_t)(n - 1) * NCP + HSP + NCP +
                     (size_t)(NC + 2) * threads + (size_t)(NC + 1) * threads;
  return dbl * sizeof(double);
}

// dynamic shared memory (doubles):  colP | colN | zero | low0 | low1 | dcold | xb_hot | xb_cold | Xc[NC + 2][T] | SP[NC + 1][T]
// (rows NC and NC+1 of Xc hold each thread's running sum and evaluated-block count for its current chunk: they
// are kept out of the register file, which the block loop needs, and share the thread's X address register;
// SP[j] = product of the cold rows j .. NC-1, SP[NC] = 1)
template <int B, int S0, int S, int R, int THREADS, int MINBLOCKS, bool SKIP>
__global__ void __launch_bounds__(THREADS, MINBLOCKS)
level_reg_kernel(const __grid_constant__ LevelArgs a) {
  using LL = LevelLayout<B, S0, S, R>;
  constexpr int HS = LL::HS, HT = LL::HT, HSP = LL::HSP, LB = LL::LB, NB = 1 << B, WARPS = THREADS / 32;
  constexpr int G = (HT > 24) ? 2 : 4;          // cold rows refreshed together (fewer where registers are short)
  constexpr bool UROP = spl_uniform_low(HS) != 0;   // low-column entries as uniform-register operands (no image in shared memory)
  constexpr bool PAIR = true;                   // neighbouring slots fetch their high-column entries together
  extern __shared__ __align__(16) double dsm[];
  const int n = a.n, NC = a.NC, NCP = a.NCP;
  double* s_colP = dsm;                         //  D
  double* s_colN = s_colP + (size_t)(n - 1) * HSP;   // -D
  double* s_zero = s_colN + (size_t)(n - 1) * HSP;   //  0 (first block of a tile: X is already explicit)
  double* s_low0 = s_zero + HSP;
  double* s_low1 = s_low0 + (UROP ? 0 : HS * LB);   // column B-1 negated
  double* s_dcold = s_low1 + (UROP ? 0 : HS * LB);
  double* s_xbh = s_dcold + (size_t)(n - 1) * NCP;
  double* s_xbc = s_xbh + HSP;
  double* s_X = s_xbc + NCP;                    // [NC + 2][THREADS]
  double* s_SP = s_X + (size_t)(NC + 2) * THREADS;   // [NC + 1][THREADS]: suffix products of the cold rows
  __shared__ unsigned long long tq[WARPS][64];  // surviving tiles of the warp's current chunk
  __shared__ unsigned int s_chunk[WARPS];       // the chunk each warp is working on (same reason)

  for (int e = threadIdx.x; e < (n - 1) * HSP; e += THREADS) {
    const double v = a.colT_hot[e];
    s_colP[e] = v;
    s_colN[e] = -v;
  }
  for (int e = threadIdx.x; e < HSP; e += THREADS) s_zero[e] = 0.0;
  for (int e = threadIdx.x; e < (UROP ? 0 : HS * LB); e += THREADS) {
    const double v = a.lowR[e];
    s_low0[e] = v;
    s_low1[e] = (e % LB == B - 1) ? -v : v;
  }
  for (int e = threadIdx.x; e < (n - 1) * NCP; e += THREADS) s_dcold[e] = a.dcold[e];
  for (int e = threadIdx.x; e < HT; e += THREADS) s_xbh[e] = a.xb_hot[e];
  for (int e = threadIdx.x; e < NC; e += THREADS) s_xbc[e] = a.xb_cold[e];
  s_SP[(size_t)NC * THREADS + threadIdx.x] = 1.0;
  __syncthreads();

  const uint32_t sm_colP = (uint32_t)__cvta_generic_to_shared(s_colP);
  const uint32_t neg_off = (uint32_t)((size_t)(n - 1) * HSP * 8);       // colN - colP
  const uint32_t zero_off = 2u * neg_off;                              // zero - colP
  const uint32_t sm_low0 = (uint32_t)__cvta_generic_to_shared(s_low0);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* myX = s_X + threadIdx.x;
  double* mySP = s_SP + threadIdx.x;

#pragma unroll 1
  for (;;) {
    // ---- next chunk of tiles for this warp ----
    unsigned int chunk = 0;
    if (lane == 0) chunk = atomicAdd(&a.queue[0], 1u);
    chunk = __shfl_sync(0xffffffffu, chunk, 0);
    if (chunk >= a.n_chunks) break;
    if (lane == 0) s_chunk[wib] = chunk;
    // (SpaRyser launches have one phase: c stays a launch constant, out of the register file)
    const bool big = !SKIP || chunk < a.n_chunks_big;
    const int c = big ? a.c : a.c_small;
    const unsigned long long tile_first = big ? a.tile_first : a.tile_first_small;
    const unsigned long long n_tiles = big ? a.n_tiles : a.n_tiles_small;
    const int tc_first = a.cold_start[c - B];     // cold rows >= this index are constant over a tile
    const int nblk = 1 << (c - B);
    unsigned long long cand = (unsigned long long)(chunk - (big ? 0u : a.n_chunks_big)) * (unsigned long long)a.tiles_per_warp;
    unsigned long long cand_hi = cand + (unsigned long long)a.tiles_per_warp;
    if (cand_hi > n_tiles) cand_hi = n_tiles;

    myX[(size_t)NC * THREADS] = 0.0;
    reinterpret_cast<unsigned long long*>(myX)[(size_t)(NC + 1) * THREADS] = 0ull;
    int q = 0;
#pragma unroll 1
    for (;;) {
      // ---- refill: test 32 candidate tiles per trip until a full warp of survivors is queued ----
      while (q < 32 && cand < cand_hi) {
        const unsigned long long t = cand + lane;
        bool alive = t < cand_hi;
        if (SKIP && tc_first < NC) {
          const unsigned long long s = (tile_first + (alive ? t : cand)) << c;
          const unsigned long long g = s ^ (s >> 1);
          int jc = tc_first;
          for (; jc + 4 <= NC; jc += 4) {              // four rows at a time (independent chains)
            double x0 = s_xbc[jc], x1 = s_xbc[jc + 1], x2 = s_xbc[jc + 2], x3 = s_xbc[jc + 3];
            for (int k = c; k < n - 1; ++k) {
              const double f = (double)((g >> k) & 1ull);
              const double* dk = s_dcold + k * NCP + jc;
              x0 = fma(f, dk[0], x0); x1 = fma(f, dk[1], x1); x2 = fma(f, dk[2], x2); x3 = fma(f, dk[3], x3);
            }
            alive = alive && (x0 != 0.0) && (x1 != 0.0) && (x2 != 0.0) && (x3 != 0.0);
          }
          for (; jc < NC; ++jc) {
            double xr = s_xbc[jc];
            for (int k = c; k < n - 1; ++k)
              xr = fma((double)((g >> k) & 1ull), s_dcold[k * NCP + jc], xr);
            alive = alive && (xr != 0.0);
          }
        }
        const unsigned m = __ballot_sync(0xffffffffu, alive);
        if (alive) tq[wib][q + __popc(m & ((1u << lane) - 1u))] = t;
        q += __popc(m);
        cand += 32;
      }
      __syncwarp();
      if (q == 0) break;
      const int take = q < 32 ? q : 32;
      const bool active = lane < take;
      const unsigned long long my_tile = tq[wib][active ? lane : 0];
      __syncwarp();
      if (lane + 32 < q) tq[wib][lane] = tq[wib][lane + 32];
      q -= take;
      __syncwarp();

      // ---- explicit X at the tile start (cf. gpu_exact_sparse.cu:497-503) -----------------------
      const unsigned long long s = (tile_first + my_tile) << c;
      const unsigned long long g = s ^ (s >> 1);
      double xh[HT];
#pragma unroll
      for (int i = 0; i < HT; ++i) xh[i] = s_xbh[i];
      for (int k = c - 1; k < n - 1; ++k) {
        const double f = (double)((g >> k) & 1ull);
        const double* col = s_colP + k * HSP;
#pragma unroll
        for (int i = 0; i < HT; ++i) xh[i] = fma(f, col[i], xh[i]);
      }
      {
        // cold rows, from the last (highest level) to the first, building the suffix products
        double run = 1.0;
        for (int jc = NC - 1; jc >= 0; --jc) {
          double x = s_xbc[jc];
          for (int k = c - 1; k < n - 1; ++k)
            x = fma((double)((g >> k) & 1ull), s_dcold[k * NCP + jc], x);
          myX[jc * THREADS] = x;
          run *= x;
          mySP[jc * THREADS] = run;
        }
      }
      const int tile_odd = (int)((tile_first + my_tile) & 1ull);

#pragma unroll 1
      for (int blk = 0; blk < nblk; ++blk) {
        // high column flipped at the block start: k = B + ctz(blk), the same for the whole grid; it is
        // added when bit k+1 of the index is clear.  That bit is a bit of blk, except for the middle
        // block of the tile (k = c-1), where it is the tile's parity: handled by the correction below.
        const int z = (blk != 0) ? (__ffs(blk) - 1) : 0;     // k - B
        const int k = B + z;
        const int up = (k + 1 < c) ? ((blk >> (z + 1)) & 1) : 0;
        if (blk == (nblk >> 1)) {
          const double f = -2.0 * (double)tile_odd;
          const double* col = s_colP + (c - 1) * HSP;
#pragma unroll
          for (int i = 0; i < HT; ++i) xh[i] = fma(f, col[i], xh[i]);
        }
        const uint32_t hi_addr = sm_colP + ((blk != 0) ? (uint32_t)(k * HSP * 8) + (up ? neg_off : 0u) : zero_off);
        // column B-1 flips in the middle of the block; its direction is bit B of the index
        const uint32_t low_addr = sm_low0 + (uint32_t)((blk & 1) * (HS * LB * 8));
        const double sB = (blk & 1) ? -1.0 : 1.0;            // (the uniform-operand form of the same choice)

        // ---- cold rows of level <= k (rows [0, top)): update, refresh their suffix products (per-thread
        // direction).  G rows at a time, every load before the first store, so that the shared-memory
        // round trips of a group overlap; only the product chain is serial. ----
        double Q;
        {
          const int top = (blk != 0) ? a.cold_start[z + 1] : 0;
          double run = mySP[(size_t)top * THREADS];
          if (top > 0) {
            const int upt = (k + 1 < c) ? up : tile_odd;
            const double sg = upt ? -1.0 : 1.0;
            const double* dk = s_dcold + k * NCP;
            int jc = top;
            if (jc & (G - 1)) {
              // the rows that do not fill a group: same code with the missing rows switched off
              const int r = jc & (G - 1);
              jc -= r;
              double* px = myX + (size_t)jc * THREADS;
              double* ps = mySP + (size_t)jc * THREADS;
              double x[G], pr[G];
#pragma unroll
              for (int i = 0; i < G - 1; ++i) x[i] = fma(sg, (i < r) ? dk[jc + i] : 0.0, (i < r) ? px[i * THREADS] : 1.0);
#pragma unroll
              for (int i = G - 2; i >= 0; --i) { run *= x[i]; pr[i] = run; }
#pragma unroll
              for (int i = 0; i < G - 1; ++i)
                if (i < r) { px[i * THREADS] = x[i]; ps[i * THREADS] = pr[i]; }
            }
            while (jc > 0) {
              jc -= G;
              double* px = myX + (size_t)jc * THREADS;
              double* ps = mySP + (size_t)jc * THREADS;
              double x[G], pr[G];
#pragma unroll
              for (int i = 0; i < G; ++i) x[i] = fma(sg, dk[jc + i], px[i * THREADS]);
#pragma unroll
              for (int i = G - 1; i >= 0; --i) { run *= x[i]; pr[i] = run; }
#pragma unroll
              for (int i = 0; i < G; ++i) { px[i * THREADS] = x[i]; ps[i * THREADS] = pr[i]; }
            }
          }
          Q = run;
        }

        // ---- register-cold rows: one update, one multiply ----
        if (R > 0) {
          double r0 = 1.0, r1 = 1.0;
#pragma unroll
          for (int i = HS; i < HT; ++i) {
            if ((i & 1) == 0 && i + 1 < HT) {          // an aligned pair: one 16-byte load
              double d0, d1;
              lds_f64x2(hi_addr + (uint32_t)(i * 8), d0, d1);
              xh[i] += d0;
              xh[i + 1 < HT ? i + 1 : i] += d1;
              r0 *= xh[i];
              r1 *= xh[i + 1 < HT ? i + 1 : i];
            } else if ((i & 1) == 0 || i == HS) {
              double d;
              lds_f64(hi_addr + (uint32_t)(i * 8), d);
              xh[i] += d;
              if (i & 1) r1 *= xh[i]; else r0 *= xh[i];
            }
          }
          Q *= r0 * r1;
        }
        const bool skip_blk = SKIP && __all_sync(0xffffffffu, !active || Q == 0.0);
        if (skip_blk) {
          // exact zeros: only the block's net effect on the hot slots (high column + column B-1)
#pragma unroll
          for (int i = 0; i < HS; ++i) {
            double d;
            lds_f64(hi_addr + (uint32_t)(i * 8), d);
            if constexpr (UROP) {
              xh[i] = fma(sB, a.low[i * LB + (B - 1)], xh[i] + d);
            } else {
              double mt;
              lds_f64(low_addr + (uint32_t)((i * LB + (B - 1)) * 8), mt);
              xh[i] = (xh[i] + d) + mt;
            }
          }
        } else {
          // ---- hot slots: level L takes 2^(B-L) values; the slots of a level advance together, CH
          // independent chains at a time; the product of the level's slots for value w is folded into
          // the pair sums T as soon as the last chain group has contributed (all indices below are
          // compile-time constants once the loops are unrolled) ----
          double T[NB / 2];
          static_for<0, B>([&](auto Lc) {
            constexpr int L = decltype(Lc)::value;
            constexpr int SL = LL::count(L), base = LL::base(L);
            constexpr int NG = (SL + 2) / 3;        // chain groups of at most 3 slots, sizes as even as possible
            constexpr int cnt = NB >> L;
            double PL[NG > 1 ? cnt : 1];            // products across chain groups
            double tmp = 0.0, prev = 0.0;
            static_for<0, NG>([&](auto Gc) {
              constexpr int gi = decltype(Gc)::value;
              constexpr int t0 = gi * (SL / NG) + (gi < SL % NG ? gi : SL % NG);
              constexpr int CH = SL / NG + (gi < SL % NG ? 1 : 0);
              [[maybe_unused]] double m[UROP ? 1 : CH][LB];
              double v[CH], d[CH];
              // the high column's entries of the group's slots: neighbours share one 16-byte load
#pragma unroll
              for (int t = 0; t < CH; ++t) {
                const int i = base + t0 + t;
                if (!PAIR) {
                  lds_f64(hi_addr + (uint32_t)(i * 8), d[t]);
                } else if ((i & 1) == 0) {
                  if (t + 1 < CH) lds_f64x2(hi_addr + (uint32_t)(i * 8), d[t], d[t + 1 < CH ? t + 1 : t]);
                  else lds_f64(hi_addr + (uint32_t)(i * 8), d[t]);
                } else if (t == 0) {
                  lds_f64(hi_addr + (uint32_t)(i * 8), d[t]);
                }                                       // (an odd slot after the first came with its neighbour)
              }
#pragma unroll
              for (int t = 0; t < CH; ++t) {
                const int i = base + t0 + t;
                if constexpr (!UROP) {
#pragma unroll
                  for (int qq = (L & ~1); qq < LB; qq += 2)     // only columns >= L flip inside this level
                    lds_f64x2(low_addr + (uint32_t)((i * LB + qq) * 8), m[t][qq], m[t][qq + 1]);
                }
                v[t] = xh[i] + d[t];
              }
#pragma unroll
              for (int w = 0; w < cnt; ++w) {
                if (w > 0) {
                  const int u = w << L;
                  const int K = ctz_c(u);
#pragma unroll
                  for (int t = 0; t < CH; ++t) {
                    if constexpr (UROP) {
                      const double mk = a.low[(base + t0 + t) * LB + K];    // constant bank -> uniform register
                      if (K == B - 1) v[t] = fma(sB, mk, v[t]);
                      else if (((u >> (K + 1)) & 1) == 0) v[t] += mk;
                      else v[t] -= mk;
                    } else {
                      if (K == B - 1 || ((u >> (K + 1)) & 1) == 0) v[t] += m[t][K];
                      else v[t] -= m[t][K];
                    }
                  }
                }
                double pr = v[0];
#pragma unroll
                for (int t = 1; t < CH; ++t) pr *= v[t];
                if constexpr (NG > 1) {
                  if (gi > 0) pr *= PL[w];
                  if (gi < NG - 1) { PL[w] = pr; continue; }
                }
                // fold: level 0 pairs up with alternating signs, level L combines the sums below it
                if (L == 0) {
                  if ((w & 1) == 0) prev = pr; else T[w >> 1] = prev - pr;
                } else {
                  if ((w & 1) == 0) tmp = pr * T[w]; else T[w >> 1] = fma(pr, T[w], tmp);
                }
              }
#pragma unroll
              for (int t = 0; t < CH; ++t) xh[base + t0 + t] = v[t];
            });
          });
          if (active) {                               // idle lanes of a last, partial round repeat lane 0's tile
            myX[(size_t)NC * THREADS] = fma(Q, T[0], myX[(size_t)NC * THREADS]);
            if (SKIP) reinterpret_cast<unsigned long long*>(myX)[(size_t)(NC + 1) * THREADS] += 1ull;
          }
        }
      }
      if (!SKIP && active)                            // SpaRyser evaluates every block: counted once per tile
        reinterpret_cast<unsigned long long*>(myX)[(size_t)(NC + 1) * THREADS] += (unsigned long long)nblk;
      __syncwarp();
    }

    const double acc = warp_sum(myX[(size_t)NC * THREADS]);
    unsigned long long vis = reinterpret_cast<unsigned long long*>(myX)[(size_t)(NC + 1) * THREADS];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vis += __shfl_down_sync(0xffffffffu, vis, o);
    if (lane == 0) { const unsigned int ch = s_chunk[wib]; a.partials[ch] = acc; a.visited[ch] = vis; }
  }

  if (lane == 0) {
    __threadfence();
    if (atomicAdd(&a.queue[1], 1u) == gridDim.x * WARPS - 1) {   // nobody will touch the counters any more
      a.queue[0] = 0u;
      a.queue[1] = 0u;
      __threadfence();
    }
  }
}

}  // namespace spb
