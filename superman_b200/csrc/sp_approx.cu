// Rasmussen and Sinkhorn-scaled permanent estimators on one device, warp-per-trial.
//
// Reference paths replaced: kernel_rasmussen_sparse / kernel_approximation_sparse
// (gpu_approximation_sparse.cu:198-290, 292-452) and their dense twins kernel_rasmussen /
// kernel_approximation (gpu_approximation_dense.cu:155-229, 231-369), which run one trial per
// THREAD with the bit masks in local memory, rescan every CRS row at every step and keep the
// scaling vectors d_r / d_c in global memory at stride nov (2 x 2.7 GB per GPU at nov = 648,
// gpu_approximation_sparse.cu:742-743).
//
// Here one WARP runs one trial; all of a trial's state lives in shared memory:
//   deg[r]      remaining column count of row r (bytes; maintained incrementally when a column
//               is removed, instead of recounted from CRS at every step),
//   rowx/colx   "extracted" bit masks,
//   d_r, d_c    Sinkhorn scaling vectors (scaling estimator only).
// The lanes split the rows (minimum-degree search, redux.min on (deg << 16 | row) = first minimum
// in ascending row order, as the reference's scan), the entries of the chosen row (ballot + rank
// select of the r-th remaining column) and the columns / rows of a Sinkhorn sweep.  Every branch
// is warp-uniform.
// Random numbers: counter-based Philox4x32-10, counter = (trial index, draw/4), key = seed, so a
// trial's outcome depends only on (seed, trial index) -- reproducible and independent of how the
// trials are split over devices (the reference seeds XORWOW with time(0)*tid,
// gpu_approximation_sparse.cu:226,473).
//
// Compiled with -fmad=false: the oracle restates the same float / double operations in C and the
// per-trial values are compared bit for bit.
#include "sp_internal.cuh"
#include <string.h>
#include <new>
#include <type_traits>
#include <vector>

namespace spb {

#define APX_WARPS 16
#define APX_THREADS (APX_WARPS * 32)

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct ApproxArgs {
  const int* rptrs; const int* cols;   // CRS pattern
  const int* cptrs; const int* rows;   // CCS pattern
  const double* rvals; const double* cvals;  // entry weights (dense twins) or nullptr
  const unsigned short* ell_rows;      // [nov * W] rows of column j, CCS order, padded with nov (W > 0 only)
  const unsigned short* ell_cols;      // [nov * W] columns of row i, CRS order, padded with nov
  double* est;                         // [trial_hi - trial_lo] every trial's estimate (0 for a dead end)
  unsigned int* queue;                 // next trial to hand out, relative to trial_lo (zeroed before the launch)
  double* trace;                       // tests only (else nullptr): per trial {steps completed, running product}
  unsigned long long trial_lo, trial_hi;
  unsigned long long seed;
  double sq_scale;
  int nov, nnz;
  int scaling;                         // 0 Rasmussen, 1 scaled
  int scale_intervals, scale_times;
};

__device__ __forceinline__ bool bit_test(const unsigned* m, int i) { return (m[i >> 5] >> (i & 31)) & 1u; }

// WEIGHTED: Sinkhorn sums use the entry values in double (dense twin, gpu_approximation_dense.cu:
// 286-313); otherwise pattern only with float sums (gpu_approximation_sparse.cu:361-396).
// W > 0 (pattern only, every row and column has at most W entries): the Sinkhorn sweeps read the
// pattern in ELL form, W 16-bit indices per row / column padded with the index nov, whose scaling
// factor is a constant 0 -- fixed trip count, no pointer loads, and adding the padding's exact
// zeros after the real entries leaves every float sum bit-identical.
// With W > 0 the ELL arrays are the only copy of the pattern in shared memory (the column pick and the degree
// updates read them too), which together with byte-sized degrees (WIDE = false: no row has more than 255
// entries) lets 32 warps = 32 concurrent trials share an SM on the 36 x 36 grid instead of 16.
template <bool WEIGHTED, int W, bool WIDE>
__global__ void __launch_bounds__(APX_THREADS, 2)
approx_kernel(const ApproxArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  typedef typename std::conditional<WIDE, unsigned short, unsigned char>::type deg_t;
  const int nov = a.nov, nnz = (W > 0) ? 0 : a.nnz;
  const int words = (nov + 31) >> 5;
  // block-shared pattern (CRS + CCS; absent when the ELL form is used)
  int* s_rptrs = reinterpret_cast<int*>(smraw);
  int* s_cptrs = s_rptrs + (W > 0 ? 0 : nov + 1);
  int* s_cols = s_cptrs + (W > 0 ? 0 : nov + 1);
  int* s_rows = s_cols + nnz;
  size_t off = (W > 0) ? 0 : (size_t)(2 * (nov + 1) + 2 * nnz) * sizeof(int);
  off = (off + 15) & ~(size_t)15;
  unsigned short* s_ell_rows = reinterpret_cast<unsigned short*>(smraw + off);
  unsigned short* s_ell_cols = s_ell_rows + (size_t)nov * W;
  off += 2 * (size_t)nov * W * sizeof(unsigned short);
  off = (off + 15) & ~(size_t)15;
  // per-warp state
  const int deg_bytes = ((int)sizeof(deg_t) * nov + 15) & ~15;   // 16-bit degrees when a row may hold more than 255 entries
  const size_t warp_bytes = (size_t)deg_bytes + 2 * (size_t)words * 4 + (a.scaling ? 2 * (size_t)(nov + 1) * 4 : 0);
  const size_t warp_stride = (warp_bytes + 15) & ~(size_t)15;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  unsigned char* wbase = smraw + off + wib * warp_stride;
  deg_t* deg = reinterpret_cast<deg_t*>(wbase);
  unsigned* rowx = reinterpret_cast<unsigned*>(wbase + deg_bytes);
  unsigned* colx = rowx + words;
  float* d_r = reinterpret_cast<float*>(colx + words);
  float* d_c = d_r + (nov + 1);                 // d_r[nov] = d_c[nov] = 0: the ELL padding
  __shared__ unsigned int s_tix[APX_WARPS];

  if (W == 0) {
    for (int e = threadIdx.x; e <= nov; e += APX_THREADS) { s_rptrs[e] = a.rptrs[e]; s_cptrs[e] = a.cptrs[e]; }
    for (int e = threadIdx.x; e < nnz; e += APX_THREADS) { s_cols[e] = a.cols[e]; s_rows[e] = a.rows[e]; }
  } else {
    for (int e = threadIdx.x; e < nov * W; e += APX_THREADS) { s_ell_rows[e] = a.ell_rows[e]; s_ell_cols[e] = a.ell_cols[e]; }
  }
  __syncthreads();

  const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
  const unsigned int n_trials = (unsigned int)(a.trial_hi - a.trial_lo);

  // Trials differ wildly in length (most hit a dead end early, a survivor runs nov steps with a Sinkhorn
  // phase every scale_intervals of them): every warp pulls its next trial from a global counter instead of
  // owning a fixed stride of them (a static split left 19 % of the warp-time idle at the end of the launch).
  // Each estimate goes to its own slot, so the sums do not depend on which warp ran which trial.
  for (;;) {
    unsigned int tix = 0;
    if (lane == 0) tix = atomicAdd(a.queue, 1u);
    tix = __shfl_sync(0xffffffffu, tix, 0);
    if (tix >= n_trials) break;
    // the trial index is read back from shared memory where it is needed (every fourth step and at the
    // end): one register less to carry through the whole trial
    if (lane == 0) s_tix[wib] = tix;
    __syncwarp();
#define SPB_TRIAL() (a.trial_lo + (unsigned long long)s_tix[wib])
    // ---- reset state ----
    for (int r = lane; r < nov; r += 32) {
      int d0;
      if (W > 0) {
        d0 = 0;
#pragma unroll
        for (int q = 0; q < W; ++q) d0 += (s_ell_cols[r * W + q] != (unsigned short)nov);
      } else {
        d0 = s_rptrs[r + 1] - s_rptrs[r];
      }
      deg[r] = (deg_t)d0;
    }
    for (int w = lane; w < words; w += 32) { rowx[w] = 0u; colx[w] = 0u; }
    if (a.scaling) {
      for (int i = lane; i < nov; i += 32) { d_r[i] = 1.0f; d_c[i] = 1.0f; }
      if (lane == 0) { d_r[nov] = 0.0f; d_c[nov] = 0.0f; }
    }
    __syncwarp();
    double perm = 1.0;
    uint32_t rnd[4];
    bool dead = false;
    int step = 0;                 // steps completed so far
    for (; step < nov && !dead; ++step) {
      if ((step & 3) == 0) {
        const unsigned long long trial = SPB_TRIAL();
        philox4x32_10((uint32_t)trial, (uint32_t)(trial >> 32), (uint32_t)(step >> 2), 0u, k0, k1, rnd);
      }
      const int rw = step & 3;   // select, not index: keeps the four words in registers
      const uint32_t draw = (rw == 0) ? rnd[0] : (rw == 1) ? rnd[1] : (rw == 2) ? rnd[2] : rnd[3];
      // ---- minimum-degree remaining row, first in ascending order ----
      unsigned best = 0xffffffffu;
      for (int r = lane; r < nov; r += 32)
        if (!bit_test(rowx, r)) best = min(best, ((unsigned)deg[r] << 16) | (unsigned)r);
      best = __reduce_min_sync(0xffffffffu, best);
      const int row = (int)(best & 0xffffu);
      const int dmin = (int)(best >> 16);
      if (dmin == 0) { dead = true; break; }
      const int rb = (W > 0) ? 0 : s_rptrs[row], re = (W > 0) ? 0 : s_rptrs[row + 1];
      int col = -1;
      if (W == 0 && !a.scaling) {
        // ---- Rasmussen: perm *= deg; uniform pick of the r-th remaining column ----
        perm *= (double)dmin;
        int want = (int)(((uint64_t)draw * (uint64_t)dmin) >> 32);
        for (int base = rb; base < re && col < 0; base += 32) {
          const int e = base + lane;
          const int cc = (e < re) ? s_cols[e] : 0;
          const bool rem = (e < re) && !bit_test(colx, cc);
          const unsigned m = __ballot_sync(0xffffffffu, rem);
          const int cnt = __popc(m);
          if (want < cnt) {
            const int src = __fns(m, 0, want + 1);
            col = __shfl_sync(0xffffffffu, cc, src);
          } else {
            want -= cnt;
          }
        }
      } else {
        // ---- scaled estimator: Sinkhorn sweeps every scale_intervals steps ----
        if (step % a.scale_intervals == 0) {
          for (int sweep = 0; sweep < a.scale_times && !dead; ++sweep) {
            bool zero = false;
            for (int j = lane; j < nov; j += 32) {
              if (bit_test(colx, j)) continue;
              if (WEIGHTED) {
                double cs = 0.0;
                for (int t = s_cptrs[j]; t < s_cptrs[j + 1]; ++t) {
                  cs += (double)d_r[s_rows[t]] * a.cvals[t];          // extracted rows hold d_r = 0
                }
                if (cs == 0.0) zero = true; else d_c[j] = (float)(1.0 / cs);
              } else if (W > 0) {
                const ushort4* er = reinterpret_cast<const ushort4*>(s_ell_rows + j * W);
                float cs = 0.0f;
#pragma unroll
                for (int q = 0; q < W / 4; ++q) {
                  const ushort4 e4 = er[q];
                  cs += d_r[e4.x]; cs += d_r[e4.y]; cs += d_r[e4.z]; cs += d_r[e4.w];
                }
                if (cs == 0.0f) zero = true; else d_c[j] = 1.0f / cs;
              } else {
                float cs = 0.0f;
                for (int t = s_cptrs[j]; t < s_cptrs[j + 1]; ++t) {
                  cs += d_r[s_rows[t]];                               // extracted rows hold d_r = 0
                }
                if (cs == 0.0f) zero = true; else d_c[j] = 1.0f / cs;
              }
            }
            if (__any_sync(0xffffffffu, zero)) { dead = true; break; }
            __syncwarp();
            for (int i = lane; i < nov; i += 32) {
              if (bit_test(rowx, i)) continue;
              if (WEIGHTED) {
                double rs = 0.0;
                for (int t = s_rptrs[i]; t < s_rptrs[i + 1]; ++t) {
                  rs += a.rvals[t] * (double)d_c[s_cols[t]];          // extracted columns hold d_c = 0
                }
                if (rs == 0.0) zero = true; else d_r[i] = (float)(1.0 / rs);
              } else if (W > 0) {
                const ushort4* ec = reinterpret_cast<const ushort4*>(s_ell_cols + i * W);
                float rs = 0.0f;
#pragma unroll
                for (int q = 0; q < W / 4; ++q) {
                  const ushort4 e4 = ec[q];
                  rs += d_c[e4.x]; rs += d_c[e4.y]; rs += d_c[e4.z]; rs += d_c[e4.w];
                }
                if (rs == 0.0f) zero = true; else d_r[i] = 1.0f / rs;
              } else {
                float rs = 0.0f;
                for (int t = s_rptrs[i]; t < s_rptrs[i + 1]; ++t) {
                  rs += d_c[s_cols[t]];                               // extracted columns hold d_c = 0
                }
                if (rs == 0.0f) zero = true; else d_r[i] = 1.0f / rs;
              }
            }
            if (__any_sync(0xffffffffu, zero)) { dead = true; break; }
            __syncwarp();
          }
          if (dead) break;
        }
        // ---- column with probability d_r[row]*d_c[c] / sum (all lanes compute the same) ----
        const float dr = d_r[row];
        double tot = 0.0;
        if (W > 0) {
          // the row's entries in CRS order, then padding (index nov, factor 0): the same sums bit for bit
#pragma unroll 1
          for (int q = 0; q < W; ++q) tot += (double)(dr * d_c[s_ell_cols[row * W + q]]);
        } else {
          for (int t = rb; t < re; ++t) tot += (double)(dr * d_c[s_cols[t]]);   // + 0 for extracted columns
        }
        if (tot == 0.0) { dead = true; break; }
        const double target = ((double)draw + 1.0) * (1.0 / 4294967296.0) * tot;
        double run = 0.0;
        if (W > 0) {
#pragma unroll 1
          for (int q = 0; q < W; ++q) {
            const int cc = s_ell_cols[row * W + q];
            const double s = (double)(dr * d_c[cc]);
            run += s;
            if (col < 0 && cc < nov && target <= run) { col = cc; perm /= (s / tot); }
          }
        } else {
          for (int t = rb; t < re; ++t) {
            const int cc = s_cols[t];
            const double s = (double)(dr * d_c[cc]);     // 0 for an extracted column: run does not move
            run += s;
            if (target <= run) { col = cc; perm /= (s / tot); break; }
          }
        }
        if (col < 0) { dead = true; break; }   // cannot happen: run reaches tot exactly
      }
      // ---- extract row and column; lower the degree of the other rows of that column ----
      __syncwarp();
      if (lane == 0) {
        rowx[row >> 5] |= 1u << (row & 31);
        colx[col >> 5] |= 1u << (col & 31);
        // an extracted row / column keeps scaling factor 0: the Sinkhorn sums and the column pick
        // add exact zeros for it instead of testing the bit sets per entry
        if (a.scaling) { d_r[row] = 0.0f; d_c[col] = 0.0f; }
      }
      __syncwarp();
      if (W > 0) {
        if (lane < W) {
          const int r = s_ell_rows[col * W + lane];
          if (r < nov && !bit_test(rowx, r)) deg[r] -= 1;
        }
      } else {
        for (int t = s_cptrs[col] + lane; t < s_cptrs[col + 1]; t += 32) {
          const int r = s_rows[t];
          if (!bit_test(rowx, r)) deg[r] -= 1;
        }
      }
      __syncwarp();
    }
    // a dead end estimates 0; `perm` is still the product over the `step` steps it completed
    if (a.trace && lane == 0) {
      a.trace[2 * (size_t)s_tix[wib]] = (double)step;
      a.trace[2 * (size_t)s_tix[wib] + 1] = perm;
    }
    if (lane == 0) a.est[s_tix[wib]] = dead ? 0.0 : perm;
    __syncwarp();
#undef SPB_TRIAL
  }
}


// ---- thread-per-trial variant for nov <= 64 -----------------------------------------------------
// With at most 64 rows and columns a row's pattern is one 64-bit word: the remaining degree of a
// row is popc(rowmask[r] & ~colx), read from shared memory at a warp-uniform address (broadcast),
// the two "extracted" sets live in two registers, and one THREAD runs a trial -- 32 trials per warp
// instead of one.  Lanes are persistent: a lane whose trial ends (dead end or last step) starts
// its next trial in the following loop trip, so the warp stays full until the range is exhausted.
// Same estimators, same Philox stream, same order of every floating-point operation as the
// warp-per-trial kernel above and as the oracle (per-trial values are bit-identical).
#define APS_THREADS 128

struct SmallArgs {
  const unsigned long long* rowmask;   // [nov] columns of row r
  const unsigned long long* colmask;   // [nov] rows of column c
  const double* wdense;                // [nov*nov] entry weights (scaled dense twin) or nullptr
  double* est;
  unsigned int* queue;
  double* trace;
  unsigned long long trial_lo, trial_hi;
  unsigned long long seed;
  double sq_scale;
  int nov;
  int scaling, scale_intervals, scale_times;
};

template <bool SCALING, bool WEIGHTED>
__global__ void __launch_bounds__(APS_THREADS)
approx_small_kernel(const SmallArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int nov = a.nov;
  unsigned long long* s_row = reinterpret_cast<unsigned long long*>(smraw);
  unsigned long long* s_col = s_row + nov;
  double* s_w = reinterpret_cast<double*>(s_col + nov);                 // WEIGHTED: nov*nov
  float* s_d = reinterpret_cast<float*>(s_w + (WEIGHTED ? nov * nov : 0));   // SCALING: d_r | d_c, [i][thread]
  float* d_r = s_d + threadIdx.x;
  float* d_c = s_d + (size_t)nov * APS_THREADS + threadIdx.x;
  for (int e = threadIdx.x; e < nov; e += APS_THREADS) { s_row[e] = a.rowmask[e]; s_col[e] = a.colmask[e]; }
  if (WEIGHTED) for (int e = threadIdx.x; e < nov * nov; e += APS_THREADS) s_w[e] = a.wdense[e];
  __syncthreads();

  // persistent lanes pulling trial indices from a global counter (see approx_kernel); one slot per estimate
  const unsigned int n_trials = (unsigned int)(a.trial_hi - a.trial_lo);
  unsigned int tix = atomicAdd(a.queue, 1u);
  unsigned long long trial = a.trial_lo + tix;
  const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
  double perm = 1.0;
  unsigned long long rowx = 0ull, colx = 0ull;
  uint32_t rnd0 = 0, rnd1 = 0, rnd2 = 0, rnd3 = 0;
  int step = 0;

  for (int it = 0;; ++it) {
    const bool have = tix < n_trials;
    if (!__any_sync(0xffffffffu, have)) break;
    if (!have) continue;
    // scaled estimator: a new trial only starts on a loop trip that is a multiple of the scaling
    // interval, so the (expensive) Sinkhorn steps of the 32 lanes coincide instead of making every
    // trip pay for the one lane that happens to be scaling
    if (SCALING && step == 0 && (it % a.scale_intervals) != 0) continue;
    if (step == 0) {
      rowx = 0ull; colx = 0ull; perm = 1.0;
      if (SCALING) for (int i = 0; i < nov; ++i) { d_r[i * APS_THREADS] = 1.0f; d_c[i * APS_THREADS] = 1.0f; }
    }
    if ((step & 3) == 0) {
      uint32_t r[4];
      philox4x32_10((uint32_t)trial, (uint32_t)(trial >> 32), (uint32_t)(step >> 2), 0u, k0, k1, r);
      rnd0 = r[0]; rnd1 = r[1]; rnd2 = r[2]; rnd3 = r[3];
    }
    const int rw = step & 3;
    const uint32_t draw = (rw == 0) ? rnd0 : (rw == 1) ? rnd1 : (rw == 2) ? rnd2 : rnd3;
    // ---- minimum-degree remaining row, first in ascending order ----
    unsigned best = 0xffffffffu;
    for (int r = 0; r < nov; ++r) {
      const unsigned d = (unsigned)__popcll(s_row[r] & ~colx);
      const unsigned key = ((rowx >> r) & 1ull) ? 0xffffffffu : ((d << 8) | (unsigned)r);
      best = min(best, key);
    }
    const int row = (int)(best & 0xffu);
    const int dmin = (int)(best >> 8);
    bool dead = (dmin == 0);
    int col = 0;
    if (!dead) {
      const unsigned long long avail = s_row[row] & ~colx;
      if (!SCALING) {
        perm *= (double)dmin;
        const int want = (int)(((uint64_t)draw * (uint64_t)dmin) >> 32);
        // want-th (0-based) set bit of avail
        const unsigned lo32 = (unsigned)avail, hi32 = (unsigned)(avail >> 32);
        const int nlo = __popc(lo32);
        col = (want < nlo) ? (int)__fns(lo32, 0, want + 1) : 32 + (int)__fns(hi32, 0, want - nlo + 1);
      } else {
        if (step % a.scale_intervals == 0) {
          for (int sweep = 0; sweep < a.scale_times && !dead; ++sweep) {
            for (int j = 0; j < nov && !dead; ++j) {
              if ((colx >> j) & 1ull) continue;
              unsigned long long m = s_col[j] & ~rowx;
              if (WEIGHTED) {
                double cs = 0.0;
                while (m) { const int r = __ffsll((long long)m) - 1; m &= m - 1ull; cs += (double)d_r[r * APS_THREADS] * s_w[r * nov + j]; }
                if (cs == 0.0) dead = true; else d_c[j * APS_THREADS] = (float)(1.0 / cs);
              } else {
                float cs = 0.0f;
                while (m) { const int r = __ffsll((long long)m) - 1; m &= m - 1ull; cs += d_r[r * APS_THREADS]; }
                if (cs == 0.0f) dead = true; else d_c[j * APS_THREADS] = 1.0f / cs;
              }
            }
            for (int i = 0; i < nov && !dead; ++i) {
              if ((rowx >> i) & 1ull) continue;
              unsigned long long m = s_row[i] & ~colx;
              if (WEIGHTED) {
                double rs = 0.0;
                while (m) { const int cc = __ffsll((long long)m) - 1; m &= m - 1ull; rs += s_w[i * nov + cc] * (double)d_c[cc * APS_THREADS]; }
                if (rs == 0.0) dead = true; else d_r[i * APS_THREADS] = (float)(1.0 / rs);
              } else {
                float rs = 0.0f;
                while (m) { const int cc = __ffsll((long long)m) - 1; m &= m - 1ull; rs += d_c[cc * APS_THREADS]; }
                if (rs == 0.0f) dead = true; else d_r[i * APS_THREADS] = 1.0f / rs;
              }
            }
          }
        }
        if (!dead) {
          const float dr = d_r[row * APS_THREADS];
          double tot = 0.0;
          unsigned long long m = avail;
          while (m) { const int cc = __ffsll((long long)m) - 1; m &= m - 1ull; tot += (double)(dr * d_c[cc * APS_THREADS]); }
          if (tot == 0.0) {
            dead = true;
          } else {
            const double target = ((double)draw + 1.0) * (1.0 / 4294967296.0) * tot;
            double run = 0.0;
            bool picked = false;
            m = avail;
            while (m) {
              const int cc = __ffsll((long long)m) - 1; m &= m - 1ull;
              const double sv = (double)(dr * d_c[cc * APS_THREADS]);
              run += sv;
              if (target <= run) { col = cc; perm /= (sv / tot); picked = true; break; }
            }
            if (!picked) dead = true;
          }
        }
      }
    }
    if (!dead) { rowx |= 1ull << row; colx |= 1ull << col; ++step; }
    if (dead || step == nov) {
      // a dead end estimates 0; `perm` is still the product over the `step` steps it completed
      if (a.trace) {
        a.trace[2 * (trial - a.trial_lo)] = (double)step;
        a.trace[2 * (trial - a.trial_lo) + 1] = perm;
      }
      a.est[tix] = dead ? 0.0 : perm;
      tix = atomicAdd(a.queue, 1u);
      trial = a.trial_lo + tix;
      step = 0;
    }
  }
}

// ---- thread-per-trial Rasmussen for 64 < nov (sparse patterns, e.g. the 36x36 grid: nov = 648) ----
// One thread runs a trial; its state lives in shared memory as 32-bit words laid out [word][thread]
// (every lane stays in its own bank):
//   deg   one BITS-wide field per row: remaining column count, all ones once the row is extracted
//         (BITS = 4 when no row has more than 14 entries -- grids have 4 -- else 8)
//   gmin  one byte per group of 32 rows: the minimum of the group's deg fields
//   colx  one bit per column: extracted
// The minimum-degree row (first in ascending order, as gpu_approximation_sparse.cu:242-256 scans) is
// found through gmin: SIMD byte minimum over the ngroups bytes, first group holding it, first row of
// that group holding it -- about 100 instructions per step instead of a scan of every CRS row.
// The pattern (CRS + CCS, 16-bit) is shared by the block.  With 4-bit degrees a trial's state on the
// 36 x 36 grid is 432 bytes instead of 756, so two blocks of 224 threads share an SM (14 warps instead of 8).
// Persistent lanes pulling trials from a global counter.  Per-trial values are bit-identical to the other
// engines and to the oracle.
#define APM_MAX_THREADS 256

struct MidArgs {
  const int* rptrs; const int* cols; const int* cptrs; const int* rows;
  double* est;
  unsigned int* queue;
  double* trace;
  unsigned long long trial_lo, trial_hi;
  unsigned long long seed;
  double sq_scale;
  int nov, nnz;
};

__device__ __forceinline__ unsigned byte_min4(unsigned w) {   // minimum of the four bytes of w
  unsigned m = __vminu4(w, w >> 16);
  m = __vminu4(m, m >> 8);
  return m & 0xffu;
}

// BITS-wide unsigned fields packed in 32-bit words
template <int BITS> struct Fields;
template <> struct Fields<8> {
  static constexpr unsigned EX = 0xffu;
  // running field-wise minimum accumulator over several words, and its final value
  __device__ static void acc_init(unsigned& a0, unsigned& a1) { a0 = 0xffffffffu; a1 = 0xffffffffu; }
  __device__ static void acc_add(unsigned& a0, unsigned& a1, unsigned w) { a0 = __vminu4(a0, w); (void)a1; }
  __device__ static unsigned acc_min(unsigned a0, unsigned a1) { (void)a1; return byte_min4(a0); }
  // index of the first field of w equal to v, or -1
  __device__ static int first_eq(unsigned w, unsigned v) {
    const unsigned eq = __vcmpeq4(w, v * 0x01010101u);
    return eq ? ((__ffs(eq) - 1) >> 3) : -1;
  }
};
template <> struct Fields<4> {
  static constexpr unsigned EX = 0xfu;
  __device__ static void acc_init(unsigned& a0, unsigned& a1) { a0 = 0xffffffffu; a1 = 0xffffffffu; }
  __device__ static void acc_add(unsigned& a0, unsigned& a1, unsigned w) {
    a0 = __vminu4(a0, w & 0x0f0f0f0fu);              // even fields
    a1 = __vminu4(a1, (w >> 4) & 0x0f0f0f0fu);       // odd fields
  }
  __device__ static unsigned acc_min(unsigned a0, unsigned a1) { return byte_min4(__vminu4(a0, a1)); }
  __device__ static int first_eq(unsigned w, unsigned v) {
    const unsigned pat = v * 0x01010101u;
    const unsigned m = (__vcmpeq4(w & 0x0f0f0f0fu, pat) & 0x0f0f0f0fu) | (__vcmpeq4((w >> 4) & 0x0f0f0f0fu, pat) & 0xf0f0f0f0u);
    return m ? ((__ffs(m) - 1) >> 2) : -1;
  }
};

template <int BITS>
__global__ void __launch_bounds__(APM_MAX_THREADS)
rasmussen_mid_kernel(const MidArgs a) {
  using F = Fields<BITS>;
  constexpr int RPW = 32 / BITS;                 // rows per deg word
  constexpr int GW = 32 / RPW;                   // deg words per group of 32 rows
  extern __shared__ __align__(16) unsigned char smraw[];
  const int nov = a.nov, nnz = a.nnz;
  const int T = blockDim.x;
  const int WD = (nov + RPW - 1) / RPW;          // deg words
  const int NG = (nov + 31) >> 5;                // groups of 32 rows
  const int WG = (NG + 3) >> 2;                  // gmin words
  const int WC = (nov + 31) >> 5;                // colx words
  unsigned short* s_rptrs = reinterpret_cast<unsigned short*>(smraw);
  unsigned short* s_cptrs = s_rptrs + (nov + 1);
  unsigned short* s_cols = s_cptrs + (nov + 1);
  unsigned short* s_rows = s_cols + nnz;
  unsigned* s_deg0 = reinterpret_cast<unsigned*>(smraw + (((size_t)(2 * (nov + 1) + 2 * nnz) * 2 + 15) & ~(size_t)15));   // initial deg words
  unsigned* s_gmin0 = s_deg0 + WD;                                    // initial gmin words
  unsigned* st = s_gmin0 + WG;                                        // per-thread state, [word][thread]
  unsigned* deg = st + threadIdx.x;
  unsigned* gmin = deg + (size_t)WD * T;
  unsigned* colx = gmin + (size_t)WG * T;

  for (int e = threadIdx.x; e <= nov; e += T) { s_rptrs[e] = (unsigned short)a.rptrs[e]; s_cptrs[e] = (unsigned short)a.cptrs[e]; }
  for (int e = threadIdx.x; e < nnz; e += T) { s_cols[e] = (unsigned short)a.cols[e]; s_rows[e] = (unsigned short)a.rows[e]; }
  __syncthreads();
  for (int w = threadIdx.x; w < WD; w += T) {
    unsigned v = 0;
    for (int b = 0; b < RPW; ++b) {
      const int r = RPW * w + b;
      const unsigned d = (r < nov) ? (unsigned)(s_rptrs[r + 1] - s_rptrs[r]) : F::EX;   // padding rows: extracted
      v |= d << (BITS * b);
    }
    s_deg0[w] = v;
  }
  __syncthreads();
  for (int w = threadIdx.x; w < WG; w += T) {
    unsigned v = 0;
    for (int b = 0; b < 4; ++b) {
      const int g = 4 * w + b;
      unsigned m = 0xffu;
      if (g < NG) {
        unsigned a0, a1;
        F::acc_init(a0, a1);
        for (int q = 0; q < GW; ++q) { const int wi = GW * g + q; if (wi < WD) F::acc_add(a0, a1, s_deg0[wi]); }
        m = F::acc_min(a0, a1);
        if (m == F::EX) m = 0xffu;                 // a group without rows
      }
      v |= m << (8 * b);
    }
    s_gmin0[w] = v;
  }
  __syncthreads();

  const unsigned int n_trials = (unsigned int)(a.trial_hi - a.trial_lo);
  unsigned int tix = atomicAdd(a.queue, 1u);
  unsigned long long trial = a.trial_lo + tix;
  const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
  double perm = 1.0;
  uint32_t rnd0 = 0, rnd1 = 0, rnd2 = 0, rnd3 = 0;
  int step = 0;

  for (;;) {
    const bool have = tix < n_trials;
    if (!__any_sync(0xffffffffu, have)) break;
    if (!have) continue;
    if (step == 0) {
      perm = 1.0;
      for (int w = 0; w < WD; ++w) deg[w * T] = s_deg0[w];
      for (int w = 0; w < WG; ++w) gmin[w * T] = s_gmin0[w];
      for (int w = 0; w < WC; ++w) colx[w * T] = 0u;
    }
    if ((step & 3) == 0) {
      uint32_t r[4];
      philox4x32_10((uint32_t)trial, (uint32_t)(trial >> 32), (uint32_t)(step >> 2), 0u, k0, k1, r);
      rnd0 = r[0]; rnd1 = r[1]; rnd2 = r[2]; rnd3 = r[3];
    }
    const int rw = step & 3;
    const uint32_t draw = (rw == 0) ? rnd0 : (rw == 1) ? rnd1 : (rw == 2) ? rnd2 : rnd3;
    // ---- minimum remaining degree, first group and first row holding it ----
    unsigned acc4 = 0xffffffffu;
    for (int w = 0; w < WG; ++w) acc4 = __vminu4(acc4, gmin[w * T]);
    const unsigned dmin = byte_min4(acc4);
    bool dead = (dmin == 0u);
    if (!dead) {
      int g = 0;
      for (int w = 0; w < WG; ++w) {
        const int f = Fields<8>::first_eq(gmin[w * T], dmin);
        if (f >= 0) { g = 4 * w + f; break; }
      }
      int row = 0;
      for (int q = 0; q < GW; ++q) {
        const int wi = GW * g + q;
        if (wi >= WD) break;
        const int f = F::first_eq(deg[wi * T], dmin);
        if (f >= 0) { row = RPW * wi + f; break; }
      }
      perm *= (double)dmin;
      int want = (int)(((uint64_t)draw * (uint64_t)dmin) >> 32);
      int col = -1;
      for (int t = s_rptrs[row]; t < s_rptrs[row + 1]; ++t) {
        const int c = s_cols[t];
        if ((colx[(c >> 5) * T] >> (c & 31)) & 1u) continue;
        if (want == 0) { col = c; break; }
        --want;
      }
      // ---- extract row and column ----
      colx[(col >> 5) * T] |= 1u << (col & 31);
      deg[(row / RPW) * T] |= F::EX << (BITS * (row % RPW));
      {
        unsigned a0, a1;
        F::acc_init(a0, a1);
        const int gb = row >> 5;
        for (int q = 0; q < GW; ++q) { const int wi = GW * gb + q; if (wi < WD) F::acc_add(a0, a1, deg[wi * T]); }
        unsigned gm = F::acc_min(a0, a1);
        if (gm == F::EX) gm = 0xffu;               // the whole group is extracted
        unsigned gw = gmin[(gb >> 2) * T];
        gw = (gw & ~(0xffu << (8 * (gb & 3)))) | (gm << (8 * (gb & 3)));
        gmin[(gb >> 2) * T] = gw;
      }
      for (int t = s_cptrs[col]; t < s_cptrs[col + 1]; ++t) {
        const int r2 = s_rows[t];
        const unsigned dw = deg[(r2 / RPW) * T];
        const unsigned d = (dw >> (BITS * (r2 % RPW))) & F::EX;
        if (d == F::EX) continue;                                      // extracted
        deg[(r2 / RPW) * T] = dw - (1u << (BITS * (r2 % RPW)));
        const int g2 = r2 >> 5;
        const unsigned gw = gmin[(g2 >> 2) * T];
        const unsigned cur = (gw >> (8 * (g2 & 3))) & 0xffu;
        if (d - 1u < cur) gmin[(g2 >> 2) * T] = (gw & ~(0xffu << (8 * (g2 & 3)))) | ((d - 1u) << (8 * (g2 & 3)));
      }
    }
    if (!dead) ++step;
    if (dead || step == nov) {
      // a dead end estimates 0; `perm` is still the product over the `step` steps it completed
      if (a.trace) {
        a.trace[2 * (trial - a.trial_lo)] = (double)step;
        a.trace[2 * (trial - a.trial_lo) + 1] = perm;
      }
      a.est[tix] = dead ? 0.0 : perm;
      tix = atomicAdd(a.queue, 1u);
      trial = a.trial_lo + tix;
      step = 0;
    }
  }
}

}  // namespace spb

using namespace spb;

struct spd_approx_plan {
  Lane* lanep = nullptr;
  int nov = 0, nnz = 0;
  int scaling = 0, y = 4, z = 5;
  bool weighted = false;
  unsigned long long seed = 0;
  double sq_scale = 1.0;
  int *d_rptrs = nullptr, *d_cols = nullptr, *d_cptrs = nullptr, *d_rows = nullptr;
  double *d_rvals = nullptr, *d_cvals = nullptr;
  size_t smem_bytes = 0;
  int blocks = 0;
  int ellW = 0;                                  // 0: CSR sweeps; 4 or 8: ELL width (pattern-only scaling)
  bool wide = false;                             // some row has more than 255 entries: 16-bit degrees
  unsigned short *d_ell_rows = nullptr, *d_ell_cols = nullptr;
  // thread-per-trial engine (nov <= 64)
  bool small = false;
  unsigned long long *d_rowmask = nullptr, *d_colmask = nullptr;
  double* d_wdense = nullptr;
  size_t small_smem = 0;
  int small_blocks = 0;
  // thread-per-trial Rasmussen for larger sparse patterns
  bool mid = false;
  int mid_bits = 8, mid_threads = APM_MAX_THREADS;
  size_t mid_smem = 0;
  int mid_blocks = 0;
  double* d_trace = nullptr;                     // set by spd_approx_plan_trace for the duration of one run
  bool pending = false;
  spd_run_info info;
};

typedef void (*warp_kernel_t)(const ApproxArgs);
static warp_kernel_t warp_kernel_of(const spd_approx_plan* p) {
  if (p->weighted) return p->wide ? approx_kernel<true, 0, true> : approx_kernel<true, 0, false>;
  if (p->ellW == 4) return approx_kernel<false, 4, false>;
  if (p->ellW == 8) return approx_kernel<false, 8, false>;
  return p->wide ? approx_kernel<false, 0, true> : approx_kernel<false, 0, false>;
}

template <typename K>
static int small_prepare(spd_approx_plan* p, K kern) {
  cudaError_t e;
  if (p->small_smem > 40 * 1024) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SPB_SMEM_OPTIN_BYTES);
    if (e != cudaSuccess) { set_error("smem opt-in (%zu B): %s", p->small_smem, cudaGetErrorString(e)); return SPD_ECUDA; }
  }
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, APS_THREADS, p->small_smem);
  if (e != cudaSuccess || per_sm < 1) { set_error("small approx kernel does not fit on an SM"); return SPD_ECUDA; }
  p->small_blocks = per_sm * p->lanep->sm_count;
  return SPD_OK;
}

extern "C" {

int spd_approx_plan_create(int device, const int* rptrs, const int* cols, const int* cptrs,
                           const int* rows, const double* rvals, const double* cvals, int nov, int nnz,
                           int scaling, int scale_intervals, int scale_times, unsigned long long seed,
                           spd_approx_plan** out) {
  if (!rptrs || !cols || !cptrs || !rows || !out) { set_error("null argument"); return SPD_EINVAL; }
  if (nov < 1 || nov > 65535) { set_error("approximation supports 1 <= n <= 65535 (got %d)", nov); return SPD_ELIMIT; }
  if (nnz < 0) { set_error("negative nnz"); return SPD_EINVAL; }
  if (scaling && (scale_intervals < 1 || scale_times < 0)) { set_error("bad scaling parameters"); return SPD_EINVAL; }
  // the pattern comes straight from the caller over the C ABI: one pass of validation, so that a bad index
  // cannot become an out-of-range shift or shared-memory access in the kernels
  if (rptrs[0] != 0 || cptrs[0] != 0 || rptrs[nov] != nnz || cptrs[nov] != nnz) {
    set_error("CRS/CCS pointers do not span nnz = %d (rptrs[0]=%d rptrs[n]=%d cptrs[0]=%d cptrs[n]=%d)", nnz, rptrs[0],
              rptrs[nov], cptrs[0], cptrs[nov]);
    return SPD_EINVAL;
  }
  for (int r = 0; r < nov; ++r) {
    if (rptrs[r + 1] < rptrs[r] || cptrs[r + 1] < cptrs[r]) { set_error("CRS/CCS pointers are not monotone at %d", r); return SPD_EINVAL; }
  }
  for (int t = 0; t < nnz; ++t) {
    if (cols[t] < 0 || cols[t] >= nov || rows[t] < 0 || rows[t] >= nov) {
      set_error("CRS/CCS index out of range at entry %d (col %d, row %d, n = %d)", t, cols[t], rows[t], nov);
      return SPD_EINVAL;
    }
  }
  spd_approx_plan* p = new (std::nothrow) spd_approx_plan();
  if (!p) return SPD_ENOMEM;
  int rc = lane_acquire(device, &p->lanep);
  if (rc != SPD_OK) { delete p; return rc; }
  Lane& L = *p->lanep;
  auto fail = [&](int code) { lane_release(p->lanep); delete p; return code; };
  p->nov = nov; p->nnz = nnz; p->scaling = scaling ? 1 : 0;
  p->y = scale_intervals; p->z = scale_times; p->seed = seed;
  p->weighted = (rvals != nullptr && cvals != nullptr && scaling);
  // estimates of a large pattern are ~1e159 (36x36 grid): square them in a scaled domain
  p->sq_scale = (nov > 64) ? ldexp(1.0, -4 * nov / 5) : 1.0;
  const int words = (nov + 31) / 32;
  for (int i = 0; i < nov; ++i)
    if (rptrs[i + 1] - rptrs[i] > 255) p->wide = true;
  // ELL form of the pattern for the Sinkhorn sweeps when every row and column has <= 8 entries
  if (p->scaling && !p->weighted && env_int("SP_APPROX_ELL", 1) != 0) {
    int maxdeg = 0;
    for (int i = 0; i < nov; ++i) {
      maxdeg = std::max(maxdeg, rptrs[i + 1] - rptrs[i]);
      maxdeg = std::max(maxdeg, cptrs[i + 1] - cptrs[i]);
    }
    if (maxdeg <= 8) p->ellW = maxdeg <= 4 ? 4 : 8;
  }
  // the ELL arrays replace CRS + CCS in shared memory when they are used
  size_t off = p->ellW ? 0 : (size_t)(2 * (nov + 1) + 2 * nnz) * sizeof(int);
  off = (off + 15) & ~(size_t)15;
  off += 2 * (size_t)nov * p->ellW * sizeof(unsigned short);
  off = (off + 15) & ~(size_t)15;
  const size_t deg_bytes = ((p->wide ? 2 : 1) * (size_t)nov + 15) & ~(size_t)15;
  size_t warp_bytes = deg_bytes + 2 * (size_t)words * 4 + (scaling ? 2 * (size_t)(nov + 1) * 4 : 0);
  warp_bytes = (warp_bytes + 15) & ~(size_t)15;
  p->smem_bytes = off + APX_WARPS * warp_bytes;
  if (p->smem_bytes > SPB_SMEM_OPTIN_BYTES) {
    set_error("pattern too large for shared memory (%zu B needed)", p->smem_bytes);
    return fail(SPD_ELIMIT);
  }
  cudaError_t e;
  if ((e = cudaSetDevice(device)) != cudaSuccess) { set_error("%s", cudaGetErrorString(e)); return fail(SPD_ECUDA); }
  const size_t ib = (size_t)(nov + 1) * sizeof(int), nb = (size_t)(nnz > 0 ? nnz : 1) * sizeof(int);
  if ((rc = lane_arena_alloc(&L, ib, (void**)&p->d_rptrs)) != SPD_OK) return fail(rc);
  if ((rc = lane_arena_alloc(&L, ib, (void**)&p->d_cptrs)) != SPD_OK) return fail(rc);
  if ((rc = lane_arena_alloc(&L, nb, (void**)&p->d_cols)) != SPD_OK) return fail(rc);
  if ((rc = lane_arena_alloc(&L, nb, (void**)&p->d_rows)) != SPD_OK) return fail(rc);
  if (p->weighted) {
    if ((rc = lane_arena_alloc(&L, (size_t)(nnz > 0 ? nnz : 1) * 8, (void**)&p->d_rvals)) != SPD_OK) return fail(rc);
    if ((rc = lane_arena_alloc(&L, (size_t)(nnz > 0 ? nnz : 1) * 8, (void**)&p->d_cvals)) != SPD_OK) return fail(rc);
  }
  if ((e = cudaMemcpyAsync(p->d_rptrs, rptrs, ib, cudaMemcpyHostToDevice, L.stream)) != cudaSuccess ||
      (e = cudaMemcpyAsync(p->d_cptrs, cptrs, ib, cudaMemcpyHostToDevice, L.stream)) != cudaSuccess ||
      (nnz > 0 && (e = cudaMemcpyAsync(p->d_cols, cols, (size_t)nnz * 4, cudaMemcpyHostToDevice, L.stream)) != cudaSuccess) ||
      (nnz > 0 && (e = cudaMemcpyAsync(p->d_rows, rows, (size_t)nnz * 4, cudaMemcpyHostToDevice, L.stream)) != cudaSuccess) ||
      (p->weighted && nnz > 0 && (e = cudaMemcpyAsync(p->d_rvals, rvals, (size_t)nnz * 8, cudaMemcpyHostToDevice, L.stream)) != cudaSuccess) ||
      (p->weighted && nnz > 0 && (e = cudaMemcpyAsync(p->d_cvals, cvals, (size_t)nnz * 8, cudaMemcpyHostToDevice, L.stream)) != cudaSuccess) ||
      (e = cudaStreamSynchronize(L.stream)) != cudaSuccess) {
    set_error("approx plan upload: %s", cudaGetErrorString(e));
    return fail(SPD_ECUDA);
  }
  if (p->ellW) {
    const int W = p->ellW;
    std::vector<unsigned short> er((size_t)nov * W, (unsigned short)nov), ec((size_t)nov * W, (unsigned short)nov);
    for (int i = 0; i < nov; ++i) {
      for (int t = cptrs[i]; t < cptrs[i + 1]; ++t) er[(size_t)i * W + (t - cptrs[i])] = (unsigned short)rows[t];
      for (int t = rptrs[i]; t < rptrs[i + 1]; ++t) ec[(size_t)i * W + (t - rptrs[i])] = (unsigned short)cols[t];
    }
    const size_t eb = er.size() * sizeof(unsigned short);
    if ((rc = lane_arena_alloc(&L, eb, (void**)&p->d_ell_rows)) != SPD_OK) return fail(rc);
    if ((rc = lane_arena_alloc(&L, eb, (void**)&p->d_ell_cols)) != SPD_OK) return fail(rc);
    if ((e = cudaMemcpyAsync(p->d_ell_rows, er.data(), eb, cudaMemcpyHostToDevice, L.stream)) != cudaSuccess ||
        (e = cudaMemcpyAsync(p->d_ell_cols, ec.data(), eb, cudaMemcpyHostToDevice, L.stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(L.stream)) != cudaSuccess) {
      set_error("approx plan upload: %s", cudaGetErrorString(e));
      return fail(SPD_ECUDA);
    }
  }
  if (p->smem_bytes > 40 * 1024) {
    e = cudaFuncSetAttribute(warp_kernel_of(p), cudaFuncAttributeMaxDynamicSharedMemorySize, SPB_SMEM_OPTIN_BYTES);
    if (e != cudaSuccess) { set_error("smem opt-in (%zu B): %s", p->smem_bytes, cudaGetErrorString(e)); return fail(SPD_ECUDA); }
  }
  // ---- thread-per-trial engine for nov <= 64: 64-bit pattern words ----
  if (nov <= 64 && env_int("SP_APPROX_FORCE_WARP", 0) == 0) {
    unsigned long long rm[64], cm[64];
    for (int i = 0; i < nov; ++i) {
      rm[i] = 0ull; cm[i] = 0ull;
      for (int t = rptrs[i]; t < rptrs[i + 1]; ++t) rm[i] |= 1ull << cols[t];
      for (int t = cptrs[i]; t < cptrs[i + 1]; ++t) cm[i] |= 1ull << rows[t];
    }
    if ((rc = lane_arena_alloc(&L, 64 * 8, (void**)&p->d_rowmask)) != SPD_OK) return fail(rc);
    if ((rc = lane_arena_alloc(&L, 64 * 8, (void**)&p->d_colmask)) != SPD_OK) return fail(rc);
    std::vector<double> wd;
    if (p->weighted) {
      wd.assign((size_t)nov * nov, 0.0);
      for (int i = 0; i < nov; ++i)
        for (int t = rptrs[i]; t < rptrs[i + 1]; ++t) wd[(size_t)i * nov + cols[t]] = rvals[t];
      if ((rc = lane_arena_alloc(&L, wd.size() * 8, (void**)&p->d_wdense)) != SPD_OK) return fail(rc);
    }
    if ((e = cudaMemcpyAsync(p->d_rowmask, rm, (size_t)nov * 8, cudaMemcpyHostToDevice, L.stream)) != cudaSuccess ||
        (e = cudaMemcpyAsync(p->d_colmask, cm, (size_t)nov * 8, cudaMemcpyHostToDevice, L.stream)) != cudaSuccess ||
        (p->weighted && (e = cudaMemcpyAsync(p->d_wdense, wd.data(), wd.size() * 8, cudaMemcpyHostToDevice, L.stream)) != cudaSuccess) ||
        (e = cudaStreamSynchronize(L.stream)) != cudaSuccess) {
      set_error("approx plan upload: %s", cudaGetErrorString(e));
      return fail(SPD_ECUDA);
    }
    p->small_smem = (size_t)2 * nov * 8 + (p->weighted ? (size_t)nov * nov * 8 : 0) +
                    (scaling ? (size_t)2 * nov * APS_THREADS * 4 : 0);
    if (!scaling) rc = small_prepare(p, approx_small_kernel<false, false>);
    else if (p->weighted) rc = small_prepare(p, approx_small_kernel<true, true>);
    else rc = small_prepare(p, approx_small_kernel<true, false>);
    if (rc != SPD_OK) return fail(rc);
    p->small = true;

  }
  // ---- thread-per-trial Rasmussen for nov > 64 when the per-thread state fits shared memory ----
  int maxdeg = 0;
  for (int r = 0; r < nov; ++r) maxdeg = (rptrs[r + 1] - rptrs[r] > maxdeg) ? rptrs[r + 1] - rptrs[r] : maxdeg;
  if (!p->small && !scaling && maxdeg <= 254 && nnz < 65536 && env_int("SP_APPROX_FORCE_WARP", 0) == 0) {
    // 4-bit degrees when they fit (no row with more than 14 entries), and the largest block of whole warps
    // that lets two blocks share an SM (else one block of 256 threads, if that fits at all)
    p->mid_bits = (maxdeg <= 14 && env_int("SP_APPROX_MID_BITS", 4) == 4) ? 4 : 8;
    const int RPW = 32 / p->mid_bits;
    const int WD = (nov + RPW - 1) / RPW, NG = (nov + 31) / 32, WG = (NG + 3) / 4, WC = (nov + 31) / 32;
    const size_t shared_part = (((size_t)(2 * (nov + 1) + 2 * nnz) * 2 + 15) & ~(size_t)15) + (size_t)(WD + WG) * 4;
    const size_t per_thread = (size_t)(WD + WG + WC) * 4;
    const size_t half_sm = (227 * 1024) / 2 - 2048;
    int threads = 0;
    if (shared_part < half_sm) threads = (int)((half_sm - shared_part) / per_thread) / 32 * 32;
    if (threads > APM_MAX_THREADS) threads = APM_MAX_THREADS;
    if (threads < 128) threads = APM_MAX_THREADS;                 // two blocks would be too thin: one full block
    p->mid_threads = threads;
    p->mid_smem = shared_part + per_thread * threads;
    if (p->mid_smem <= 220 * 1024) {
      const void* kern = (p->mid_bits == 4) ? (const void*)rasmussen_mid_kernel<4> : (const void*)rasmussen_mid_kernel<8>;
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SPB_SMEM_OPTIN_BYTES);
      int per = 0;
      if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kern, threads, p->mid_smem);
      if (e == cudaSuccess && per >= 1) {
        p->mid = true;
        p->mid_blocks = per * L.sm_count;
      } else {
        (void)cudaGetLastError();
      }
    }
  }
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, warp_kernel_of(p), APX_THREADS, p->smem_bytes);
  if (e != cudaSuccess || per_sm < 1) { set_error("approx kernel does not fit on an SM: %s", cudaGetErrorString(e)); return fail(SPD_ECUDA); }
  p->blocks = per_sm * L.sm_count;     // persistent grid: resident blocks x SM count
  if ((rc = lane_reserve_partials(&L, (size_t)1 << 17)) != SPD_OK) return fail(rc);    // estimate slots; grows on demand
  *out = p;
  return SPD_OK;
}

void spd_approx_plan_destroy(spd_approx_plan* p) {
  if (!p) return;
  lane_release(p->lanep);
  delete p;
}

int spd_approx_plan_launch(spd_approx_plan* p, unsigned long long lo, unsigned long long hi) {
  if (!p) { set_error("null plan"); return SPD_EINVAL; }
  if (p->pending) { set_error("plan already has a pending run"); return SPD_EINVAL; }
  if (hi < lo) { set_error("bad trial range"); return SPD_EINVAL; }
  Lane& L = *p->lanep;
  SPB_CUDA(cudaSetDevice(L.device));
  memset(&p->info, 0, sizeof(p->info));
  p->info.units = hi - lo;
  p->info.path = p->scaling ? SPD_PATH_SCALING : SPD_PATH_RASMUSSEN;
  SPB_CUDA(cudaEventRecord(L.ev0, L.stream));
  // one slot per estimate (at most 2^22 trials = 32 MiB per launch), a global trial counter, and a
  // fixed-order reduction of the slots into {sum, sum of scaled squares, survivors}
  const unsigned long long max_trials = 1ull << 22;
  int launches = 0, rc;
  bool first = true;
  unsigned long long t0 = lo;
  if (hi == lo) {
    if ((rc = launch_reduce_estimates(L, L.d_partials, 0, p->sq_scale, L.d_result, false)) != SPD_OK) return rc;
    ++launches;
  }
  while (t0 < hi) {
    const unsigned long long cnt = (hi - t0 < max_trials) ? hi - t0 : max_trials;
    if ((rc = lane_reserve_partials(&L, (size_t)cnt)) != SPD_OK) return rc;
    SPB_CUDA(cudaMemsetAsync(L.d_queue, 0, sizeof(unsigned int), L.stream));
    double* d_trace = p->d_trace ? p->d_trace + 2 * (t0 - lo) : nullptr;
    if (p->small) {
      int blocks = p->small_blocks;
      const unsigned long long need = (cnt + APS_THREADS - 1) / APS_THREADS;
      if (need < (unsigned long long)blocks) blocks = (int)need;
      SmallArgs sa;
      sa.rowmask = p->d_rowmask; sa.colmask = p->d_colmask; sa.wdense = p->d_wdense;
      sa.est = L.d_partials; sa.queue = L.d_queue; sa.trace = d_trace;
      sa.trial_lo = t0; sa.trial_hi = t0 + cnt; sa.seed = p->seed; sa.sq_scale = p->sq_scale;
      sa.nov = p->nov; sa.scaling = p->scaling; sa.scale_intervals = p->y; sa.scale_times = p->z;
      if (!p->scaling) approx_small_kernel<false, false><<<blocks, APS_THREADS, p->small_smem, L.stream>>>(sa);
      else if (p->weighted) approx_small_kernel<true, true><<<blocks, APS_THREADS, p->small_smem, L.stream>>>(sa);
      else approx_small_kernel<true, false><<<blocks, APS_THREADS, p->small_smem, L.stream>>>(sa);
    } else if (p->mid) {
      int blocks = p->mid_blocks;
      const unsigned long long need = (cnt + p->mid_threads - 1) / p->mid_threads;
      if (need < (unsigned long long)blocks) blocks = (int)need;
      MidArgs ma;
      ma.rptrs = p->d_rptrs; ma.cols = p->d_cols; ma.cptrs = p->d_cptrs; ma.rows = p->d_rows;
      ma.est = L.d_partials; ma.queue = L.d_queue; ma.trace = d_trace;
      ma.trial_lo = t0; ma.trial_hi = t0 + cnt; ma.seed = p->seed; ma.sq_scale = p->sq_scale;
      ma.nov = p->nov; ma.nnz = p->nnz;
      if (p->mid_bits == 4) rasmussen_mid_kernel<4><<<blocks, p->mid_threads, p->mid_smem, L.stream>>>(ma);
      else rasmussen_mid_kernel<8><<<blocks, p->mid_threads, p->mid_smem, L.stream>>>(ma);
    } else {
      int blocks = p->blocks;
      const unsigned long long need = (cnt + APX_WARPS - 1) / APX_WARPS;
      if (need < (unsigned long long)blocks) blocks = (int)need;
      ApproxArgs a;
      a.rptrs = p->d_rptrs; a.cols = p->d_cols; a.cptrs = p->d_cptrs; a.rows = p->d_rows;
      a.rvals = p->d_rvals; a.cvals = p->d_cvals;
      a.ell_rows = p->d_ell_rows; a.ell_cols = p->d_ell_cols;
      a.est = L.d_partials; a.queue = L.d_queue; a.trace = d_trace;
      a.trial_lo = t0; a.trial_hi = t0 + cnt; a.seed = p->seed; a.sq_scale = p->sq_scale;
      a.nov = p->nov; a.nnz = p->nnz; a.scaling = p->scaling;
      a.scale_intervals = p->y; a.scale_times = p->z;
      warp_kernel_of(p)<<<blocks, APX_THREADS, p->smem_bytes, L.stream>>>(a);
    }
    SPB_CUDA(cudaGetLastError());
    if ((rc = launch_reduce_estimates(L, L.d_partials, (size_t)cnt, p->sq_scale, L.d_result, !first)) != SPD_OK) return rc;
    launches += 2;
    first = false;
    t0 += cnt;
  }
  SPB_CUDA(cudaMemcpyAsync(L.h_result, L.d_result, 3 * sizeof(double), cudaMemcpyDeviceToHost, L.stream));
  SPB_CUDA(cudaEventRecord(L.ev1, L.stream));
  p->info.launches = launches;
  p->pending = true;
  return SPD_OK;
}

int spd_approx_plan_wait(spd_approx_plan* p, double* sum, spd_run_info* info) {
  if (!p || !p->pending) { set_error("no pending run"); return SPD_EINVAL; }
  p->pending = false;
  SPB_CUDA(cudaSetDevice(p->lanep->device));
  SPB_CUDA(cudaEventSynchronize(p->lanep->ev1));
  float ms = 0.f;
  SPB_CUDA(cudaEventElapsedTime(&ms, p->lanep->ev0, p->lanep->ev1));
  p->info.kernel_ms = ms;
  p->info.aux0 = p->lanep->h_result[1];   // sum of (estimate * sq_scale)^2
  p->info.aux1 = p->sq_scale;
  memcpy(&p->info.visited, &p->lanep->h_result[2], sizeof(unsigned long long));   // trials that reached the last step
  if (sum) *sum = p->lanep->h_result[0];
  if (info) *info = p->info;
  return SPD_OK;
}

int spd_approx_plan_run(spd_approx_plan* p, unsigned long long lo, unsigned long long hi, double* sum,
                        spd_run_info* info) {
  int rc = spd_approx_plan_launch(p, lo, hi);
  if (rc != SPD_OK) return rc;
  return spd_approx_plan_wait(p, sum, info);
}

// Trials [lo, hi) in one launch with their per-trial record, for bit-exact comparison with the oracle
// (tests; <= 2^16 trials): estimate[i] (0 for a dead end), steps[i] completed and the running product at
// that point (== estimate when the trial reached the last step).  A dead trial still has to agree with the
// CPU restatement in how far it got and in the product it had built -- on patterns where nearly every trial dies
// (the 36 x 36 grid) this is what makes the comparison say something.
int spd_approx_plan_trace(spd_approx_plan* p, unsigned long long lo, unsigned long long hi, double* estimate,
                          int* steps, double* partial) {
  if (!p || hi < lo || hi - lo > 65536ull) { set_error("bad trace range"); return SPD_EINVAL; }
  const size_t cnt = (size_t)(hi - lo);
  if (cnt == 0) return SPD_OK;
  Lane& L = *p->lanep;
  SPB_CUDA(cudaSetDevice(L.device));
  double* d_trace = nullptr;
  SPB_CUDA(cudaMalloc(&d_trace, 2 * cnt * sizeof(double)));
  p->d_trace = d_trace;
  double s = 0.0;
  spd_run_info info;
  int rc = spd_approx_plan_run(p, lo, hi, &s, &info);
  p->d_trace = nullptr;
  std::vector<double> h(2 * cnt);
  if (rc == SPD_OK && cudaMemcpy(h.data(), d_trace, 2 * cnt * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) {
    set_error("trace read-back failed");
    rc = SPD_ECUDA;
  }
  cudaFree(d_trace);
  if (rc != SPD_OK) return rc;
  for (size_t i = 0; i < cnt; ++i) {
    const int st = (int)h[2 * i];
    if (steps) steps[i] = st;
    if (partial) partial[i] = h[2 * i + 1];
    if (estimate) estimate[i] = (st == p->nov) ? h[2 * i + 1] : 0.0;
  }
  return SPD_OK;
}

// One trial on the device (kept for callers that want a single estimate).
int spd_approx_plan_trial(spd_approx_plan* p, unsigned long long trial, double* value) {
  return spd_approx_plan_trace(p, trial, trial + 1, value, nullptr, nullptr);
}

}  // extern "C"
