/* oracle_approx.c -- CPU restatement of SUPerman's Monte-Carlo estimators.  TEST INFRASTRUCTURE
 * ONLY (see the header of oracle.c for who may load liboracle.so).
 *
 * One trial, sequential, written the way the reference kernels are written (rescan every
 * remaining row's CRS entries at every step; bit masks; no incremental degrees), so that it is an
 * independent check of the warp-cooperative CUDA kernels:
 *   Rasmussen : kernel_rasmussen_sparse        gpu_approximation_sparse.cu:238-287
 *   scaling   : kernel_approximation_sparse    gpu_approximation_sparse.cu:342-449 (float sums)
 *               kernel_approximation (dense)   gpu_approximation_dense.cu:262-365 (weighted, double sums)
 * The reference's RNG is cuRAND XORWOW seeded from time(0) (unpinned); both the
 * product and this oracle use Philox4x32-10 (Salmon et al., SC'11: 10 rounds, multipliers
 * 0xD2511F53 / 0xCD9E8D57, Weyl constants 0x9E3779B9 / 0xBB67AE85) with
 *   counter = (trial_lo, trial_hi, draw_index / 4, 0), key = (seed_lo, seed_hi), word draw_index % 4,
 * one draw per step.  Rasmussen picks column rank (draw * deg) >> 32; the scaled estimator
 * compares ((draw + 1) / 2^32) * total with the running sum, as the reference does with
 * curand_uniform in (0, 1].
 * Parity status: PINNED statistically against closed forms (Kasteleyn, n!) and, for the estimator
 * logic, against the reference CPU functions rasmussen_sparse / approximation_perman64_sparse of
 * oracle/_ref (same distribution, different RNG) in tests/test_oracle.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned long long u64;

static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                          uint32_t out[4]) {
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_philox(u64 seed, u64 trial, uint32_t block, uint32_t out[4]) {
  philox4x32_10((uint32_t)trial, (uint32_t)(trial >> 32), block, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), out);
}

static uint32_t draw_of(u64 seed, u64 trial, int step) {
  uint32_t w[4];
  orc_philox(seed, trial, (uint32_t)(step >> 2), w);
  return w[step & 3];
}

#define TESTBIT(m, i) (((m)[(i) >> 5] >> ((i) & 31)) & 1u)
#define SETBIT(m, i) ((m)[(i) >> 5] |= 1u << ((i) & 31))

/* first remaining row with the fewest remaining columns (gpu_approximation_sparse.cu:242-256) */
static int min_degree_row(const int *rptrs, const int *cols, int nov, const unsigned *rowx,
                          const unsigned *colx, int *deg_out) {
  int best = nov + 1, row = -1;
  for (int r = 0; r < nov; ++r) {
    if (TESTBIT(rowx, r)) continue;
    int d = 0;
    for (int t = rptrs[r]; t < rptrs[r + 1]; ++t)
      if (!TESTBIT(colx, cols[t])) ++d;
    if (best > d) { best = d; row = r; }
  }
  *deg_out = best;
  return row;
}

/* Every trial function also reports how far the trial got: *steps_out = steps completed (nov when it
 * reached the last step) and *partial_out = the running product at that point; the estimate is 0 for a
 * dead end and the full product otherwise.  (The reference only returns the estimate; the record lets a
 * parity test say something on patterns where almost every trial dies.) */
double orc_rasmussen_trace(const int *rptrs, const int *cols, int nov, u64 seed, u64 trial, int *steps_out,
                           double *partial_out) {
  const int words = (nov + 31) / 32;
  unsigned *rowx = (unsigned *)calloc((size_t)words, sizeof(unsigned));
  unsigned *colx = (unsigned *)calloc((size_t)words, sizeof(unsigned));
  double perm = 1.0;
  int step = 0, dead = 0;
  for (; step < nov; ++step) {
    int deg;
    const int row = min_degree_row(rptrs, cols, nov, rowx, colx, &deg);
    if (deg == 0) { dead = 1; break; }
    perm *= deg;
    int want = (int)(((uint64_t)draw_of(seed, trial, step) * (uint64_t)deg) >> 32);
    int col = -1;
    for (int t = rptrs[row]; t < rptrs[row + 1]; ++t) {
      const int c = cols[t];
      if (TESTBIT(colx, c)) continue;
      if (want == 0) { col = c; break; }
      --want;
    }
    SETBIT(colx, col);
    SETBIT(rowx, row);
  }
  free(rowx); free(colx);
  if (steps_out) *steps_out = step;
  if (partial_out) *partial_out = perm;
  return dead ? 0.0 : perm;
}

double orc_rasmussen_trial(const int *rptrs, const int *cols, int nov, u64 seed, u64 trial) {
  return orc_rasmussen_trace(rptrs, cols, nov, seed, trial, NULL, NULL);
}

/* rvals / cvals == NULL: pattern-only sweeps with float sums (sparse kernel);
 * otherwise sums weighted by the entries, in double (dense kernel). */
double orc_scaling_trace(const int *rptrs, const int *cols, const int *cptrs, const int *rows,
                         const double *rvals, const double *cvals, int nov, int scale_intervals,
                         int scale_times, u64 seed, u64 trial, int *steps_out, double *partial_out) {
  const int words = (nov + 31) / 32;
  unsigned *rowx = (unsigned *)calloc((size_t)words, sizeof(unsigned));
  unsigned *colx = (unsigned *)calloc((size_t)words, sizeof(unsigned));
  float *d_r = (float *)malloc((size_t)nov * sizeof(float));
  float *d_c = (float *)malloc((size_t)nov * sizeof(float));
  for (int i = 0; i < nov; ++i) { d_r[i] = 1.0f; d_c[i] = 1.0f; }
  const int weighted = (rvals != NULL && cvals != NULL);
  double perm = 1.0;
  int step = 0, dead = 0;
  for (; step < nov; ++step) {
    int deg;
    const int row = min_degree_row(rptrs, cols, nov, rowx, colx, &deg);
    if (deg == 0) { dead = 1; break; }   /* the sum below would be 0: same outcome */
    if (step % scale_intervals == 0) {
      for (int k = 0; k < scale_times && !dead; ++k) {
        for (int j = 0; j < nov && !dead; ++j) {
          if (TESTBIT(colx, j)) continue;
          if (weighted) {
            double cs = 0.0;
            for (int t = cptrs[j]; t < cptrs[j + 1]; ++t)
              if (!TESTBIT(rowx, rows[t])) cs += (double)d_r[rows[t]] * cvals[t];
            if (cs == 0.0) dead = 1; else d_c[j] = (float)(1.0 / cs);
          } else {
            float cs = 0.0f;
            for (int t = cptrs[j]; t < cptrs[j + 1]; ++t)
              if (!TESTBIT(rowx, rows[t])) cs += d_r[rows[t]];
            if (cs == 0.0f) dead = 1; else d_c[j] = 1.0f / cs;
          }
        }
        for (int i = 0; i < nov && !dead; ++i) {
          if (TESTBIT(rowx, i)) continue;
          if (weighted) {
            double rs = 0.0;
            for (int t = rptrs[i]; t < rptrs[i + 1]; ++t)
              if (!TESTBIT(colx, cols[t])) rs += rvals[t] * (double)d_c[cols[t]];
            if (rs == 0.0) dead = 1; else d_r[i] = (float)(1.0 / rs);
          } else {
            float rs = 0.0f;
            for (int t = rptrs[i]; t < rptrs[i + 1]; ++t)
              if (!TESTBIT(colx, cols[t])) rs += d_c[cols[t]];
            if (rs == 0.0f) dead = 1; else d_r[i] = 1.0f / rs;
          }
        }
      }
    }
    if (dead) break;
    const float dr = d_r[row];
    double tot = 0.0;
    for (int t = rptrs[row]; t < rptrs[row + 1]; ++t)
      if (!TESTBIT(colx, cols[t])) tot += (double)(dr * d_c[cols[t]]);
    if (tot == 0.0) { dead = 1; break; }
    const double target = ((double)draw_of(seed, trial, step) + 1.0) * (1.0 / 4294967296.0) * tot;
    double run = 0.0;
    int col = -1;
    for (int t = rptrs[row]; t < rptrs[row + 1]; ++t) {
      const int c = cols[t];
      if (TESTBIT(colx, c)) continue;
      const double s = (double)(dr * d_c[c]);
      run += s;
      if (target <= run) { col = c; perm /= (s / tot); break; }
    }
    if (col < 0) { dead = 1; break; }
    SETBIT(colx, col);
    SETBIT(rowx, row);
  }
  free(rowx); free(colx); free(d_r); free(d_c);
  if (steps_out) *steps_out = step;
  if (partial_out) *partial_out = perm;
  return dead ? 0.0 : perm;
}

double orc_scaling_trial(const int *rptrs, const int *cols, const int *cptrs, const int *rows,
                         const double *rvals, const double *cvals, int nov, int scale_intervals,
                         int scale_times, u64 seed, u64 trial) {
  return orc_scaling_trace(rptrs, cols, cptrs, rows, rvals, cvals, nov, scale_intervals, scale_times, seed, trial, NULL,
                           NULL);
}

/* mean over trials [lo, hi), sequential sum in trial order */
double orc_rasmussen_mean(const int *rptrs, const int *cols, int nov, u64 seed, u64 lo, u64 hi) {
  double s = 0.0;
  for (u64 t = lo; t < hi; ++t) s += orc_rasmussen_trial(rptrs, cols, nov, seed, t);
  return s / (double)(hi - lo);
}
