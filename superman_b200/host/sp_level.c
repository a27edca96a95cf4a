/* sp_level.c -- host-side planning of the sparse exact paths (SpaRyser / SkipPer), in C: everything that is
 * decided from the matrix before a kernel runs.
 *
 *   1. column order: a plan that covers the whole index space may walk the columns 0 .. n-2 in any order
 *      (the Ryser sum runs over every subset of them); the B most frequently flipped ones are chosen so that
 *      the rows' level populations fit the cheapest slot configuration of the LevelRyser engine;
 *   2. row order: ascending by level (the lowest flippable column holding a non-zero of the row);
 *   3. engine and parameters: LevelRyser (B, S0, S, R) by an FP64-instruction cost model, or the hot/cold
 *      register kernel when some level has more rows than any slot configuration takes;
 *   4. the packed images the kernels stage into shared memory.
 *
 * The device layer (csrc/sp_sparse.cu, spd_sparse_plan_create_packed) only uploads what is prepared here and
 * launches.  The reference does none of this: its sparse kernels walk the CCS arrays per step
 * (gpu_exact_sparse.cu:455-670) and the only preprocessing is SortOrder / SkipOrder (util.h:553-684), which
 * stays what -r1 / -r2 select in front of this. */
#define _POSIX_C_SOURCE 200809L
#include "sp_sched.h"
#include "superman_b200_level.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int env_int_c(const char *name, int dflt) {
  const char *s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}

static const int k_slot_opts[6] = {1, 2, 3, 4, 6, 8};

/* hot-slot FP64 instructions per index of the cheapest (S0, S) that takes the level populations pops[0..B),
 * or < 0 when none does.  Rows that do not fit their own level's slots go to free slots of lower levels. */
static double fit_slots(const int *pops, int B) {
  for (int si = 0; si < 6; ++si) {
    const int S = k_slot_opts[si];
    for (int S0 = (S - 2 > 1 ? S - 2 : 1); S0 <= S; ++S0) {
      int free_[4] = {S0, S, S, S};
      int ok = 1;
      for (int L = 0; L < B && ok; ++L) {
        int need = pops[L];
        for (int LL = L; LL >= 0 && need > 0; --LL) {
          const int t = need < free_[LL] ? need : free_[LL];
          free_[LL] -= t; need -= t;
        }
        ok = (need == 0);
      }
      if (ok) {
        double c = 0.0;
        for (int L = 0; L < B; ++L) c += (double)(L == 0 ? S0 : S) * 2.0 * (double)(1 << (B - L));
        return c / (double)(1 << B);
      }
    }
  }
  return -1.0;
}

/* Column order for a whole-space plan: among the ordered B-tuples (B = 3, 4) of the 9 sparsest columns take
 * the one whose level populations fit the cheapest slot configuration; the other columns keep their order;
 * column n-1 stays last (it is the one the Nijenhuis-Wilf start vector is built around).  SortOrder puts the
 * sparsest columns first but knows nothing about slots: level populations (4, 5, 3, 3) need S = 6, the same
 * matrix with columns 1 and 2 exchanged has (4, 4, 4, 3) and fits S = 4: 15 instead of 18.5 FP64
 * instructions per index.  perm[k'] = original column at position k'. */
static void choose_low_columns(int n, const double *dmat_t, int *perm) {
  for (int k = 0; k < n; ++k) perm[k] = k;
  if (n < 10) return;
  int cnt[64], order[64];
  for (int k = 0; k < n - 1; ++k) {
    int c = 0;
    for (int j = 0; j < n; ++j) c += dmat_t[(size_t)k * n + j] != 0.0;
    cnt[k] = c; order[k] = k;
  }
  for (int i = 1; i < n - 1; ++i) {            /* stable insertion sort by count */
    const int k = order[i];
    int p = i;
    while (p > 0 && cnt[order[p - 1]] > cnt[k]) { order[p] = order[p - 1]; --p; }
    order[p] = k;
  }
  const int K = n - 1 < 9 ? n - 1 : 9;
  unsigned long long rows_of[64];
  for (int k = 0; k < n - 1; ++k) {
    rows_of[k] = 0ull;
    for (int j = 0; j < n; ++j)
      if (dmat_t[(size_t)k * n + j] != 0.0) rows_of[k] |= 1ull << j;
  }
  double best = 1e300;
  int best_t[4] = {0, 1, 2, 3}, best_B = 0;
  for (int B = 3; B <= 4; ++B) {               /* the caller's order is the incumbent */
    unsigned long long seen = 0ull;
    int pops[4];
    for (int L = 0; L < B; ++L) { pops[L] = __builtin_popcountll(rows_of[L] & ~seen); seen |= rows_of[L]; }
    const double c = fit_slots(pops, B);
    if (c >= 0 && c < best) best = c;
  }
  for (int B = 3; B <= 4; ++B) {
    int t[4];
    for (t[0] = 0; t[0] < K; ++t[0])
      for (t[1] = 0; t[1] < K; ++t[1]) {
        if (t[1] == t[0]) continue;
        for (t[2] = 0; t[2] < K; ++t[2]) {
          if (t[2] == t[0] || t[2] == t[1]) continue;
          for (t[3] = 0; t[3] < (B == 4 ? K : 1); ++t[3]) {
            if (B == 4 && (t[3] == t[0] || t[3] == t[1] || t[3] == t[2])) continue;
            unsigned long long seen = 0ull;
            int pops[4];
            for (int L = 0; L < B; ++L) {
              const unsigned long long r = rows_of[order[t[L]]];
              pops[L] = __builtin_popcountll(r & ~seen);
              seen |= r;
            }
            const double c = fit_slots(pops, B);
            if (c >= 0 && c < best - 0.26) {   /* at least a quarter instruction per index better */
              best = c; best_B = B;
              for (int L = 0; L < B; ++L) best_t[L] = order[t[L]];
            }
          }
        }
      }
  }
  if (best_B == 0) return;
  char used[64];
  memset(used, 0, sizeof(used));
  int pos = 0;
  for (int L = 0; L < best_B; ++L) { perm[pos++] = best_t[L]; used[best_t[L]] = 1; }
  for (int k = 0; k < n - 1; ++k)
    if (!used[k]) perm[pos++] = k;
  perm[n - 1] = n - 1;
}

static int count_level_below(const int *level_sorted, int n, int bound) {
  int c = 0;
  while (c < n && level_sorted[c] < bound) ++c;
  return c;
}

/* FP64 instructions per Gray index of the hot/cold register kernel (sparse_reg.cuh) */
static double hotcold_cost(const int *level_sorted, int n, int B) {
  int H = count_level_below(level_sorted, n, B);
  H = ((H + 3) / 4) * 4;
  if (H > n) H = n;
  return 2.0 * H + (2.0 * (n - H) + 8.0) / (double)(1 << B);
}

typedef struct level_pack_out {
  double *colT_hot, *lowR, *dcold, *xb_hot, *xb_cold;
  int *cold_start;
  int NC, NCP, HSP;
  double cost;                 /* planning cost per index (FP64 instructions, the cold refresh weighted x3) */
  double fp64_per_index;       /* FP64 instructions the kernel executes per index for this packing */
} level_pack_out;

static void pack_free(level_pack_out *o) {
  free(o->colT_hot); free(o->lowR); free(o->dcold); free(o->xb_hot); free(o->xb_cold); free(o->cold_start);
  memset(o, 0, sizeof(*o));
}

/* Packs the rows into the LevelRyser layout for (B, S0, S, R): S0 register slots for level 0, S for each other
 * level < B, R register-cold rows, the other rows cold, sorted by level.  lvl[] / dmat_t / xbase are in the
 * caller's row order.  Returns 0 when some level has more rows than the slots at or below it can take,
 * < 0 on allocation failure, 1 on success. */
static int level_pack(int n, int B, int S0, int S, int R, int fill_free, const int *lvl, const double *dmat_t,
                      const double *xbase, level_pack_out *o) {
  memset(o, 0, sizeof(*o));
  const int HS = S0 + (B - 1) * S, HT = HS + R, HSP = HT + (HT & 1), LB = B + (B & 1);
  int slot_row[64], order[64], cold[64];
  for (int s = 0; s < HT; ++s) slot_row[s] = -1;
  for (int j = 0; j < n; ++j) order[j] = j;
  for (int i = 1; i < n; ++i) {                /* stable insertion sort by level */
    const int j = order[i];
    int p = i;
    while (p > 0 && lvl[order[p - 1]] > lvl[j]) { order[p] = order[p - 1]; --p; }
    order[p] = j;
  }
  int ncold = 0;
  /* hot rows by ascending level: own level's slots first, then any free slot of a lower level */
  for (int idx = 0; idx < n; ++idx) {
    const int j = order[idx];
    if (lvl[j] >= B) { cold[ncold++] = j; continue; }
    int placed = -1;
    for (int L = lvl[j]; L >= 0 && placed < 0; --L) {
      const int base = L == 0 ? 0 : S0 + (L - 1) * S, count = L == 0 ? S0 : S;
      for (int t = 0; t < count; ++t)
        if (slot_row[base + t] < 0) { placed = base + t; break; }
    }
    if (placed < 0) return 0;
    slot_row[placed] = j;
  }
  /* SpaRyser: a level slot that is still free does its work anyway (on the neutral row), so it takes a cold row
   * instead -- a row of level >= B has no entry in the low columns and is simply constant over the block.  The
   * cold rows of lowest level go first (they are refreshed most often), into the cheapest free slots (highest
   * level: fewest values per block).  SkipPer keeps them cold: its block test looks at the cold product only. */
  int moved = 0;
  if (fill_free) {
    for (int L = B - 1; L >= 0 && moved < ncold; --L) {
      const int base = L == 0 ? 0 : S0 + (L - 1) * S, count = L == 0 ? S0 : S;
      for (int t = 0; t < count && moved < ncold; ++t)
        if (slot_row[base + t] < 0) slot_row[base + t] = cold[moved++];
    }
  }
  /* the R cold rows of lowest level (refreshed most often) stay in registers too */
  const int nrc = R < ncold - moved ? R : ncold - moved;
  for (int t = 0; t < nrc; ++t) slot_row[HS + t] = cold[moved + t];
  const int NC = ncold - moved - nrc;
  const int *coldp = cold + moved + nrc;
  const int NCP = NC + (NC & 1) + (NC == 0 ? 2 : 0);
  o->colT_hot = (double *)calloc((size_t)(n - 1) * HSP, sizeof(double));
  o->lowR = (double *)calloc((size_t)(HS > 0 ? HS : 1) * LB, sizeof(double));
  o->dcold = (double *)calloc((size_t)(n - 1) * NCP, sizeof(double));
  o->xb_hot = (double *)malloc((size_t)HSP * sizeof(double));
  o->xb_cold = (double *)malloc((size_t)NCP * sizeof(double));
  o->cold_start = (int *)malloc((size_t)(n - B + 2) * sizeof(int));
  if (!o->colT_hot || !o->lowR || !o->dcold || !o->xb_hot || !o->xb_cold || !o->cold_start) { pack_free(o); return -1; }
  for (int s = 0; s < HSP; ++s) o->xb_hot[s] = 1.0;      /* neutral slot: x = 1, all entries 0 */
  for (int s = 0; s < NCP; ++s) o->xb_cold[s] = 1.0;
  for (int sl = 0; sl < HT; ++sl) {
    const int j = slot_row[sl];
    if (j < 0) continue;
    o->xb_hot[sl] = xbase[j];
    for (int k = 0; k < n - 1; ++k) o->colT_hot[(size_t)k * HSP + sl] = dmat_t[(size_t)k * n + j];
    if (sl < HS)
      for (int q = 0; q < B; ++q) o->lowR[(size_t)sl * LB + q] = dmat_t[(size_t)q * n + j];
  }
  for (int jc = 0; jc < NC; ++jc) {
    const int j = coldp[jc];
    o->xb_cold[jc] = xbase[j];
    for (int k = 0; k < n - 1; ++k) o->dcold[(size_t)k * NCP + jc] = dmat_t[(size_t)k * n + j];
  }
  /* cold_start[i] = first cold row with level >= B+i  (levels run up to n = "never touched") */
  {
    int jc = 0;
    for (int i = 0; i <= n - B; ++i) {
      while (jc < NC && lvl[coldp[jc]] < B + i) ++jc;
      o->cold_start[i] = jc;
    }
    o->cold_start[n - B + 1] = NC;
  }
  /* cost per index: hot slots + register-cold rows + pair sums + expected cold refresh (x3: it runs from
   * shared memory, one dependent chain) */
  double hot = 0.0;
  for (int L = 0; L < B; ++L) hot += (double)(L == 0 ? S0 : S) * 2.0 * (double)(1 << (B - L));
  double coldc = 0.0, w = 0.5;
  for (int z = 0; z < 16 && B + z <= n; ++z, w *= 0.5)
    coldc += w * 3.0 * (double)o->cold_start[(z + 1 <= n - B + 1) ? z + 1 : n - B + 1];
  o->cost = (hot + 2.0 * R + (double)((1 << B) + B) + coldc) / (double)(1 << B);
  /* what the block loop executes (level_reg.cuh): per level SL slots x 2^(B-L) values, one add and one multiply
   * each, less one multiply per value (a product of SL factors); pair sums 2^(B-1) + sum_{L>=1} 2^(B-L); the
   * block's accumulate; R register-cold rows (add + multiply) and their two-chain product; the expected cold
   * refresh (one FMA and one multiply per row touched) */
  {
    double ex = 0.0;
    for (int L = 0; L < B; ++L) {
      const double SL = (double)(L == 0 ? S0 : S), cnt = (double)(1 << (B - L));
      ex += SL * cnt + (SL - 1.0) * cnt + (L == 0 ? 0.5 * cnt : cnt);
    }
    ex += 1.0 + (R > 0 ? 2.0 * R + 2.0 : 0.0) + coldc * (2.0 / 3.0);
    o->fp64_per_index = ex / (double)(1 << B);
  }
  o->NC = NC; o->NCP = NCP; o->HSP = HSP;
  return 1;
}

int sp_level_plan_build(const double *dmat_in, const double *xbase, int nov, int skip, int flags, sp_level_plan *out) {
  if (!dmat_in || !xbase || !out) { sp_set_error("null argument"); return SP_EINVAL; }
  memset(out, 0, sizeof(*out));
  if (nov < 2 || nov > 64) { sp_set_error("sparse Ryser supports 2 <= n <= 64 (got %d)", nov); return SP_ELIMIT; }
  const int n = nov;
  int rc = SP_OK;
  double *dperm = NULL, *mt = NULL;
  const double *dmat_t = dmat_in;
  level_pack_out best;
  memset(&best, 0, sizeof(best));

  /* 1. column order: only for plans that will cover the whole index space */
  if ((flags & SP_PLAN_WHOLE_SPACE) && env_int_c("SP_SPARSE_REORDER", 1) != 0) {
    int cperm[64], moved = 0;
    choose_low_columns(n, dmat_in, cperm);
    for (int k = 0; k < n; ++k) moved |= (cperm[k] != k);
    if (moved) {
      dperm = (double *)malloc((size_t)n * n * sizeof(double));
      if (!dperm) { sp_set_error("out of memory"); return SP_ENOMEM; }
      for (int k = 0; k < n; ++k) memcpy(&dperm[(size_t)k * n], &dmat_in[(size_t)cperm[k] * n], (size_t)n * sizeof(double));
      dmat_t = dperm;
    }
  }

  /* 2. row order: ascending by the lowest flippable column (0 .. n-2) holding a non-zero of the row; rows
   * touched by no such column come last.  Stable, so equal rows keep the caller's order. */
  int lvl[64], perm[64], level_sorted[64];
  double xb[64];
  for (int j = 0; j < n; ++j) {
    int l = n;
    for (int k = 0; k < n - 1; ++k)
      if (dmat_t[(size_t)k * n + j] != 0.0) { l = k; break; }
    lvl[j] = l; perm[j] = j;
  }
  for (int i = 1; i < n; ++i) {
    const int j = perm[i];
    int p = i;
    while (p > 0 && lvl[perm[p - 1]] > lvl[j]) { perm[p] = perm[p - 1]; --p; }
    perm[p] = j;
  }
  mt = (double *)malloc((size_t)n * n * sizeof(double));
  if (!mt) { free(dperm); sp_set_error("out of memory"); return SP_ENOMEM; }
  for (int j = 0; j < n; ++j) {
    level_sorted[j] = lvl[perm[j]];
    xb[j] = xbase[perm[j]];
    for (int k = 0; k < n; ++k) mt[(size_t)k * n + j] = dmat_t[(size_t)k * n + perm[j]];
  }

  /* development aid: SP_LEVEL_DUMP=<file> gets the matrix as the plan walks it (rows by level, columns in plan
   * order) for tools/skip_structure.py */
  if (getenv("SP_LEVEL_DUMP")) {
    FILE *f = fopen(getenv("SP_LEVEL_DUMP"), "w");
    if (f) {
      fprintf(f, "%d %d\n", n, skip);
      for (int j = 0; j < n; ++j) {
        fprintf(f, "%d %.17g", level_sorted[j], xb[j]);
        for (int k = 0; k < n; ++k) fprintf(f, " %.17g", mt[(size_t)k * n + j]);
        fprintf(f, "\n");
      }
      fclose(f);
    }
  }

  /* 3. engine: LevelRyser (B, S0, S) with the lowest modelled cost, if the matrix fits its slots */
  spd_level_image img;
  memset(&img, 0, sizeof(img));
  const double hc3 = hotcold_cost(level_sorted, n, 3), hc4 = hotcold_cost(level_sorted, n, 4);
  const double hc_instr = hc3 < hc4 ? hc3 : hc4;
  const double hc_cost = hc_instr * 1.45;                    /* measured: ~68 % of the pipe */
  img.instr_per_index = hc_instr;
  const int engine = env_int_c("SP_SPARSE_ENGINE", 0);       /* 0 auto, 1 hot/cold, 2 level */
  double lv_cost = 1e300;
  if (engine != 1 && n >= 6) {
    const int forceB = env_int_c("SP_SPARSE_LOWCOLS", 0), forceS = env_int_c("SP_LEVEL_SLOTS", 0),
              forceS0 = env_int_c("SP_LEVEL_SLOTS0", 0);
    for (int B = 3; B <= 4; ++B) {
      if (B + 2 > n - 1) continue;
      if (forceB && forceB != B) continue;
      for (int si = 0; si < 6; ++si) {
        const int S = k_slot_opts[si];
        if (forceS && forceS != S) continue;
        int fits = 0;
        /* level 0 (2^B values per block, the most expensive level) may have up to two slots fewer */
        for (int S0 = (S - 2 > 1 ? S - 2 : 1); S0 <= S; ++S0) {
          if (forceS0 && forceS0 != S0) continue;
          const int R = spl_regcold(B, S0, S, skip != 0);
          level_pack_out o;
          const int r = level_pack(n, B, S0, S, R, skip == 0 && env_int_c("SP_LEVEL_FILL", 1) != 0, lvl, dmat_t, xbase, &o);
          if (r < 0) { rc = SP_ENOMEM; sp_set_error("out of memory"); goto done; }
          if (r == 0) continue;
          fits = 1;
          double cost = o.cost * 1.15;
          if (spl_minblocks(B, S0, S, skip != 0) < 4) cost *= 1.1;      /* 3 instead of 4 blocks per SM */
          if (cost < lv_cost) {
            lv_cost = cost;
            pack_free(&best);
            best = o;
            img.B = B; img.S0 = S0; img.S = S; img.R = R;
          } else {
            pack_free(&o);
          }
          break;   /* more level-0 slots for the same S only cost more */
        }
        if (fits) break;   /* a larger S for the same B only costs more */
      }
    }
  }
  if (img.B && (engine == 2 || lv_cost < hc_cost || n > 48)) {
    img.NC = best.NC; img.NCP = best.NCP; img.HSP = best.HSP;
    img.colT_hot = best.colT_hot; img.lowR = best.lowR; img.dcold = best.dcold;
    img.xb_hot = best.xb_hot; img.xb_cold = best.xb_cold; img.cold_start = best.cold_start;
    img.instr_per_index = best.fp64_per_index;
  } else {
    img.B = 0;
  }

  /* 3b. SkipPer tile length.  The kernel drops a tile when one of its tile-constant rows (level >= c, c = log2 of
   * the tile length) is zero; shorter tiles have more such rows and may drop more, longer tiles pay the tile
   * prologue and the filter less often.  With the survival rates s(c) measured on sampled tiles, the time per index
   * goes as   s(c) * (1 + 0.116 * 2^(11-c)) + 0.035 * 2^(11-c)
   * (prologue 11.6 % and filter 3.5 % of the block work at c = 11, ncu at n = 33): take c = 12 over the default 11
   * when that is smaller.  Measured on four n = 33 matrices (s(11) = s(12) on all of them): -7 ... -10 % on three,
   * +3 % on the fourth, whose lanes share fewer zero blocks when the tile bits move up by one. */
  if (skip && img.B && n - 1 > 12 && !(flags & SP_PLAN_WHOLE_SPACE)) {
    img.skip_long_tiles = 1;      /* pieces of a range: their launches are short of tiles as it is */
  } else if (skip && img.B && n - 1 > 12) {
    double surv[2] = {0.0, 0.0};
    unsigned long long rs = 0x9E3779B97F4A7C15ull;
    const int samples = 512;
    for (int ci = 0; ci < 2; ++ci) {
      const int c = 11 + ci;
      int alive = 0;
      for (int smp = 0; smp < samples; ++smp) {
        rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17;                  /* xorshift64 */
        const unsigned long long i = (rs >> 8) & ((1ull << (n - 1)) - 1) & ~((1ull << c) - 1);
        const unsigned long long g = i ^ (i >> 1);
        int ok = 1;
        for (int j = n - 1; j >= 0 && ok && level_sorted[j] >= c; --j) {
          double x = xb[j];
          for (int k = c; k < n - 1; ++k)
            if ((g >> k) & 1ull) x += mt[(size_t)k * n + j];
          if (x == 0.0) ok = 0;
        }
        alive += ok;
      }
      surv[ci] = (double)alive / samples;
    }
    const double t11 = surv[0] * 1.116 + 0.035, t12 = surv[1] * 1.058 + 0.0175;
    img.skip_long_tiles = (t12 < t11 && env_int_c("SP_SKIP_LONG_TILES", 1) != 0) ? 1 : 0;
  }

  /* the plan owns what the image points to */
  out->n = n; out->skip = skip;
  out->mat_t = mt; mt = NULL;
  memcpy(out->xbase, xb, sizeof(double) * (size_t)n);
  memcpy(out->level_sorted, level_sorted, sizeof(int) * (size_t)n);
  out->img = img;
  if (img.B) {
    out->owned[0] = best.colT_hot; out->owned[1] = best.lowR; out->owned[2] = best.dcold;
    out->owned[3] = best.xb_hot; out->owned[4] = best.xb_cold; out->owned[5] = best.cold_start;
    memset(&best, 0, sizeof(best));
  }
done:
  pack_free(&best);
  free(mt);
  free(dperm);
  return rc;
}

void sp_level_plan_free(sp_level_plan *p) {
  if (!p) return;
  free(p->mat_t);
  for (int i = 0; i < 6; ++i) free(p->owned[i]);
  memset(p, 0, sizeof(*p));
}

int sp_sparse_plan_open(int device, const double *dmat_t, const double *xbase, int nov, int skip, int flags,
                        spd_sparse_plan **out) {
  if (!out) { sp_set_error("null argument"); return SP_EINVAL; }
  sp_level_plan plan;
  int rc = sp_level_plan_build(dmat_t, xbase, nov, skip, flags, &plan);
  if (rc != SP_OK) return rc;
  /* hand the images to the device */
  rc = spd_sparse_plan_create_packed(device, plan.mat_t, plan.xbase, plan.level_sorted, plan.n, skip, &plan.img, out);
  if (rc != SPD_OK) sp_set_error("%s", spd_last_error());
  sp_level_plan_free(&plan);
  return rc;
}
