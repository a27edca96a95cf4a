/* oracle.c -- CPU restatement of SUPerman's permanent algorithms.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing under oracle/ is part of the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load liboracle.so, and only as the checker
 * or the timed CPU baseline.  libsuperman_b200.so never links, loads or calls it.
 *
 * Parity status: PINNED.  Every function below is checked (tests/test_oracle.py) against
 *   - the golden values of SURVEY.md 8(c) (reference functions run with 17-digit printing),
 *   - oracle/_ref/libref.so, the reference's own util.h / algo.h compiled unmodified from
 *     /root/reference by oracle/Makefile (in the build container; fixtures generated from it are
 *     committed under tests/golden/ with the script that made them),
 *   - closed forms (n!, derangements, Kasteleyn's grid formula).
 *
 * Each function cites the reference lines it restates (paths relative to /root/reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned long long u64;
typedef __int128 i128;
typedef unsigned __int128 u128;

/* ------------------------------------------------------------------------------------------
 * Nijenhuis-Wilf start vector: x[j] = A[j][n-1] - rowsum_j / 2
 * (algo.h:1043-1049, gpu_exact_dense.cu:647-654; row sum accumulated in double)
 * ---------------------------------------------------------------------------------------- */
void orc_nw_base(const double *mat, int n, double *x) {
  for (int j = 0; j < n; ++j) {
    double rs = 0.0;
    for (int k = 0; k < n; ++k) rs += mat[j * n + k];
    x[j] = mat[j * n + (n - 1)] - rs / 2;
  }
}

double orc_nw_factor(int n) { return (double)(4 * (n & 1) - 2); }   /* algo.h:1087 */

/* ------------------------------------------------------------------------------------------
 * Dense Ryser partial sum over Gray indices [lo, hi) in double, ONE serial chain:
 * the arithmetic of perman64 (algo.h:1031-1088) / cpu_perman64 with one thread
 * (gpu_exact_dense.cu:6-69): explicit X at the range start from gray(lo-1), then
 * k = ctz(i), s = +-1, x += s*col_k, prod = x[0]*x[1]*...*x[n-1] (ascending, starting from 1.0),
 * acc += (-1)^i prod.  lo == 0 includes index 0 (the base term, algo.h:1049).
 * ---------------------------------------------------------------------------------------- */
double orc_ryser_range_f64(const double *mat, int n, u64 lo, u64 hi) {
  double x[64], acc = 0.0;
  orc_nw_base(mat, n, x);
  u64 i = lo;
  if (i >= hi) return 0.0;
  u64 g = 0;
  if (i == 0) {
    double p = 1.0;
    for (int j = 0; j < n; ++j) p *= x[j];
    acc = p;
    i = 1;
  } else {
    g = (i - 1) ^ ((i - 1) >> 1);
    for (int k = 0; k < n - 1; ++k)
      if ((g >> k) & 1ull)
        for (int j = 0; j < n; ++j) x[j] += mat[j * n + k];
  }
  for (; i < hi; ++i) {
    const int k = __builtin_ctzll(i);
    g ^= (1ull << k);
    const double s = ((g >> k) & 1ull) ? 1.0 : -1.0;
    double prod = 1.0;
    for (int j = 0; j < n; ++j) {
      x[j] += s * mat[j * n + k];
      prod *= x[j];
    }
    acc += (i & 1ull) ? -prod : prod;
  }
  return acc;
}

/* Same sum in long double, cut into 2^14-index pieces that are summed pairwise: the
 * high-precision value parity is asserted against ("the long-double oracle"). */
static long double ryser_piece_ld(const double *mat, int n, const long double *xb, u64 lo, u64 hi) {
  long double x[64], acc = 0.0L;
  for (int j = 0; j < n; ++j) x[j] = xb[j];
  u64 i = lo, g = 0;
  if (i == 0) {
    long double p = 1.0L;
    for (int j = 0; j < n; ++j) p *= x[j];
    acc = p;
    i = 1;
  } else {
    g = (i - 1) ^ ((i - 1) >> 1);
    for (int k = 0; k < n - 1; ++k)
      if ((g >> k) & 1ull)
        for (int j = 0; j < n; ++j) x[j] += (long double)mat[j * n + k];
  }
  for (; i < hi; ++i) {
    const int k = __builtin_ctzll(i);
    g ^= (1ull << k);
    const long double s = ((g >> k) & 1ull) ? 1.0L : -1.0L;
    long double prod = 1.0L;
    for (int j = 0; j < n; ++j) {
      x[j] += s * (long double)mat[j * n + k];
      prod *= x[j];
    }
    acc += (i & 1ull) ? -prod : prod;
  }
  return acc;
}

long double orc_ryser_range_ld(const double *mat, int n, u64 lo, u64 hi) {
  if (hi <= lo) return 0.0L;
  long double xb[64];
  for (int j = 0; j < n; ++j) {
    long double rs = 0.0L;
    for (int k = 0; k < n; ++k) rs += (long double)mat[j * n + k];
    xb[j] = (long double)mat[j * n + (n - 1)] - rs / 2;
  }
  const u64 piece = 1ull << 14;
  const u64 first = lo / piece, last = (hi - 1) / piece;
  const long long np = (long long)(last - first + 1);
  long double *part = (long double *)malloc((size_t)np * sizeof(long double));
#pragma omp parallel for schedule(dynamic, 64)
  for (long long p = 0; p < np; ++p) {
    u64 a = (first + (u64)p) * piece, b = a + piece;
    if (a < lo) a = lo;
    if (b > hi) b = hi;
    part[p] = ryser_piece_ld(mat, n, xb, a, b);
  }
  /* pairwise tree */
  long long m = np;
  while (m > 1) {
    const long long h = (m + 1) / 2;
    for (long long p = 0; p + h < m; ++p) part[p] += part[p + h];
    m = h;
  }
  const long double r = part[0];
  free(part);
  return r;
}

/* full permanent = (4*(n&1)-2) * sum over [0, 2^(n-1)) ; n == 1 -> A[0][0] */
long double orc_perm_ld(const double *mat, int n) {
  if (n == 1) return (long double)mat[0];
  return (long double)orc_nw_factor(n) * orc_ryser_range_ld(mat, n, 0, 1ull << (n - 1));
}
double orc_perm_ld_as_double(const double *mat, int n) { return (double)orc_perm_ld(mat, n); }
double orc_ryser_range_ld_as_double(const double *mat, int n, u64 lo, u64 hi) {
  return (double)orc_ryser_range_ld(mat, n, lo, hi);
}

/* serial double permanent: perman64 (algo.h:1031-1088) */
double orc_perm_f64(const double *mat, int n) {
  if (n == 1) return mat[0];
  return orc_nw_factor(n) * orc_ryser_range_f64(mat, n, 0, 1ull << (n - 1));
}

/* ------------------------------------------------------------------------------------------
 * Exact integer Ryser for small-integer matrices: work with
 * y_j = 2*x_j = 2*A[j][n-1] - rowsum_j (integers); perm = (4(n&1)-2) * sum / 2^n ... precisely
 * prod_j x_j = prod_j y_j / 2^n, so perm = (4(n&1)-2) * S / 2^n with S = sum_i (-1)^i prod_j y_j(i).
 * S is accumulated modulo 2^128 (wrap-around is harmless as long as |S| < 2^127, which holds for
 * n <= 24 with entries <= 1 and for the small-entry cases the tests use).  Returns S / 2^(n-1)
 * with the sign factor applied, split in two 64-bit halves (two's complement).
 * ---------------------------------------------------------------------------------------- */
void orc_perm_i128(const int *mat, int n, long long *hi_out, u64 *lo_out) {
  if (n == 1) { i128 v = mat[0]; *hi_out = (long long)(v >> 64); *lo_out = (u64)v; return; }
  i128 y[64];
  for (int j = 0; j < n; ++j) {
    long long rs = 0;
    for (int k = 0; k < n; ++k) rs += mat[j * n + k];
    y[j] = (i128)(2LL * mat[j * n + (n - 1)] - rs);
  }
  u128 S = 0;
  {
    u128 p = 1;
    for (int j = 0; j < n; ++j) p *= (u128)y[j];
    S = p;
  }
  u64 g = 0;
  const u64 end = 1ull << (n - 1);
  for (u64 i = 1; i < end; ++i) {
    const int k = __builtin_ctzll(i);
    g ^= (1ull << k);
    const int s = ((g >> k) & 1ull) ? 2 : -2;
    u128 p = 1;
    for (int j = 0; j < n; ++j) {
      y[j] += (i128)(s * mat[j * n + k]);
      p *= (u128)y[j];
    }
    if (i & 1ull) S -= p; else S += p;
  }
  /* perm = f * S / 2^n, f = +2 (n odd) or -2 (n even)  ->  +-S / 2^(n-1) */
  i128 Ss = (i128)S;
  i128 v = Ss / ((i128)1 << (n - 1));   /* exact: S is divisible by 2^(n-1) */
  if ((n & 1) == 0) v = -v;
  *hi_out = (long long)(v >> 64);
  *lo_out = (u64)v;
}

/* ------------------------------------------------------------------------------------------
 * CRS / CCS builders and orderings (util.h:522-684), element type double.
 * Arrays are caller-allocated: cptrs, rptrs [n+1]; rows, cols, cvals, rvals [>= count of >0 entries].
 * ---------------------------------------------------------------------------------------- */
/* matrix2compressed (util.h:522-551): only entries > 0 are kept; returns their count */
int orc_matrix2compressed(const double *mat, int n, int *cptrs, int *rows, double *cvals, int *rptrs,
                          int *cols, double *rvals) {
  int er = 0, ec = 0;
  for (int i = 0; i < n; ++i) {
    rptrs[i] = er;
    cptrs[i] = ec;
    for (int j = 0; j < n; ++j) {
      if (mat[i * n + j] > 0) { cols[er] = j; rvals[er] = mat[i * n + j]; ++er; }
      if (mat[j * n + i] > 0) { rows[ec] = j; cvals[ec] = mat[j * n + i]; ++ec; }
    }
  }
  rptrs[n] = er;
  cptrs[n] = ec;
  return er;
}

/* matrix2compressed_sortOrder (util.h:553-619): columns ascending by their count of > 0 entries.
 * The reference's qsort comparator returns only 0/1 (util.h:566-570); on glibc it behaves as a
 * STABLE ascending sort (ties keep the original column order) -- SURVEY.md Appendix B.  `mat` is
 * rewritten in the new column order (entries <= 0 become 0, util.h:608-618); colperm[new] = old. */
int orc_sort_order(double *mat, int n, int *cptrs, int *rows, double *cvals, int *rptrs, int *cols,
                   double *rvals, int *colperm) {
  int cnt[64 * 16];
  int *count = (n <= 1024) ? cnt : (int *)malloc((size_t)n * sizeof(int));
  for (int j = 0; j < n; ++j) {
    int c = 0;
    for (int i = 0; i < n; ++i) if (mat[i * n + j] > 0) ++c;
    count[j] = c;
    colperm[j] = j;
  }
  /* stable insertion sort by count */
  for (int a = 1; a < n; ++a) {
    const int v = colperm[a];
    int b = a - 1;
    while (b >= 0 && count[colperm[b]] > count[v]) { colperm[b + 1] = colperm[b]; --b; }
    colperm[b + 1] = v;
  }
  int er = 0, ec = 0;
  for (int i = 0; i < n; ++i) {
    rptrs[i] = er;
    for (int idx = 0; idx < n; ++idx) {
      const int j = colperm[idx];
      if (mat[i * n + j] > 0) { cols[er] = idx; rvals[er] = mat[i * n + j]; ++er; }
    }
  }
  rptrs[n] = er;
  for (int idx = 0; idx < n; ++idx) {
    const int j = colperm[idx];
    cptrs[idx] = ec;
    for (int i = 0; i < n; ++i)
      if (mat[i * n + j] > 0) { rows[ec] = i; cvals[ec] = mat[i * n + j]; ++ec; }
  }
  cptrs[n] = ec;
  for (int i = 0; i < n * n; ++i) mat[i] = 0;
  for (int j = 0; j < n; ++j)
    for (int t = cptrs[j]; t < cptrs[j + 1]; ++t) mat[rows[t] * n + j] = cvals[t];
  if (count != cnt) free(count);
  return er;
}

/* matrix2compressed_skipOrder (util.h:621-684): greedy min-degree column (first minimum, sentinel
 * INT8_MAX = 127), rows appended at first touch, degrees of the other columns those rows touch
 * decremented; mat permuted by both; then matrix2compressed.  rowperm/colperm[new] = old.
 * Rows never touched keep whatever the reference's uninitialised rowPerm held; here they are
 * appended in ascending order (only reachable for matrices with an all-zero row, permanent 0). */
int orc_skip_order(double *mat, int n, int *cptrs, int *rows, double *cvals, int *rptrs, int *cols,
                   double *rvals, int *rowperm, int *colperm) {
  int *degs = (int *)calloc((size_t)n, sizeof(int));
  char *vis = (char *)calloc((size_t)n, 1);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j)
      if (mat[i * n + j] != 0) degs[j]++;
  int placed = 0;
  for (int j = 0; j < n; ++j) {
    int cur = 0, best = 127;   /* INT8_MAX */
    int found = 0;
    for (int l = 0; l < n; ++l)
      if (degs[l] < best) { best = degs[l]; cur = l; found = 1; }
    if (!found) {               /* every remaining column has degree >= 127: take the first unplaced */
      for (int l = 0; l < n; ++l) if (degs[l] != 127) { cur = l; break; }
    }
    degs[cur] = 127;
    colperm[j] = cur;
    for (int l = 0; l < n; ++l) {
      if (mat[l * n + cur] != 0 && !vis[l]) {
        vis[l] = 1;
        rowperm[placed++] = l;
        for (int k = 0; k < n; ++k)
          if (mat[l * n + k] != 0 && degs[k] != 127) degs[k]--;
      }
    }
  }
  for (int l = 0; l < n; ++l) if (!vis[l]) rowperm[placed++] = l;
  double *prev = (double *)malloc((size_t)n * n * sizeof(double));
  memcpy(prev, mat, (size_t)n * n * sizeof(double));
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) mat[r * n + c] = prev[rowperm[r] * n + colperm[c]];
  free(prev);
  free(degs);
  free(vis);
  return orc_matrix2compressed(mat, n, cptrs, rows, cvals, rptrs, cols, rvals);
}

/* gridGraph2compressed (util.h:403-520): bipartite biadjacency (0/1) of the m x n grid graph,
 * nov = m*n/2.  mat must hold nov*nov ints (zeroed here).  Returns nnz, or -1 when both are odd. */
int orc_grid_graph(int m, int n, int *mat) {
  if ((m % 2 == 1) && (n % 2 == 1)) return -1;
  int row, col;
  if (m % 2 == 0) { row = n; col = m; } else { row = m; col = n; }
  const int nov = m * n / 2;
  const int h = col / 2;
  memset(mat, 0, (size_t)nov * nov * sizeof(int));
  for (int i = 0; i < row; ++i) {
    for (int j = 0; j < col; ++j) {
      const int x = i * h + j / 2;
      int nb[4], cnt = 0;
      if (x - h >= 0) nb[cnt++] = x - h;
      if (x + h < nov) nb[cnt++] = x + h;
      if (j % 2 == 0) {
        if (j != 0) nb[cnt++] = x - 1;
        nb[cnt++] = x;
      } else {
        nb[cnt++] = x;
        if (j != col - 1) nb[cnt++] = x + 1;
      }
      const int same = ((i % 2 == 0) && (j % 2 == 0)) || ((i % 2 == 1) && (j % 2 == 1));
      for (int e = 0; e < cnt; ++e) {
        if (same) mat[x * nov + nb[e]] = 1;     /* edges1: (x, neighbour)   util.h:463-465 */
        else      mat[nb[e] * nov + x] = 1;     /* edges2: transposed       util.h:466-468 */
      }
    }
  }
  int nnz = 0;
  for (int i = 0; i < nov * nov; ++i) if (mat[i] > 0) ++nnz;
  return nnz;
}

/* Kasteleyn / Temperley-Fisher: number of perfect matchings (= permanent of the biadjacency
 * matrix) of the m x n grid: prod_{j=1..ceil(m/2)} prod_{k=1..ceil(n/2)} (4cos^2(pi j/(m+1)) + 4cos^2(pi k/(n+1))) */
double orc_kasteleyn(int m, int n) {
  if ((m * n) % 2) return 0.0;
  long double p = 1.0L;
  const long double pi = 3.14159265358979323846264338327950288L;
  for (int j = 1; j <= (m + 1) / 2; ++j)
    for (int k = 1; k <= (n + 1) / 2; ++k) {
      const long double a = cosl(pi * j / (m + 1)), b = cosl(pi * k / (n + 1));
      p *= 4 * a * a + 4 * b * b;   /* an odd dimension's middle index has cos = 0: the factor
                                       degenerates to the other direction's term, as it should */
    }
  return (double)p;
}

/* ------------------------------------------------------------------------------------------
 * SpaRyser partial sum over [lo, hi), lo >= 1: the arithmetic of cpu_perman64_sparse with one
 * thread (gpu_exact_sparse.cu:7-87) == kernel_xshared_coalescing_mshared_sparse (:455-552):
 * incremental product with divide and a zero counter.  x = NW base computed from the dense mat
 * (gpu_exact_sparse.cu:861-871).
 * ---------------------------------------------------------------------------------------- */
double orc_sparyser_range_f64(const double *mat, const int *cptrs, const int *rows, const double *cvals,
                              int n, u64 lo, u64 hi) {
  double x[64];
  for (int j = 0; j < n; ++j) {
    double rs = 0.0;
    for (int k = 0; k < n; ++k) if (mat[j * n + k] != 0) rs += mat[j * n + k];
    x[j] = mat[j * n + (n - 1)] - rs / 2;
  }
  if (lo == 0) lo = 1;
  if (lo >= hi) return 0.0;
  u64 i = lo;
  u64 g = (i - 1) ^ ((i - 1) >> 1);
  for (int k = 0; k < n - 1; ++k)
    if ((g >> k) & 1ull)
      for (int t = cptrs[k]; t < cptrs[k + 1]; ++t) x[rows[t]] += cvals[t];
  double prod = 1.0, acc = 0.0;
  int zero_num = 0;
  for (int j = 0; j < n; ++j) { if (x[j] == 0) zero_num++; else prod *= x[j]; }
  for (; i < hi; ++i) {
    const int k = __builtin_ctzll(i);
    g ^= (1ull << k);
    const double s = ((g >> k) & 1ull) ? 1.0 : -1.0;
    for (int t = cptrs[k]; t < cptrs[k + 1]; ++t) {
      const int r = rows[t];
      if (x[r] == 0) {
        zero_num--;
        x[r] += s * cvals[t];
        prod *= x[r];
      } else {
        prod /= x[r];
        x[r] += s * cvals[t];
        if (x[r] == 0) zero_num++; else prod *= x[r];
      }
    }
    if (zero_num == 0) acc += (i & 1ull) ? -prod : prod;
  }
  return acc;
}

/* ------------------------------------------------------------------------------------------
 * SkipPer partial sum over [lo, hi), lo >= 1: cpu_perman64_skipper with one chunk
 * (gpu_exact_sparse.cu:89-191) == kernel_xshared_coalescing_mshared_skipper (:555-670).
 * *visited receives the number of indices actually evaluated.
 * ---------------------------------------------------------------------------------------- */
double orc_skipper_range_f64(const double *mat, const int *rptrs, const int *cols, const int *cptrs,
                             const int *rows, const double *cvals, int n, u64 lo, u64 hi, u64 *visited) {
  double x[64];
  for (int j = 0; j < n; ++j) {
    double rs = 0.0;
    for (int k = 0; k < n; ++k) if (mat[j * n + k] != 0) rs += mat[j * n + k];
    x[j] = mat[j * n + (n - 1)] - rs / 2;
  }
  if (lo == 0) lo = 1;
  u64 i = lo, prev_gray = 0, seen = 0;
  double acc = 0.0;
  while (i < hi) {
    const u64 gray = i ^ (i >> 1);
    u64 diff = prev_gray ^ gray;
    for (int j = 0; diff; ++j) {
      const u64 onej = 1ull << j;
      if (diff & onej) {
        diff ^= onej;
        if (gray & onej) for (int t = cptrs[j]; t < cptrs[j + 1]; ++t) x[rows[t]] += cvals[t];
        else             for (int t = cptrs[j]; t < cptrs[j + 1]; ++t) x[rows[t]] -= cvals[t];
      }
    }
    prev_gray = gray;
    ++seen;
    int last_zero = -1;
    double prod = 1.0;
    for (int j = n - 1; j >= 0; --j) {
      prod *= x[j];
      if (x[j] == 0) { last_zero = j; break; }
    }
    if (prod != 0) {
      acc += (i & 1ull) ? -prod : prod;
      ++i;
    } else if (last_zero < 0) {
      ++i;   /* product underflowed to zero without a zero row: nothing to skip */
    } else {
      u64 change = ~0ull;
      for (int t = rptrs[last_zero]; t < rptrs[last_zero + 1]; ++t) {
        const int c = cols[t];
        if (c >= 63) continue;
        const u64 start = 1ull << c, period = start << 1;
        u64 ci = start;
        if (i >= start) ci = start + ((i - start) / period + 1) * period;
        if (ci < change) change = ci;
      }
      ++i;
      if (change > i) i = change;
    }
  }
  if (visited) *visited = seen;
  return acc;
}
