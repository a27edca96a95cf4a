/* Structural preprocessing in C: the exact reductions the reference's revised
 * front-end applies BEFORE the exponential kernel, so that the kernel sees a smaller n.
 *
 * Mirrors, for the drop-in boundary:
 *   getMinNnz / checkEmpty                      revised_perman/util.h:1165-1197
 *   d1compress, d2compress                      revised_perman/util.h:1200-1330
 *   d34compress                                 revised_perman/util.h:1333-1407
 *   scalesk + scaleMatrix                       revised_perman/util.h:1445-1593
 *   dulmage_mendehlson                          revised_perman/util.h:143-440 (dead code upstream)
 *   compress_singleton_and_then_recurse,
 *   compress_and_calculate_recursive,
 *   scale_and_calculate                         revised_perman/main.cpp:993-1260
 *
 * Every step picks the same row / column and evaluates the same expressions in the same order as
 * the reference, so the reduced matrices are bit-identical to its -- with one deliberate change:
 * d1compress folds the removed entry into the first matrix row (util.h:1251-1253: "that's where
 * matrix could go out of 0-1 form"); here it is accumulated in a separate scalar factor, so a 0/1
 * matrix stays 0/1 (which keeps SkipPer's exact-zero skipping alive) and perm(original) =
 * factor * perm(reduced).  An entry counts as non-zero when it is != 0; the reference counts > 0 in
 * d1/d2 and != 0 in d34 -- identical on the non-negative matrices it supports.
 */
#define _POSIX_C_SOURCE 200809L
#include "superman_b200.h"
#include "sp_sched.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static void drop_compressed(sp_matrix *m) {
  free(m->cptrs); free(m->rows); free(m->cvals); free(m->rptrs); free(m->cols); free(m->rvals);
  m->cptrs = m->rows = m->rptrs = m->cols = NULL;
  m->cvals = m->rvals = NULL;
  m->nnz = 0;
}

static int row_deg(const double *a, int n, int i) {
  int d = 0;
  for (int j = 0; j < n; ++j) d += a[(size_t)i * n + j] != 0;
  return d;
}
static int col_deg(const double *a, int n, int j) {
  int d = 0;
  for (int i = 0; i < n; ++i) d += a[(size_t)i * n + j] != 0;
  return d;
}

/* getMinNnz (util.h:1181): smallest number of non-zeros over all rows and columns */
int sp_matrix_min_degree(const sp_matrix *m) {
  if (!m || !m->mat) { sp_set_error("null argument"); return SP_EINVAL; }
  const int n = m->nov;
  int best = n;
  for (int i = 0; i < n; ++i) {
    int d = row_deg(m->mat, n, i);
    if (d < best) best = d;
    d = col_deg(m->mat, n, i);
    if (d < best) best = d;
  }
  return best;
}

/* remove row r and column c in place (n -> n-1) */
static void drop_row_col(double *a, int n, int r, int c) {
  size_t w = 0;
  for (int i = 0; i < n; ++i) {
    if (i == r) continue;
    for (int j = 0; j < n; ++j) {
      if (j == c) continue;
      a[w++] = a[(size_t)i * n + j];
    }
  }
}

/* d1compress (util.h:1200-1258): the LAST row with one non-zero wins, else the last such column */
static int step_d1(double *a, int *pn, double *factor) {
  const int n = *pn;
  int r = -1, c = -1;
  for (int i = 0; i < n; ++i) {
    if (row_deg(a, n, i) == 1) r = i;
    if (col_deg(a, n, i) == 1) c = i;
  }
  if (r == -1 && c == -1) return 0;
  if (r != -1) {
    for (int j = 0; j < n; ++j)
      if (a[(size_t)r * n + j] != 0) { c = j; break; }
  } else {
    for (int i = 0; i < n; ++i)
      if (a[(size_t)i * n + c] != 0) { r = i; break; }
  }
  *factor *= a[(size_t)r * n + c];
  drop_row_col(a, n, r, c);
  *pn = n - 1;
  return 1;
}

/* d2compress (util.h:1260-1330): the first index i whose row (preferred) or column has exactly two
 * non-zeros.  Row r with entries in columns p < q: drop row r and column q, column p becomes
 * A[i][p]*A[r][q] + A[i][q]*A[r][p] (expansion along the row + multilinearity); symmetric for a column. */
static int step_d2(double *a, int *pn) {
  const int n = *pn;
  int r = -1, c = -1;
  for (int i = 0; i < n; ++i) {
    if (row_deg(a, n, i) == 2) r = i;
    if (col_deg(a, n, i) == 2) c = i;
    if (r != -1 || c != -1) break;
  }
  if (r == -1 && c == -1) return 0;
  int p = -1, q = -1;
  if (r != -1) {
    for (int j = 0; j < n; ++j)
      if (a[(size_t)r * n + j] != 0) { if (p == -1) p = j; else { q = j; break; } }
    const double arp = a[(size_t)r * n + p], arq = a[(size_t)r * n + q];
    for (int i = 0; i < n; ++i)
      if (i != r) a[(size_t)i * n + p] = a[(size_t)i * n + p] * arq + a[(size_t)i * n + q] * arp;
    drop_row_col(a, n, r, q);
  } else {
    for (int i = 0; i < n; ++i)
      if (a[(size_t)i * n + c] != 0) { if (p == -1) p = i; else { q = i; break; } }
    const double apc = a[(size_t)p * n + c], aqc = a[(size_t)q * n + c];
    for (int j = 0; j < n; ++j)
      if (j != c) a[(size_t)p * n + j] = a[(size_t)p * n + j] * aqc + a[(size_t)q * n + j] * apc;
    drop_row_col(a, n, q, c);
  }
  *pn = n - 1;
  return 2;
}

static int has_empty_line(const double *a, int n) {          /* checkEmpty, util.h:1165 */
  for (int i = 0; i < n; ++i)
    if (row_deg(a, n, i) == 0 || col_deg(a, n, i) == 0) return 1;
  return 0;
}

int sp_matrix_reduce_step(sp_matrix *m, double *factor) {
  if (!m || !m->mat || !factor) { sp_set_error("null argument"); return SP_EINVAL; }
  drop_compressed(m);
  if (m->nov <= 1) return 0;
  int rc = step_d1(m->mat, &m->nov, factor);
  if (!rc) rc = step_d2(m->mat, &m->nov);
  return rc;
}

/* the loop of compress_singleton_and_then_recurse (main.cpp:1058-1090); a row or column without
 * non-zeros ("Matrix is rank deficient! Perman is 0", exit(1) upstream) collapses the matrix to the
 * 1x1 zero matrix with factor 0 instead of ending the process */
int sp_matrix_reduce(sp_matrix *m, double *factor) {
  if (!m || !m->mat || !factor) { sp_set_error("null argument"); return SP_EINVAL; }
  drop_compressed(m);
  *factor = 1.0;
  int removed = 0;
  for (;;) {
    if (has_empty_line(m->mat, m->nov)) {
      removed += m->nov - 1;
      m->nov = 1;
      m->mat[0] = 0.0;
      *factor = 0.0;
      break;
    }
    if (m->nov <= 1) break;
    int rc = step_d1(m->mat, &m->nov, factor);
    if (!rc) rc = step_d2(m->mat, &m->nov);
    if (!rc) break;
    ++removed;
  }
  return removed;
}

/* d34compress (util.h:1333-1407).  Row r (or, when no row qualifies at that index, column r -- the
 * matrix is then transposed, as upstream) with min_deg in {3, 4} non-zeros in columns c0<c1<c2<c3
 * (for three non-zeros c3 is the last zero column of the row):
 *   perm(A) = perm(A1) + perm(A2),
 *   A1 = A without row r and column c1, column c0 := A[r][c0]*A[i][c1] + A[r][c1]*A[i][c0],
 *   A2 = A without row r and column c3, column c2 := A[r][c2]*A[i][c3] + A[r][c3]*A[i][c2]. */
int sp_matrix_split34(sp_matrix *m, int min_deg, sp_matrix *second) {
  if (!m || !m->mat || !second) { sp_set_error("null argument"); return SP_EINVAL; }
  if (min_deg != 3 && min_deg != 4) { sp_set_error("split34: degree %d is not 3 or 4", min_deg); return SP_EINVAL; }
  const int n = m->nov;
  if (n < 5) { sp_set_error("split34 needs n >= 5 (got %d)", n); return SP_ELIMIT; }
  double *a = m->mat;
  int r = -1, c = -1;
  for (int i = 0; i < n; ++i) {
    if (row_deg(a, n, i) == min_deg) r = i;
    if (col_deg(a, n, i) == min_deg) c = i;
    if (r != -1 || c != -1) break;
  }
  if (r == -1 && c == -1) return 0;
  double *t = (double *)malloc((size_t)n * n * sizeof(double));
  double *b = (double *)calloc((size_t)n * n, sizeof(double));
  if (!t || !b) { free(t); free(b); sp_set_error("out of memory"); return SP_ENOMEM; }
  if (r == -1) {
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) t[(size_t)j * n + i] = a[(size_t)i * n + j];
    r = c;
  } else {
    memcpy(t, a, (size_t)n * n * sizeof(double));
  }
  int nb[4] = {-1, -1, -1, -1}, k = 0, zeroloc = -1;
  for (int j = 0; j < n; ++j) {
    if (t[(size_t)r * n + j] != 0) nb[k++] = j; else zeroloc = j;
  }
  if (nb[3] == -1) nb[3] = zeroloc;
  drop_compressed(m);
  memset(a, 0, (size_t)n * n * sizeof(double));
  const int n1 = n - 1;
  const double *tr = t + (size_t)r * n;
  for (int i = 0; i < n; ++i) {
    if (i == r) continue;
    const int il = i - (i > r);
    const double *ti = t + (size_t)i * n;
    for (int j = 0; j < n; ++j) {
      if (j != nb[1]) {
        const int jl = j - (j > nb[1]);
        a[(size_t)il * n1 + jl] = (j != nb[0]) ? ti[j] : tr[nb[0]] * ti[nb[1]] + tr[nb[1]] * ti[nb[0]];
      }
      if (j != nb[3]) {
        const int jl = j - (j > nb[3]);
        b[(size_t)il * n1 + jl] = (j != nb[2]) ? ti[j] : tr[nb[2]] * ti[nb[3]] + tr[nb[3]] * ti[nb[2]];
      }
    }
  }
  free(t);
  m->nov = n1;
  memset(second, 0, sizeof(*second));
  second->nov = n1;
  second->type = m->type;
  second->mat = b;
  return 1;
}

/* scalesk + scaleMatrix (util.h:1445-1593): Sinkhorn-Knopp sweeps (columns, then rows) until the mean
 * column sum and the mean row sum are both within 10 of `threshold`, then mat[i][j] *= rv[i] * cv[j].
 * perm(original) = perm(scaled) / prod(cv) / prod(rv).  Sums run in CCS / CRS order (ascending row /
 * column index) with the products associated as upstream: (val*cv)*rv per column, (val*rv)*cv per
 * row.  Upstream loops forever when the sweeps do not converge; here at most 1000 sweeps. */
int sp_matrix_scale(sp_matrix *m, double threshold, double *rv, double *cv) {
  if (!m || !m->mat || !rv || !cv) { sp_set_error("null argument"); return SP_EINVAL; }
  if (!(threshold > 0)) { sp_set_error("scaling threshold must be positive"); return SP_EINVAL; }
  const int n = m->nov;
  double *a = m->mat;
  for (int i = 0; i < n; ++i) rv[i] = cv[i] = 1.0;
  double max_error = 100;
  int sweeps = 0;
  while (max_error > 10.0 && sweeps < 1000) {
    for (int j = 0; j < n; ++j) {
      double sum = 0;
      int any = 0;
      for (int i = 0; i < n; ++i) {
        const double v = a[(size_t)i * n + j];
        if (v > 0) { sum += v * cv[j] * rv[i]; any = 1; }
      }
      if (any) cv[j] = threshold / sum;
    }
    for (int i = 0; i < n; ++i) {
      double sum = 0;
      int any = 0;
      for (int j = 0; j < n; ++j) {
        const double v = a[(size_t)i * n + j];
        if (v > 0) { sum += v * rv[i] * cv[j]; any = 1; }
      }
      if (any) rv[i] = threshold / sum;
    }
    double colsum = 0, rowsum = 0;
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) {
        const double v = a[(size_t)i * n + j];
        if (v > 0) colsum += v * cv[j] * rv[i];
      }
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        const double v = a[(size_t)i * n + j];
        if (v > 0) rowsum += v * rv[i] * cv[j];
      }
    const double e1 = fabs(threshold - colsum / n), e2 = fabs(threshold - rowsum / n);
    max_error = e1 > e2 ? e1 : e2;
    ++sweeps;
  }
  drop_compressed(m);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) a[(size_t)i * n + j] *= rv[i];
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) a[(size_t)i * n + j] *= cv[j];
  return sweeps;
}

/* Sinkhorn-Knopp to convergence: like sp_matrix_scale, but sweeps until EVERY column sum is within
 * 0.1 % of `threshold` right after the rows were normalised (at most 1000 sweeps).  This is what the
 * compressed driver uses: upstream's stopping rule looks at the MEAN row / column sum only, which
 * the first sweep already satisfies, and one sweep is not enough for the matrices the degree
 * compression produces -- merged columns carry products of entries, the Ryser sum over such a
 * matrix cancels catastrophically (measured on 24x24 banded matrices reduced to 8x8 leaves: relative
 * error 0.29 in FP64 and 1e-3 in long double after one sweep, 1e-15 in FP64 once converged; see
 * tests/test_host_reduce.py).  Call sp_matrix_dm first: Sinkhorn converges linearly only on a matrix
 * with total support.  Returns the number of sweeps. */
int sp_matrix_balance(sp_matrix *m, double threshold, double *rv, double *cv) {
  if (!m || !m->mat || !rv || !cv) { sp_set_error("null argument"); return SP_EINVAL; }
  if (!(threshold > 0)) { sp_set_error("scaling threshold must be positive"); return SP_EINVAL; }
  const int n = m->nov;
  double *a = m->mat;
  for (int i = 0; i < n; ++i) rv[i] = cv[i] = 1.0;
  int sweeps = 0;
  for (; sweeps < 1000;) {
    for (int j = 0; j < n; ++j) {
      double sum = 0;
      for (int i = 0; i < n; ++i) sum += a[(size_t)i * n + j] * rv[i];
      if (sum > 0) cv[j] = threshold / sum;
    }
    for (int i = 0; i < n; ++i) {
      double sum = 0;
      for (int j = 0; j < n; ++j) sum += a[(size_t)i * n + j] * cv[j];
      if (sum > 0) rv[i] = threshold / sum;
    }
    ++sweeps;
    double worst = 0;
    for (int j = 0; j < n; ++j) {
      double sum = 0;
      for (int i = 0; i < n; ++i) sum += a[(size_t)i * n + j] * rv[i];
      if (sum > 0) {
        const double e = fabs(sum * cv[j] - threshold);
        if (e > worst) worst = e;
      }
    }
    if (worst <= 1e-3 * threshold) break;
  }
  drop_compressed(m);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) a[(size_t)i * n + j] = a[(size_t)i * n + j] * rv[i] * cv[j];
  return sweeps;
}

/* Dulmage-Mendelsohn fine decomposition (util.h:309-440): with a perfect matching M, entry (i, j)
 * lies on some perfect matching iff row i and the row matched to column j are in the same strongly
 * connected component of the digraph { i -> M(j) : A[i][j] != 0, (i, j) not in M }.  Entries that
 * lie on no perfect matching contribute nothing to the permanent and are erased -- which feeds the
 * degree compression and the sparse kernels.  (Upstream compares component[i] with component[j]
 * using the column index j directly, which is only right when the matching is the diagonal; the
 * routine is never called there.)  *matching receives the size of a maximum matching; when it is
 * < nov the permanent is 0 and nothing is erased.  Returns the number of erased entries. */
static int dm_augment(const double *a, int n, int i, int *seen, int *col_row, int stamp) {
  for (int j = 0; j < n; ++j) {
    if (a[(size_t)i * n + j] == 0 || seen[j] == stamp) continue;
    seen[j] = stamp;
    if (col_row[j] < 0 || dm_augment(a, n, col_row[j], seen, col_row, stamp)) { col_row[j] = i; return 1; }
  }
  return 0;
}

int sp_matrix_dm(sp_matrix *m, int *matching) {
  if (!m || !m->mat) { sp_set_error("null argument"); return SP_EINVAL; }
  const int n = m->nov;
  double *a = m->mat;
  int *buf = (int *)malloc((size_t)n * 9 * sizeof(int));
  if (!buf) { sp_set_error("out of memory"); return SP_ENOMEM; }
  int *seen = buf, *col_row = buf + n, *index = buf + 2 * n, *low = buf + 3 * n, *comp = buf + 4 * n,
      *stack = buf + 5 * n, *call_i = buf + 6 * n, *call_j = buf + 7 * n, *onstack = buf + 8 * n;
  for (int j = 0; j < n; ++j) { seen[j] = -1; col_row[j] = -1; }
  int matched = 0;
  for (int i = 0; i < n; ++i) matched += dm_augment(a, n, i, seen, col_row, i);
  if (matching) *matching = matched;
  if (matched < n) { free(buf); return 0; }
  /* Tarjan, iterative; successor of row i through column j is col_row[j] */
  for (int i = 0; i < n; ++i) { index[i] = -1; comp[i] = -1; onstack[i] = 0; }
  int next_index = 0, ncomp = 0, sp = 0;
  for (int root = 0; root < n; ++root) {
    if (index[root] != -1) continue;
    int depth = 0;
    call_i[0] = root; call_j[0] = 0;
    index[root] = low[root] = next_index++;
    stack[sp++] = root; onstack[root] = 1;
    while (depth >= 0) {
      const int i = call_i[depth];
      int advanced = 0;
      while (call_j[depth] < n) {
        const int j = call_j[depth]++;
        if (a[(size_t)i * n + j] == 0 || col_row[j] == i) continue;
        const int w = col_row[j];
        if (index[w] == -1) {
          index[w] = low[w] = next_index++;
          stack[sp++] = w; onstack[w] = 1;
          ++depth;
          call_i[depth] = w; call_j[depth] = 0;
          advanced = 1;
          break;
        }
        if (onstack[w] && index[w] < low[i]) low[i] = index[w];
      }
      if (advanced) continue;
      if (low[i] == index[i]) {
        int w;
        do { w = stack[--sp]; onstack[w] = 0; comp[w] = ncomp; } while (w != i);
        ++ncomp;
      }
      --depth;
      if (depth >= 0 && low[i] < low[call_i[depth]]) low[call_i[depth]] = low[i];
    }
  }
  int erased = 0;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j)
      if (a[(size_t)i * n + j] != 0 && comp[i] != comp[col_row[j]]) { a[(size_t)i * n + j] = 0; ++erased; }
  free(buf);
  if (erased) drop_compressed(m);
  return erased;
}

/* ---- compress_singleton_and_then_recurse / compress_and_calculate_recursive / scale_and_calculate
 * (main.cpp:993-1260) on the GPU engine ----------------------------------------------------------
 * The recursion only PREPARES leaves (reduce, split, total support, balance, -r ordering) and
 * queues them with the product of the degree-1 factors along their path; the permanent is
 *   sum over leaves, in the order the recursion meets them, of  coeff * perm(leaf) / prod(rv cv).
 * Queued leaves are independent: when at least as many leaves are pending as the id has devices
 * (one device for the single-GPU ids, gpu_num for -p5 / -p6 / -p8), host threads -- two per device,
 * so one leaf's plan upload and result read-back overlap the other's kernel -- take whole leaves
 * from a shared counter (no exchange, one double back per leaf); otherwise every leaf goes through
 * the id's own entry point, which splits its Gray range over the devices.  The sum order is fixed,
 * so the result does not depend on which device computed which leaf. */
#define SP_LEAF_BATCH 1024

typedef struct leaf {
  sp_matrix m;                 /* owned; carries CRS / CCS for the sparse ids */
  double coeff;                /* product of the degree-1 factors on the path to this leaf */
  double rv[64], cv[64];       /* Sinkhorn factors when scaled */
  int scaled;
  int slot;                    /* host thread that computed it in a parallel flush (device = slot % devices), else -1 */
  double value;                /* coeff * perm(leaf before scaling) */
  sp_stats st;
  int rc;
  char err[200];
} leaf;

typedef struct reduce_ctx {
  int sparse, preprocessing, algo_id, gpu_num, threads, leaf_nov;
  int multi;                   /* the id partitions over several devices */
  int first_device;
  double threshold;
  sp_stats total;
#define SP_LEAF_THREADS_PER_DEVICE 2
  double seq_ms, thr_ms[SP_MAX_DEVICES * SP_LEAF_THREADS_PER_DEVICE];   /* sums of CUDA-event times */
  unsigned long long thr_units[SP_MAX_DEVICES * SP_LEAF_THREADS_PER_DEVICE];
  int par_devices;             /* devices of the parallel flushes */
  int leaves, failed, altered;
  char err[200];
  leaf *pend[SP_LEAF_BATCH];
  int npend;
  int next;                    /* shared leaf counter of a parallel flush */
  double sum;
} reduce_ctx;

static void fail_ctx(reduce_ctx *cx, int code, const char *msg) {
  if (cx->failed) return;
  cx->failed = code;
  snprintf(cx->err, sizeof(cx->err), "%s", msg);
}

/* slot < 0: through the id's own entry point (which may split the leaf over the devices);
 * slot >= 0: the whole leaf on device first_device + slot % devices */
static void run_leaf(const reduce_ctx *cx, leaf *lf, int slot, int devices) {
  const int n = lf->m.nov;
  double v;
  memset(&lf->st, 0, sizeof(lf->st));
  lf->rc = SP_OK;
  lf->slot = slot;
  if (n == 1) {
    v = lf->m.mat[0];
  } else if (slot < 0) {
    if (!cx->sparse)
      v = sp_dense_ryser(lf->m.mat, n, cx->algo_id, cx->gpu_num, 0, cx->threads, &lf->st);
    else if (cx->algo_id == 7 || cx->algo_id == 8)
      v = sp_skipper(lf->m.mat, lf->m.rptrs, lf->m.cols, lf->m.cptrs, lf->m.rows, lf->m.cvals, n, cx->algo_id,
                     cx->gpu_num, 0, cx->threads, &lf->st);
    else
      v = sp_sparse_ryser(lf->m.mat, lf->m.cptrs, lf->m.rows, lf->m.cvals, n, cx->algo_id, cx->gpu_num, 0,
                          cx->threads, &lf->st);
  } else {
    const long long end = 1ll << (n - 1);
    const int device = cx->first_device + slot % devices;
    if (!cx->sparse)
      v = sp_dense_ryser_range(lf->m.mat, n, device, 0, end, &lf->st);
    else
      /* a balanced leaf has no exact zeros left to skip (sp_api.c: some_row_can_cancel) */
      v = sp_sparse_ryser_range(lf->m.mat, lf->m.cptrs, lf->m.rows, lf->m.cvals, n,
                                (cx->algo_id == 7 || cx->algo_id == 8) && !lf->scaled, device, 0, end, &lf->st);
    v *= sp_nw_factor(n);
  }
  if (isnan(v) && lf->st.error) {
    lf->rc = lf->st.error;
    snprintf(lf->err, sizeof(lf->err), "%s", sp_last_error());
    return;
  }
  /* undo the scaling (main.cpp:1143-1149) in long double: the factors of a badly scaled leaf can be
   * far apart (rv huge, cv tiny) and must not overflow on the way to a representable result */
  long double lv = (long double)v;
  if (lf->scaled)
    for (int i = 0; i < n; ++i) { lv /= (long double)lf->cv[i]; lv /= (long double)lf->rv[i]; }
  lf->value = (double)((long double)lf->coeff * lv);
}

typedef struct flush_arg { reduce_ctx *cx; int slot; int devices; } flush_arg;

static void *flush_worker(void *p) {
  flush_arg *fa = (flush_arg *)p;
  reduce_ctx *cx = fa->cx;
  for (;;) {
    const int i = __atomic_fetch_add(&cx->next, 1, __ATOMIC_RELAXED);
    if (i >= cx->npend) break;
    run_leaf(cx, cx->pend[i], fa->slot, fa->devices);
  }
  return NULL;
}

static void flush_leaves(reduce_ctx *cx) {
  if (cx->npend == 0) return;
  int devices = 1;
  if (cx->multi && !cx->failed) {
    const int visible = sp_device_count() - cx->first_device;
    devices = cx->gpu_num < visible ? cx->gpu_num : visible;
    if (devices > SP_MAX_DEVICES) devices = SP_MAX_DEVICES;
    if (devices < 1) devices = 1;
  }
  if (cx->failed) {
    /* an earlier leaf failed: nothing more is computed */
  } else if (cx->npend >= 2 && cx->npend >= devices) {
    pthread_t th[SP_MAX_DEVICES * SP_LEAF_THREADS_PER_DEVICE];
    flush_arg fa[SP_MAX_DEVICES * SP_LEAF_THREADS_PER_DEVICE];
    int started[SP_MAX_DEVICES * SP_LEAF_THREADS_PER_DEVICE] = {0};
    int nthreads = devices * SP_LEAF_THREADS_PER_DEVICE;
    if (nthreads > cx->npend) nthreads = cx->npend;
    cx->next = 0;
    for (int t = 1; t < nthreads; ++t) {
      fa[t].cx = cx; fa[t].slot = t; fa[t].devices = devices;   /* thread t drives device t % devices */
      started[t] = pthread_create(&th[t], NULL, flush_worker, &fa[t]) == 0;
    }
    fa[0].cx = cx; fa[0].slot = 0; fa[0].devices = devices;
    flush_worker(&fa[0]);                        /* the caller's thread works too */
    for (int t = 1; t < nthreads; ++t)
      if (started[t]) pthread_join(th[t], NULL);
    if (devices > cx->total.devices) cx->total.devices = devices;
    if (devices > cx->par_devices) cx->par_devices = devices;
  } else {
    for (int i = 0; i < cx->npend; ++i) {
      run_leaf(cx, cx->pend[i], -1, 1);
      if (cx->pend[i]->rc != SP_OK) { fail_ctx(cx, cx->pend[i]->rc, cx->pend[i]->err); break; }
    }
  }
  for (int i = 0; i < cx->npend; ++i) {
    leaf *lf = cx->pend[i];
    if (lf->rc != SP_OK) fail_ctx(cx, lf->rc, lf->err);
    if (!cx->failed) {
      cx->sum += lf->value;
      cx->total.units += lf->st.units;
      cx->total.visited += lf->st.visited;
      cx->total.launches += lf->st.launches;
      if (lf->st.devices > cx->total.devices) cx->total.devices = lf->st.devices;
      cx->total.path = lf->st.path;
      cx->total.tile_log2 = lf->st.tile_log2;
      if (lf->slot >= 0) {
        cx->thr_ms[lf->slot] += lf->st.kernel_ms;
        cx->thr_units[lf->slot] += lf->st.units;
      } else {
        cx->seq_ms += lf->st.kernel_ms;
      }
      ++cx->leaves;
    }
    sp_matrix_free(&lf->m);
    free(lf);
  }
  cx->npend = 0;
}

/* consumes m: total support, balancing, ordering; queues the leaf */
static void queue_leaf(reduce_ctx *cx, sp_matrix *m, double coeff) {
  const int n = m->nov;
  leaf *lf = NULL;
  int rc;
  if (cx->failed) goto drop;
  if (n > 64) { sp_set_error("exact paths support n <= 64 (got %d after compression)", n); fail_ctx(cx, SP_ELIMIT, sp_last_error()); goto drop; }
  lf = (leaf *)calloc(1, sizeof(leaf));
  if (!lf) { fail_ctx(cx, SP_ENOMEM, "out of memory"); goto drop; }
  /* threshold 0 = automatic: scale when the compression changed the matrix (see sp_matrix_balance for
   * why that is necessary); untouched matrices are not scaled -- SkipPer's exact-zero skipping depends
   * on the values.  Before scaling, entries on no perfect matching are erased (exact), which gives
   * the matrix total support; a leaf without a perfect matching has permanent 0 and needs no kernel. */
  const double threshold = cx->threshold > 0 ? cx->threshold : (cx->threshold == 0 && cx->altered) ? 1.0 : -1.0;
  if (threshold > 0 && n > 1) {
    int matching = 0;
    rc = sp_matrix_dm(m, &matching);
    if (rc < 0) { fail_ctx(cx, rc, sp_last_error()); goto drop; }
    if (matching < n) { ++cx->leaves; goto drop; }
    rc = sp_matrix_balance(m, threshold, lf->rv, lf->cv);
    if (rc < 0) { fail_ctx(cx, rc, sp_last_error()); goto drop; }
    lf->scaled = 1;
  }
  if (cx->sparse && n > 1) {
    rc = sp_matrix_compress(m, cx->preprocessing);
    if (rc != SP_OK) { fail_ctx(cx, rc, sp_last_error()); goto drop; }
  }
  lf->m = *m;                                   /* ownership moves into the leaf */
  memset(m, 0, sizeof(*m));
  lf->coeff = coeff;
  cx->pend[cx->npend++] = lf;
  if (cx->npend == SP_LEAF_BATCH) flush_leaves(cx);
  return;
drop:
  free(lf);
  sp_matrix_free(m);
}

/* consumes m */
static void recurse(reduce_ctx *cx, sp_matrix *m, double coeff) {
  for (;;) {
    if (cx->failed) break;
    const int mind = sp_matrix_min_degree(m);
    if (mind == 0) break;                       /* an empty row or column: this branch adds 0 */
    if (!(mind < 5 && m->nov > cx->leaf_nov)) { queue_leaf(cx, m, coeff); return; }
    if (mind <= 2) {
      /* one d1 / d2 step (main.cpp:1010-1028); the d1 entry goes to the coefficient, not into row 0 */
      int rc = (mind == 1) ? step_d1(m->mat, &m->nov, &coeff) : step_d2(m->mat, &m->nov);
      if (rc <= 0) { fail_ctx(cx, SP_EINVAL, "degree compression found no candidate"); break; }
      cx->altered = 1;
      continue;
    }
    sp_matrix second;
    int rc = sp_matrix_split34(m, mind, &second);
    if (rc <= 0) { fail_ctx(cx, rc < 0 ? rc : SP_EINVAL, rc < 0 ? sp_last_error() : "d34 split found no candidate"); break; }
    cx->altered = 1;
    sp_matrix first = *m;                       /* ownership moves into the two recursive calls */
    memset(m, 0, sizeof(*m));
    recurse(cx, &first, coeff);
    recurse(cx, &second, coeff);
    return;
  }
  sp_matrix_free(m);
}

double sp_permanent_compressed(const double *mat, int nov, int sparse, int preprocessing, int algo_id, int gpu_num,
                               int threads, double scaling_threshold, int leaf_nov, sp_stats *stats) {
  const double t0 = sp_now_ms();
  if (stats) memset(stats, 0, sizeof(*stats));
  /* a matrix that compresses all the way (triangular, diagonal, structurally singular) needs no
   * kernel, but this is a GPU entry point like the others: without a device it fails, it does not
   * quietly become a CPU implementation */
  if (sp_device_count() <= 0) {
    sp_set_error("no CUDA device (libsuperman_b200 has no CPU fallback)");
    if (stats) stats->error = SP_ENODEV;
    return NAN;
  }
  reduce_ctx *cx = (reduce_ctx *)calloc(1, sizeof(reduce_ctx));
  if (!cx) { sp_set_error("out of memory"); if (stats) stats->error = SP_ENOMEM; return NAN; }
  cx->sparse = sparse; cx->preprocessing = preprocessing; cx->algo_id = algo_id;
  cx->gpu_num = gpu_num < 1 ? 1 : gpu_num;
  cx->threads = threads; cx->threshold = scaling_threshold;
  cx->multi = sparse ? (algo_id == 5 || algo_id == 6 || algo_id == 8) : (algo_id == 5 || algo_id == 6);
  cx->first_device = sp_first_device();
  /* `densemat->nov > 30`, main.cpp:1008; leaf_nov < 0: no compression at all (scaling only) */
  cx->leaf_nov = leaf_nov > 0 ? leaf_nov : leaf_nov == 0 ? 30 : SP_MAX_NOV;
  sp_matrix m;
  int rc = sp_matrix_from_dense(mat, nov, &m);
  if (rc != SP_OK) { free(cx); if (stats) stats->error = rc; return NAN; }
  /* With scaling on, balance the matrix once BEFORE compressing it, as upstream does (-u then -o,
   * main.cpp:1637-1641): the d2 / d34 merges multiply entries, which must not overflow or drown on a
   * badly scaled input.  Automatic mode (threshold 0) does it only when a first step applies -- an
   * untouched matrix stays as it is. */
  double *pre_rv = NULL, *pre_cv = NULL;
  int structurally_zero = 0;
  if (leaf_nov >= 0 && nov > 1) {
    const int mind = sp_matrix_min_degree(&m);
    const int will_alter = mind <= 2 || (mind < 5 && nov > cx->leaf_nov);
    const double thr0 = scaling_threshold > 0 ? scaling_threshold : (scaling_threshold == 0 && will_alter) ? 1.0 : -1.0;
    if (thr0 > 0 && mind > 0) {
      int matching = 0;
      rc = sp_matrix_dm(&m, &matching);
      if (rc >= 0 && matching < nov) structurally_zero = 1;
      if (rc >= 0 && !structurally_zero) {
        pre_rv = (double *)malloc(2 * (size_t)nov * sizeof(double));
        if (!pre_rv) { sp_set_error("out of memory"); rc = SP_ENOMEM; }
        else { pre_cv = pre_rv + nov; rc = sp_matrix_balance(&m, thr0, pre_rv, pre_cv); }
      }
      if (rc < 0) { free(pre_rv); sp_matrix_free(&m); free(cx); if (stats) stats->error = rc; return NAN; }
    }
  }
  double factor = 1.0;
  rc = (leaf_nov < 0 || structurally_zero) ? 0 : sp_matrix_reduce(&m, &factor);
  if (rc < 0) { free(pre_rv); sp_matrix_free(&m); free(cx); if (stats) stats->error = rc; return NAN; }
  cx->altered = rc > 0;
  if (factor == 0.0 || structurally_zero) {
    sp_matrix_free(&m);
  } else {
    recurse(cx, &m, factor);
    flush_leaves(cx);
  }
  const int failed = cx->failed;
  double perman = cx->sum;
  if (pre_rv) {
    long double t = (long double)perman;
    for (int i = 0; i < nov; ++i) { t /= (long double)pre_cv[i]; t /= (long double)pre_rv[i]; }
    perman = (double)t;
    free(pre_rv);
  }
  if (failed) sp_set_error("%s", cx->err);
  if (stats) {
    *stats = cx->total;
    /* a device's two host threads overlap their kernels: its time is the longer of the two sums */
    double par = 0.0;
    for (int t = 0; cx->par_devices > 0 && t < cx->par_devices * SP_LEAF_THREADS_PER_DEVICE; ++t) {
      const int d = t % cx->par_devices;
      if (cx->thr_ms[t] > stats->device_ms[d]) stats->device_ms[d] = cx->thr_ms[t];
      stats->device_units[d] += cx->thr_units[t];
      if (cx->thr_ms[t] > par) par = cx->thr_ms[t];
    }
    stats->kernel_ms = cx->seq_ms + par;
    stats->chunks = cx->leaves;
    stats->wall_ms = sp_now_ms() - t0;
    stats->error = failed;
  }
  free(cx);
  return failed ? NAN : perman;
}
