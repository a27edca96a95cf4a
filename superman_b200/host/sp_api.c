/* C-ABI entry points of libsuperman_b200.so: the host halves of the reference's gpu_perman64_*
 * wrappers (NW preamble, transpose, partition choice, final factor), in C, on top of the device
 * layer (superman_b200_device.h) and the chunk scheduler (sp_sched.h). */
#define _POSIX_C_SOURCE 200809L
#include "superman_b200.h"
#include "superman_b200_device.h"
#include "sp_sched.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

const char *sp_version(void) { return "superman_b200 0.1 (sm_100a)"; }

/* first device used by the permanent entry points (the revised front-end's -l device id,
 * revised_perman/main.cpp:1443-1449); multi-GPU ids use first .. first+gpu_num-1 */
static int g_first_device = 0;
int sp_set_first_device(int device) {
  if (device < 0) { sp_set_error("negative device id"); return SP_EINVAL; }
  g_first_device = device;
  return SP_OK;
}

int sp_first_device(void) { return g_first_device; }

/* the revised front-end's -q (flags.calculation_quad, revised_perman/flags.h:61-64): dense exact calls made
 * after sp_set_precision(SP_PRECISION_QUAD) run the double-double kernel */
int sp_set_precision(int precision) {
  if (precision != SP_PRECISION_DOUBLE && precision != SP_PRECISION_QUAD) { sp_set_error("unknown precision %d", precision); return SP_EINVAL; }
  spd_set_quad(precision == SP_PRECISION_QUAD);
  return SP_OK;
}

int sp_device_count(void) {
  int n = spd_device_count();
  if (n < 0) sp_set_error("%s", spd_last_error());
  return n;
}

double sp_int_peak(int device, int millis) {
  double r = spd_int_peak_instr_per_s(device, millis);
  if (r < 0) sp_set_error("%s", spd_last_error());
  return r;
}

double sp_nw_factor(int nov) { return (double)(4 * (nov & 1) - 2); }

double sp_fp64_peak(int device, int millis) {
  double r = spd_fp64_peak_instr_per_s(device, millis);
  if (r < 0) sp_set_error("%s", spd_last_error());
  return r;
}

/* Creates the CUDA contexts and the pooled streams / buffers of devices 0..gpu_num-1 (in parallel,
 * one thread per device) so that a following call starts from warm devices: primary-context
 * creation costs ~0.15 s per GPU and is not part of any algorithm. */
typedef struct warm_job { int dummy; } warm_job;
static int warm_open(const void *job, int device, void **plan) {
  (void)job;
  static const double one[4] = {1.0, 1.0, 1.0, 1.0}, xb[2] = {0.5, 0.5};
  return spd_dense_plan_create(device, one, xb, 2, (spd_dense_plan **)plan);
}
static int warm_launch(void *plan, unsigned long long lo, unsigned long long hi) {
  return spd_dense_plan_launch((spd_dense_plan *)plan, lo, hi);
}
static int warm_wait(void *plan, double *sum, spd_run_info *info) {
  return spd_dense_plan_wait((spd_dense_plan *)plan, sum, info);
}
static void warm_close(void *plan) { spd_dense_plan_destroy((spd_dense_plan *)plan); }

int sp_warmup(int gpu_num) {
  static const sp_job_ops ops = {warm_open, warm_launch, warm_wait, warm_close};
  warm_job job = {0};
  double sum = 0.0;
  sp_stats st;
  memset(&st, 0, sizeof(st));
  if (gpu_num < 1) gpu_num = 1;
  /* every device opens two plans on a 2x2 matrix: creates the context, the two pooled lanes (stream, events,
   * pinned slot, arena, partial buffers) the chunk-queue paths use, and the parked worker threads */
  return sp_sched_run(&ops, &job, SP_SCHED_PREPARE, gpu_num, g_first_device, 0ull, 0ull, 0, 0ull, &sum, &st);
}

static void stats_clear(sp_stats *st) {
  if (st) memset(st, 0, sizeof(*st));
}

static double fail(sp_stats *st, int code) {
  if (st) st->error = code;
  return NAN;
}

/* ---- Nijenhuis-Wilf preamble (gpu_exact_dense.cu:642-662) ------------------------------------ */
/* x[j] = A[j][n-1] - rowsum_j / 2, and the transposed matrix mat_t[k*n+j] = A[j][k]. */
static void nw_preamble(const double *mat, int nov, double *x, double *mat_t) {
  for (int j = 0; j < nov; ++j) {
    double rs = 0.0;
    for (int k = 0; k < nov; ++k) rs += mat[j * nov + k];
    x[j] = mat[j * nov + (nov - 1)] - rs / 2;
  }
  if (mat_t) {
    for (int i = 0; i < nov; ++i)
      for (int j = 0; j < nov; ++j) mat_t[i * nov + j] = mat[j * nov + i];
  }
}

/* ---- dense job for the scheduler ------------------------------------------------------------- */
typedef struct dense_job {
  const double *mat_t;
  const double *x;
  int nov;
} dense_job;

static int dense_open(const void *job, int device, void **plan) {
  const dense_job *j = (const dense_job *)job;
  return spd_dense_plan_create(device, j->mat_t, j->x, j->nov, (spd_dense_plan **)plan);
}
static int dense_launch(void *plan, unsigned long long lo, unsigned long long hi) {
  return spd_dense_plan_launch((spd_dense_plan *)plan, lo, hi);
}
static int dense_wait(void *plan, double *sum, spd_run_info *info) {
  return spd_dense_plan_wait((spd_dense_plan *)plan, sum, info);
}
static void dense_close(void *plan) { spd_dense_plan_destroy((spd_dense_plan *)plan); }

static const sp_job_ops g_dense_ops = {dense_open, dense_launch, dense_wait, dense_close};

/* number of chunks of the dynamic paths: the reference uses 2^(nov-29) (dense,
 * gpu_exact_dense.cu:786-793) or 2^(nov-30) (sparse, gpu_exact_sparse.cu:1005-1008), i.e. chunks of
 * about 2^28 / 2^29 indices; we keep that rule but never hand a device fewer than 8 chunks to
 * balance, as long as a chunk still has 2^22 indices. */
unsigned long long sp_dynamic_chunks(int nov, int ref_base, int gpu_num) {
  unsigned long long chunks = 1;
  for (int i = ref_base; i < nov; ++i) chunks *= 2;
  const unsigned long long want = 8ull * (unsigned long long)(gpu_num > 0 ? gpu_num : 1);
  const unsigned long long total = 1ull << (nov - 1);
  while (chunks < want && total / (chunks * 2) >= (1ull << 22)) chunks *= 2;
  return chunks;
}

/* Dense Ryser costs the same for every index, so the queue has nothing to balance but differences between
 * the devices themselves: the reference's 2^(nov-29) chunks (128 at n = 36, 2048 at n = 40) only add a launch,
 * a reduction and an 8-byte copy per chunk (measured in round 1: -p6 1.4-2 % behind -p5).  Four chunks per
 * device -- two in flight, two to even out -- keep the dynamic path at the static path's speed. */
static unsigned long long dense_dynamic_chunks(int nov, int gpu_num) {
  unsigned long long chunks = sp_dynamic_chunks(nov, 29, gpu_num);
  const unsigned long long cap = 4ull * (unsigned long long)(gpu_num > 0 ? gpu_num : 1);
  if (chunks > cap) chunks = cap;
  return chunks;
}

static int check_dense_args(const double *mat, int nov, sp_stats *st) {
  if (!mat) { sp_set_error("mat is NULL"); if (st) st->error = SP_EINVAL; return SP_EINVAL; }
  if (nov < 1 || nov > 64) {
    sp_set_error("dense Ryser supports 1 <= n <= 64 (got %d)", nov);
    if (st) st->error = SP_ELIMIT;
    return SP_ELIMIT;
  }
  return SP_OK;
}

double sp_dense_ryser(const double *mat, int nov, int algo_id, int gpu_num, int use_cpu, int threads,
                      sp_stats *stats) {
  (void)use_cpu; (void)threads;
  const double t0 = sp_now_ms();
  stats_clear(stats);
  if (check_dense_args(mat, nov, stats) != SP_OK) return NAN;
  int mode = SP_SCHED_STATIC;
  switch (algo_id) {
    case 0: case 1: case 2: case 3: case 4: gpu_num = 1; break;   /* main.cu:34-58 */
    case 5: break;                                                 /* main.cu:59-63 */
    case 6: mode = SP_SCHED_DYNAMIC; break;                        /* main.cu:64-68 */
    default:
      sp_set_error("Unknown Algorithm ID");
      return fail(stats, SP_EALGO);
  }
  if (gpu_num < 1) gpu_num = 1;
  if (nov == 1) {
    if (spd_device_count() <= 0) { sp_set_error("no CUDA device: %s", spd_last_error()); return fail(stats, SP_ENODEV); }
    if (stats) stats->wall_ms = sp_now_ms() - t0;
    return mat[0];
  }
  double x[64];
  double *mat_t = (double *)malloc((size_t)nov * nov * sizeof(double));
  if (!mat_t) { sp_set_error("out of memory"); return fail(stats, SP_ENOMEM); }
  nw_preamble(mat, nov, x, mat_t);
  dense_job job = {mat_t, x, nov};
  const unsigned long long end = 1ull << (nov - 1);
  /* never more devices than 2^16-index tile groups (the kernel's alignment unit) */
  while (gpu_num > 1 && (end >> 16) < (unsigned long long)gpu_num) gpu_num--;
  const unsigned long long chunks = (mode == SP_SCHED_DYNAMIC) ? dense_dynamic_chunks(nov, gpu_num) : 0;
  double sum = 0.0;
  /* index 0 (the base term p = prod x, gpu_exact_dense.cu:653) is part of device 0's range */
  int rc = sp_sched_run(&g_dense_ops, &job, mode, gpu_num, g_first_device, 0ull, end, 16, chunks, &sum, stats);
  free(mat_t);
  if (stats) stats->wall_ms = sp_now_ms() - t0;
  if (rc != SP_OK) return fail(stats, rc);
  return sp_nw_factor(nov) * sum;
}

/* Opens (and closes) the plans a following sp_dense_ryser call on this matrix will open, on every device it
 * will use: contexts, lanes, worker threads and the kernel instantiation for this order are in place
 * afterwards, so that the timed call measures the algorithm.  What `perman` does between reading the matrix
 * and starting its clock (the reference's clock starts after its own lazy initialisation as well: its first
 * CUDA call is in the wrapper, but the context exists since main() touched the device). */
int sp_prepare_dense(const double *mat, int nov, int gpu_num) {
  if (check_dense_args(mat, nov, NULL) != SP_OK) return SP_EINVAL;
  if (nov < 2) return SP_OK;
  double x[64];
  double *mat_t = (double *)malloc((size_t)nov * nov * sizeof(double));
  if (!mat_t) { sp_set_error("out of memory"); return SP_ENOMEM; }
  nw_preamble(mat, nov, x, mat_t);
  dense_job job = {mat_t, x, nov};
  double sum = 0.0;
  sp_stats st;
  int rc = sp_sched_run(&g_dense_ops, &job, SP_SCHED_PREPARE, gpu_num < 1 ? 1 : gpu_num, g_first_device, 0ull, 0ull, 0, 0ull, &sum, &st);
  free(mat_t);
  return rc;
}

struct sp_dense_handle {
  spd_dense_plan *plan;
  int nov;
  int device;
};

int sp_dense_open(const double *mat, int nov, int device, sp_dense_handle **h) {
  if (!h) { sp_set_error("null handle pointer"); return SP_EINVAL; }
  int rc = check_dense_args(mat, nov, NULL);
  if (rc != SP_OK) return rc;
  if (nov < 2) { sp_set_error("resident dense handle needs n >= 2"); return SP_ELIMIT; }
  double x[64];
  double *mat_t = (double *)malloc((size_t)nov * nov * sizeof(double));
  sp_dense_handle *hh = (sp_dense_handle *)calloc(1, sizeof(*hh));
  if (!mat_t || !hh) { free(mat_t); free(hh); sp_set_error("out of memory"); return SP_ENOMEM; }
  nw_preamble(mat, nov, x, mat_t);
  rc = spd_dense_plan_create(device, mat_t, x, nov, &hh->plan);
  free(mat_t);
  if (rc != SPD_OK) { sp_set_error("%s", spd_last_error()); free(hh); return rc; }
  hh->nov = nov; hh->device = device;
  *h = hh;
  return SP_OK;
}

double sp_dense_run(sp_dense_handle *h, long long start, long long end, sp_stats *stats) {
  const double t0 = sp_now_ms();
  stats_clear(stats);
  if (!h || !h->plan) { sp_set_error("null handle"); return fail(stats, SP_EINVAL); }
  if (start < 0 || end < start) { sp_set_error("bad range [%lld, %lld)", start, end); return fail(stats, SP_EINVAL); }
  double sum = 0.0;
  spd_run_info info;
  int rc = spd_dense_plan_run(h->plan, (unsigned long long)start, (unsigned long long)end, &sum, &info);
  if (rc != SPD_OK) { sp_set_error("%s", spd_last_error()); return fail(stats, rc); }
  if (stats) {
    stats->kernel_ms = info.kernel_ms;
    stats->device_ms[0] = info.kernel_ms;
    stats->device_partial[0] = sum;
    stats->device_units[0] = info.units;
    stats->units = info.units; stats->visited = info.visited;
    stats->devices = 1; stats->chunks = 1; stats->launches = info.launches;
    stats->path = info.path; stats->tile_log2 = info.tile_log2;
    stats->wall_ms = sp_now_ms() - t0;
  }
  return sum;
}

void sp_dense_close(sp_dense_handle *h) {
  if (!h) return;
  spd_dense_plan_destroy(h->plan);
  free(h);
}

double sp_dense_ryser_range(const double *mat, int nov, int device, long long start, long long end,
                            sp_stats *stats) {
  const double t0 = sp_now_ms();
  sp_dense_handle *h = NULL;
  stats_clear(stats);
  int rc = sp_dense_open(mat, nov, device, &h);
  if (rc != SP_OK) return fail(stats, rc);
  double r = sp_dense_run(h, start, end, stats);
  sp_dense_close(h);
  if (stats) stats->wall_ms = sp_now_ms() - t0;
  return r;
}

/* ================================================================================================
 * Sparse exact paths
 * ============================================================================================== */

/* NW start vector of the sparse wrappers (gpu_exact_sparse.cu:861-871): row sums over the
 * non-zeros of the dense matrix; and D, the CCS scattered back to dense (transposed). */
static int sparse_preamble(const double *mat, const int *cptrs, const int *rows, const double *cvals,
                           int nov, double *x, double *dmat_t) {
  for (int j = 0; j < nov; ++j) {
    double rs = 0.0;
    for (int k = 0; k < nov; ++k)
      if (mat[j * nov + k] != 0) rs += mat[j * nov + k];
    x[j] = mat[j * nov + (nov - 1)] - rs / 2;
  }
  memset(dmat_t, 0, (size_t)nov * nov * sizeof(double));
  for (int k = 0; k < nov; ++k) {
    if (cptrs[k] > cptrs[k + 1] || cptrs[k] < 0) { sp_set_error("malformed CCS column pointers"); return SP_EINVAL; }
    for (int t = cptrs[k]; t < cptrs[k + 1]; ++t) {
      const int r = rows[t];
      if (r < 0 || r >= nov) { sp_set_error("CCS row index %d out of range", r); return SP_EINVAL; }
      dmat_t[(size_t)k * nov + r] += cvals[t];
    }
  }
  return SP_OK;
}

typedef struct sparse_job {
  const double *dmat_t;
  const double *x;
  int nov;
  int skip;
} sparse_job;

static int sparse_open(const void *job, int device, void **plan) {
  const sparse_job *j = (const sparse_job *)job;
  /* the id entry points always cover the whole index space: the plan may choose its own column order */
  return sp_sparse_plan_open(device, j->dmat_t, j->x, j->nov, j->skip, SP_PLAN_WHOLE_SPACE, (spd_sparse_plan **)plan);
}
static int sparse_launch(void *plan, unsigned long long lo, unsigned long long hi) {
  return spd_sparse_plan_launch((spd_sparse_plan *)plan, lo, hi);
}
static int sparse_wait(void *plan, double *sum, spd_run_info *info) {
  return spd_sparse_plan_wait((spd_sparse_plan *)plan, sum, info);
}
static void sparse_close(void *plan) { spd_sparse_plan_destroy((spd_sparse_plan *)plan); }

static const sp_job_ops g_sparse_ops = {sparse_open, sparse_launch, sparse_wait, sparse_close};

/* SkipPer skips where some row's X is exactly 0 (gpu_exact_sparse.cu:644-666), i.e. where the selected
 * entries of that row cancel its Nijenhuis-Wilf shift exactly: common for 0/1 and small-integer rows,
 * essentially impossible for generic reals (the reference's own CPU SkipPer takes
 * 0.31 s on -b, 6.9 s on int/, 26.7 s on double/ input).  A row can only cancel if all its entries are
 * dyadic rationals; when no row qualifies, the skip bookkeeping (tile filter, block votes) cannot
 * pay and the same engine runs without it -- the sum is identical, skipped terms being exact zeros.
 * SP_SKIP_ALWAYS=1 keeps the skip engine on whatever the values. */
static int some_row_can_cancel(const double *mat, int nov) {
  const char *force = getenv("SP_SKIP_ALWAYS");
  if (force && force[0] == '1') return 1;
  for (int j = 0; j < nov; ++j) {
    int dyadic = 1;
    for (int k = 0; k < nov && dyadic; ++k) {
      const double v = mat[j * nov + k] * 1024.0;
      if (v != floor(v) || fabs(v) > 9007199254740992.0) dyadic = 0;
    }
    if (dyadic) return 1;
  }
  return 0;
}

int sp_prepare_sparse(const double *mat, const int *cptrs, const int *rows, const double *cvals, int nov,
                      int skipper, int gpu_num) {
  if (!mat || !cptrs || !rows || !cvals || nov < 2 || nov > 64) return SP_OK;   /* the real call reports it */
  if (skipper && !some_row_can_cancel(mat, nov)) skipper = 0;
  double x[64];
  double *dmat_t = (double *)malloc((size_t)nov * nov * sizeof(double));
  if (!dmat_t) { sp_set_error("out of memory"); return SP_ENOMEM; }
  int rc = sparse_preamble(mat, cptrs, rows, cvals, nov, x, dmat_t);
  if (rc == SP_OK) {
    sparse_job job = {dmat_t, x, nov, skipper};
    double sum = 0.0;
    sp_stats st;
    rc = sp_sched_run(&g_sparse_ops, &job, SP_SCHED_PREPARE, gpu_num < 1 ? 1 : gpu_num, g_first_device, 0ull, 0ull, 0, 0ull, &sum, &st);
  }
  free(dmat_t);
  return rc;
}

static double sparse_common(const double *mat, const int *cptrs, const int *rows, const double *cvals,
                            int nov, int skip, int mode, int gpu_num, sp_stats *stats) {
  const double t0 = sp_now_ms();
  if (!mat || !cptrs || !rows || !cvals) { sp_set_error("null argument"); return fail(stats, SP_EINVAL); }
  if (nov < 1 || nov > 64) { sp_set_error("sparse Ryser supports 1 <= n <= 64 (got %d)", nov); return fail(stats, SP_ELIMIT); }
  if (gpu_num < 1) gpu_num = 1;
  if (skip && !some_row_can_cancel(mat, nov)) skip = 0;
  if (nov == 1) {
    if (spd_device_count() <= 0) { sp_set_error("no CUDA device: %s", spd_last_error()); return fail(stats, SP_ENODEV); }
    return mat[0];
  }
  double x[64];
  double *dmat_t = (double *)malloc((size_t)nov * nov * sizeof(double));
  if (!dmat_t) { sp_set_error("out of memory"); return fail(stats, SP_ENOMEM); }
  int rc = sparse_preamble(mat, cptrs, rows, cvals, nov, x, dmat_t);
  if (rc != SP_OK) { free(dmat_t); return fail(stats, rc); }
  sparse_job job = {dmat_t, x, nov, skip};
  const unsigned long long end = 1ull << (nov - 1);
  while (gpu_num > 1 && (end >> 14) < (unsigned long long)gpu_num) gpu_num--;
  /* The sparse kernels work in tiles of 2^11-2^12 indices and need about 2^20 of them per launch to fill the
   * persistent grid: the reference's 2^(nov-30) chunks (x4 for SkipPer in round 1: 256 at n = 36) made launches of
   * 2^27 indices with 2^8-index tiles that are mostly prologue -- SkipPer n = 36 took 78 ms in 256 chunks against
   * 25 ms in one piece.  Two chunks per device, both in flight (8 x B200, n = 36: SkipPer 6.9 / 8.5 / 9.7 ms and SpaRyser
   * 4.3 / 5.1 / 6.1 ms with 2 / 4 / 8 chunks per device: what finer chunks even out is less than what their
   * shorter tiles cost). */
  unsigned long long chunks = 0;
  if (mode == SP_SCHED_DYNAMIC) {
    chunks = sp_dynamic_chunks(nov, 30, gpu_num);
    const char *e = getenv("SP_SPARSE_CHUNKS_PER_DEVICE");                 /* development knob */
    const unsigned long long per_dev = (e && atoi(e) > 0) ? (unsigned long long)atoi(e) : 2ull;
    const unsigned long long cap = per_dev * (unsigned long long)(gpu_num > 0 ? gpu_num : 1);
    if (chunks > cap) chunks = cap;
  }
  double sum = 0.0;
  rc = sp_sched_run(&g_sparse_ops, &job, mode, gpu_num, g_first_device, 0ull, end, 14, chunks, &sum, stats);
  free(dmat_t);
  if (stats) stats->wall_ms = sp_now_ms() - t0;
  if (rc != SP_OK) return fail(stats, rc);
  return sp_nw_factor(nov) * sum;
}

double sp_sparse_ryser(const double *mat, const int *cptrs, const int *rows, const double *cvals,
                       int nov, int algo_id, int gpu_num, int use_cpu, int threads, sp_stats *stats) {
  (void)use_cpu; (void)threads;
  stats_clear(stats);
  int mode = SP_SCHED_STATIC;
  switch (algo_id) {
    case 1: case 2: case 3: case 4: gpu_num = 1; break;   /* main.cu:108-127 */
    case 5: break;                                         /* main.cu:128-131 */
    case 6: mode = SP_SCHED_DYNAMIC; break;                /* main.cu:132-136 */
    default:
      sp_set_error("Unknown Algorithm ID");
      return fail(stats, SP_EALGO);
  }
  return sparse_common(mat, cptrs, rows, cvals, nov, 0, mode, gpu_num, stats);
}

double sp_skipper(const double *mat, const int *rptrs, const int *cols, const int *cptrs,
                  const int *rows, const double *cvals, int nov, int algo_id, int gpu_num, int use_cpu,
                  int threads, sp_stats *stats) {
  (void)rptrs; (void)cols; (void)use_cpu; (void)threads;
  stats_clear(stats);
  int mode = SP_SCHED_STATIC;
  switch (algo_id) {
    case 7: gpu_num = 1; break;                            /* main.cu:137-141 */
    case 8: mode = SP_SCHED_DYNAMIC; break;                /* main.cu:142-146 */
    default:
      sp_set_error("Unknown Algorithm ID");
      return fail(stats, SP_EALGO);
  }
  return sparse_common(mat, cptrs, rows, cvals, nov, 1, mode, gpu_num, stats);
}

double sp_sparse_ryser_range(const double *mat, const int *cptrs, const int *rows, const double *cvals,
                             int nov, int skipper, int device, long long start, long long end,
                             sp_stats *stats) {
  const double t0 = sp_now_ms();
  stats_clear(stats);
  if (!mat || !cptrs || !rows || !cvals) { sp_set_error("null argument"); return fail(stats, SP_EINVAL); }
  if (nov < 2 || nov > 64) { sp_set_error("sparse range needs 2 <= n <= 64 (got %d)", nov); return fail(stats, SP_ELIMIT); }
  if (start < 0 || end < start) { sp_set_error("bad range [%lld, %lld)", start, end); return fail(stats, SP_EINVAL); }
  double x[64];
  double *dmat_t = (double *)malloc((size_t)nov * nov * sizeof(double));
  if (!dmat_t) { sp_set_error("out of memory"); return fail(stats, SP_ENOMEM); }
  int rc = sparse_preamble(mat, cptrs, rows, cvals, nov, x, dmat_t);
  if (rc != SP_OK) { free(dmat_t); return fail(stats, rc); }
  spd_sparse_plan *plan = NULL;
  rc = sp_sparse_plan_open(device, dmat_t, x, nov, skipper, 0, &plan);
  free(dmat_t);
  if (rc != SPD_OK) return fail(stats, rc);
  double sum = 0.0;
  spd_run_info info;
  rc = spd_sparse_plan_run(plan, (unsigned long long)start, (unsigned long long)end, &sum, &info);
  spd_sparse_plan_destroy(plan);
  if (rc != SPD_OK) { sp_set_error("%s", spd_last_error()); return fail(stats, rc); }
  if (stats) {
    stats->kernel_ms = info.kernel_ms;
    stats->device_ms[0] = info.kernel_ms;
    stats->device_partial[0] = sum;
    stats->device_units[0] = info.units;
    stats->units = info.units; stats->visited = info.visited;
    stats->devices = 1; stats->chunks = 1; stats->launches = info.launches;
    stats->path = info.path; stats->tile_log2 = info.tile_log2;
    stats->wall_ms = sp_now_ms() - t0;
  }
  return sum;
}

/* ================================================================================================
 * Approximations
 * ============================================================================================== */
#define SP_DEFAULT_SEED 0x5355506572ULL   /* "SUPer" */

static unsigned long long pick_seed(unsigned long long seed) {
  if (seed != 0) return seed;
  const char *e = getenv("SP_SEED");
  if (e && *e) {
    unsigned long long v = strtoull(e, NULL, 0);
    if (v != 0) return v;
  }
  return SP_DEFAULT_SEED;
}

typedef struct approx_job {
  const int *rptrs, *cols, *cptrs, *rows;
  const double *rvals, *cvals;
  int nov, nnz, scaling, y, z;
  unsigned long long seed;
} approx_job;

static int approx_open(const void *job, int device, void **plan) {
  const approx_job *j = (const approx_job *)job;
  return spd_approx_plan_create(device, j->rptrs, j->cols, j->cptrs, j->rows, j->rvals, j->cvals, j->nov,
                                j->nnz, j->scaling, j->y, j->z, j->seed, (spd_approx_plan **)plan);
}
static int approx_launch(void *plan, unsigned long long lo, unsigned long long hi) {
  return spd_approx_plan_launch((spd_approx_plan *)plan, lo, hi);
}
static int approx_wait(void *plan, double *sum, spd_run_info *info) {
  return spd_approx_plan_wait((spd_approx_plan *)plan, sum, info);
}
static void approx_close(void *plan) { spd_approx_plan_destroy((spd_approx_plan *)plan); }
static const sp_job_ops g_approx_ops = {approx_open, approx_launch, approx_wait, approx_close};

static double approx_common(const approx_job *job, long long trials, int gpu_num, sp_stats *stats) {
  const double t0 = sp_now_ms();
  stats_clear(stats);
  if (!job->rptrs || !job->cols || !job->cptrs || !job->rows) { sp_set_error("null argument"); return fail(stats, SP_EINVAL); }
  if (trials < 1) { sp_set_error("number of trials must be positive (got %lld)", trials); return fail(stats, SP_EINVAL); }
  if (gpu_num < 1) gpu_num = 1;
  if ((long long)gpu_num > trials) gpu_num = (int)trials;
  sp_stats local;
  sp_stats *st = stats ? stats : &local;
  double sum = 0.0;
  /* static even split of the trial indices over the devices; every device returns its sum, its sum of
   * (scaled) squares and its count of trials that reached the last step, all carried by the scheduler */
  int rc = sp_sched_run(&g_approx_ops, job, SP_SCHED_STATIC, gpu_num, g_first_device, 0ull, (unsigned long long)trials, 0, 0, &sum, st);
  if (rc == SP_OK) {
    const double n = (double)trials, mean = sum / n;
    const double sc = st->sq_scale != 0.0 ? st->sq_scale : 1.0, ms = mean * sc;
    double var = st->sumsq_scaled / n - ms * ms;
    if (var < 0) var = 0;
    st->std_error = (n > 1) ? sqrt(var / (n - 1)) / sc : 0.0;
    /* the squares are accumulated in a scaled domain chosen from the order of the pattern; if they left the
     * double range anyway, say so instead of reporting 0 */
    if (!isfinite(st->sumsq_scaled) || (sum != 0.0 && st->sumsq_scaled == 0.0)) st->std_error = NAN;
  }
  st->wall_ms = sp_now_ms() - t0;
  if (rc != SP_OK) return fail(stats, rc);
  return sum / (double)trials;
}

int sp_prepare_approx(const int *rptrs, const int *cols, const int *cptrs, const int *rows, int nov, int nnz,
                      int scaling, int scale_intervals, int scale_times, int gpu_num) {
  if (!rptrs || !cols || !cptrs || !rows) return SP_OK;
  approx_job job = {rptrs, cols, cptrs, rows, NULL, NULL, nov, nnz, scaling, scaling ? scale_intervals : 1,
                    scaling ? scale_times : 0, pick_seed(0)};
  double sum = 0.0;
  sp_stats st;
  return sp_sched_run(&g_approx_ops, &job, SP_SCHED_PREPARE, gpu_num < 1 ? 1 : gpu_num, g_first_device, 0ull, 0ull, 0, 0ull, &sum, &st);
}

double sp_rasmussen_sparse(const int *rptrs, const int *cols, const int *cptrs, const int *rows,
                           int nov, int nnz, long long trials, int gpu_num,
                           unsigned long long seed, sp_stats *stats) {
  approx_job job = {rptrs, cols, cptrs, rows, NULL, NULL, nov, nnz, 0, 1, 0, pick_seed(seed)};
  return approx_common(&job, trials, gpu_num, stats);
}

double sp_scaling_sparse(const int *cptrs, const int *rows, const int *rptrs, const int *cols,
                         int nov, int nnz, long long trials, int scale_intervals, int scale_times,
                         int gpu_num, unsigned long long seed, sp_stats *stats) {
  approx_job job = {rptrs, cols, cptrs, rows, NULL, NULL, nov, nnz, 1, scale_intervals, scale_times, pick_seed(seed)};
  return approx_common(&job, trials, gpu_num, stats);
}

/* pattern (entries != 0) of a dense matrix as CRS + CCS with values */
static int dense_pattern(const double *mat, int nov, int **rptrs, int **cols, double **rvals, int **cptrs,
                         int **rows, double **cvals, int *nnz_out) {
  int nnz = 0;
  for (size_t e = 0; e < (size_t)nov * nov; ++e) nnz += (mat[e] != 0);
  const size_t c1 = (size_t)(nnz > 0 ? nnz : 1);
  *rptrs = (int *)malloc((size_t)(nov + 1) * sizeof(int));
  *cptrs = (int *)malloc((size_t)(nov + 1) * sizeof(int));
  *cols = (int *)malloc(c1 * sizeof(int));
  *rows = (int *)malloc(c1 * sizeof(int));
  *rvals = (double *)malloc(c1 * sizeof(double));
  *cvals = (double *)malloc(c1 * sizeof(double));
  if (!*rptrs || !*cptrs || !*cols || !*rows || !*rvals || !*cvals) { sp_set_error("out of memory"); return SP_ENOMEM; }
  int r = 0, c = 0;
  for (int a = 0; a < nov; ++a) {
    (*rptrs)[a] = r; (*cptrs)[a] = c;
    for (int b = 0; b < nov; ++b) {
      if (mat[(size_t)a * nov + b] != 0) { (*cols)[r] = b; (*rvals)[r] = mat[(size_t)a * nov + b]; ++r; }
      if (mat[(size_t)b * nov + a] != 0) { (*rows)[c] = b; (*cvals)[c] = mat[(size_t)b * nov + a]; ++c; }
    }
  }
  (*rptrs)[nov] = r; (*cptrs)[nov] = c;
  *nnz_out = nnz;
  return SP_OK;
}

static double dense_approx(const double *mat, int nov, long long trials, int scaling, int y, int z,
                           int gpu_num, unsigned long long seed, sp_stats *stats) {
  stats_clear(stats);
  if (!mat) { sp_set_error("mat is NULL"); return fail(stats, SP_EINVAL); }
  if (nov < 1 || nov > SP_MAX_NOV) { sp_set_error("matrix order %d out of range", nov); return fail(stats, SP_ELIMIT); }
  int *rptrs = NULL, *cols = NULL, *cptrs = NULL, *rows = NULL, nnz = 0;
  double *rvals = NULL, *cvals = NULL;
  int rc = dense_pattern(mat, nov, &rptrs, &cols, &rvals, &cptrs, &rows, &cvals, &nnz);
  double r = NAN;
  if (rc == SP_OK) {
    approx_job job = {rptrs, cols, cptrs, rows, scaling ? rvals : NULL, scaling ? cvals : NULL,
                      nov, nnz, scaling, y, z, pick_seed(seed)};
    r = approx_common(&job, trials, gpu_num, stats);
  } else if (stats) {
    stats->error = rc;
  }
  free(rptrs); free(cols); free(cptrs); free(rows); free(rvals); free(cvals);
  return r;
}

double sp_rasmussen_dense(const double *mat, int nov, long long trials, int gpu_num,
                          unsigned long long seed, sp_stats *stats) {
  return dense_approx(mat, nov, trials, 0, 1, 0, gpu_num, seed, stats);
}

double sp_scaling_dense(const double *mat, int nov, long long trials, int scale_intervals,
                        int scale_times, int gpu_num, unsigned long long seed, sp_stats *stats) {
  return dense_approx(mat, nov, trials, 1, scale_intervals, scale_times, gpu_num, seed, stats);
}

/* shared tail of the four per-trial entry points: one plan, one traced launch */
static double trace_common(const int *rptrs, const int *cols, const int *cptrs, const int *rows, const double *rvals,
                           const double *cvals, int nov, int nnz, int scaling, int y, int z, unsigned long long seed,
                           long long trial, int count, double *values, int *steps, double *partial, sp_stats *stats) {
  spd_approx_plan *plan = NULL;
  int rc = spd_approx_plan_create(g_first_device, rptrs, cols, cptrs, rows, rvals, cvals, nov, nnz, scaling, y, z,
                                  pick_seed(seed), &plan);
  if (rc != SPD_OK) { sp_set_error("%s", spd_last_error()); return fail(stats, rc); }
  double *est = values ? values : (double *)malloc((size_t)count * sizeof(double));
  if (!est) { spd_approx_plan_destroy(plan); sp_set_error("out of memory"); return fail(stats, SP_ENOMEM); }
  rc = spd_approx_plan_trace(plan, (unsigned long long)trial, (unsigned long long)trial + (unsigned long long)count, est, steps, partial);
  spd_approx_plan_destroy(plan);
  double total = 0.0;
  if (rc == SPD_OK) for (int i = 0; i < count; ++i) total += est[i];
  if (!values) free(est);
  if (rc != SPD_OK) { sp_set_error("%s", spd_last_error()); return fail(stats, rc); }
  return total;
}

double sp_approx_trace_sparse(const int *rptrs, const int *cols, const int *cptrs, const int *rows, int nov, int nnz,
                              int scaling, int scale_intervals, int scale_times, unsigned long long seed,
                              long long trial, int count, double *values, int *steps, double *partial,
                              sp_stats *stats) {
  stats_clear(stats);
  if (count < 1 || count > 65536 || trial < 0) { sp_set_error("bad argument"); return fail(stats, SP_EINVAL); }
  return trace_common(rptrs, cols, cptrs, rows, NULL, NULL, nov, nnz, scaling, scale_intervals, scale_times, seed, trial,
                      count, values, steps, partial, stats);
}

double sp_approx_trial_sparse(const int *rptrs, const int *cols, const int *cptrs, const int *rows,
                              int nov, int nnz, int scaling, int scale_intervals, int scale_times,
                              unsigned long long seed, long long trial, int count, double *values,
                              sp_stats *stats) {
  if (!values) { stats_clear(stats); sp_set_error("bad argument"); return fail(stats, SP_EINVAL); }
  return sp_approx_trace_sparse(rptrs, cols, cptrs, rows, nov, nnz, scaling, scale_intervals, scale_times, seed, trial,
                                count, values, NULL, NULL, stats);
}

/* per-trial records of the dense twins (pattern = entries != 0; the scaled estimator weights its
 * Sinkhorn sums by the entries, gpu_approximation_dense.cu:286-313) */
double sp_approx_trace_dense(const double *mat, int nov, int scaling, int scale_intervals, int scale_times,
                             unsigned long long seed, long long trial, int count, double *values, int *steps,
                             double *partial, sp_stats *stats) {
  stats_clear(stats);
  if (!mat || count < 1 || count > 65536 || trial < 0) { sp_set_error("bad argument"); return fail(stats, SP_EINVAL); }
  if (nov < 1 || nov > SP_MAX_NOV) { sp_set_error("matrix order %d out of range", nov); return fail(stats, SP_ELIMIT); }
  int *rptrs = NULL, *cols = NULL, *cptrs = NULL, *rows = NULL, nnz = 0;
  double *rvals = NULL, *cvals = NULL;
  int rc = dense_pattern(mat, nov, &rptrs, &cols, &rvals, &cptrs, &rows, &cvals, &nnz);
  double total = NAN;
  if (rc == SP_OK)
    total = trace_common(rptrs, cols, cptrs, rows, scaling ? rvals : NULL, scaling ? cvals : NULL, nov, nnz, scaling,
                         scale_intervals, scale_times, seed, trial, count, values, steps, partial, stats);
  else if (stats)
    stats->error = rc;
  free(rptrs); free(cols); free(cptrs); free(rows); free(rvals); free(cvals);
  return total;
}

double sp_approx_trial_dense(const double *mat, int nov, int scaling, int scale_intervals, int scale_times,
                             unsigned long long seed, long long trial, int count, double *values,
                             sp_stats *stats) {
  if (!values) { stats_clear(stats); sp_set_error("bad argument"); return fail(stats, SP_EINVAL); }
  return sp_approx_trace_dense(mat, nov, scaling, scale_intervals, scale_times, seed, trial, count, values, NULL, NULL,
                               stats);
}
