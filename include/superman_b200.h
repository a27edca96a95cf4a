/* superman_b200.h -- C-ABI of libsuperman_b200.so, the B200-native drop-in for SUPerman's
 * GPU permanent paths.
 *
 * The reference has no FFI on its GPU path: main.cu textually #includes the four .cu files and
 * RunAlgo (main.cu:20-248) calls the templated host wrappers gpu_perman64_* directly.  Its only
 * extern "C" surface is the CPU-only shim interface_connector.c:61-231.  The entry points below are
 * therefore (1) one function per gpu_perman64_* wrapper family -- same argument meaning, element
 * type widened to double (int / float matrix files are exactly representable) -- and (2) the
 * reader / CRS-CCS / ordering helpers of util.h that main.cu calls before them, so that a
 * maintainer can rebind RunAlgo to this library line by line (INTEGRATION.md shows the stub).
 *
 * Ownership: the caller allocates and frees every host array it passes in; arrays returned through
 * an sp_matrix are owned by that struct and released by sp_matrix_free.  The library owns all
 * device memory and never keeps a host pointer after a call returns.
 * Errors: a failing call returns NaN (functions returning double) or a negative SP_E* code and
 * sets sp_last_error(); `stats->error` carries the same code.  The reference checks no CUDA call
 * at all; a success path prints exactly what the reference prints.
 * There is no CPU fallback: with no CUDA device every compute entry point fails with SP_ENODEV.
 */
#ifndef SUPERMAN_B200_H
#define SUPERMAN_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define SP_OK        0
#define SP_ENODEV   -1
#define SP_EINVAL   -2
#define SP_ECUDA    -3
#define SP_ENOMEM   -4
#define SP_ELIMIT   -5
#define SP_EIO      -6
#define SP_EALGO    -7   /* "Unknown Algorithm ID" (main.cu:75) */

#define SP_MAX_DEVICES 16

typedef struct sp_stats {
  double kernel_ms;            /* max over devices of the CUDA-event time of their launches */
  double wall_ms;              /* host wall clock of the whole call */
  double device_ms[SP_MAX_DEVICES];       /* per device: sum of CUDA-event times */
  double device_partial[SP_MAX_DEVICES];  /* per device: its partial result (rank-order summed) */
  unsigned long long device_units[SP_MAX_DEVICES]; /* Gray indices / trials handled per device */
  unsigned long long units;    /* Gray indices in the range (exact) or trials run (approx) */
  unsigned long long visited;  /* Skipper: indices evaluated */
  double std_error;            /* approximations: standard error of the mean estimate */
  int devices;                 /* devices used */
  int chunks;                  /* chunks scheduled (dynamic paths), else == devices */
  int launches;                /* kernel launches */
  int path;                    /* SPD_PATH_* of the dominant kernel */
  int tile_log2;
  int error;                   /* SP_OK or SP_E* */
  /* approximations: sum over all trials of (estimate * sq_scale)^2 and that scale, from which std_error
   * is derived; `visited` counts the trials that reached the last step (the rest estimate 0) */
  double sumsq_scaled;
  double sq_scale;             /* (sparse exact paths: the chosen engine's modelled FP64 instructions per index) */
} sp_stats;

const char *sp_last_error(void);
int         sp_device_count(void);
const char *sp_version(void);
/* create contexts / streams / buffers of devices 0..gpu_num-1 ahead of time (optional) */
int         sp_warmup(int gpu_num);
/* Optional: open (and close) the plans a following call on the same input will open, on each of the gpu_num
 * devices it will use -- contexts, lanes, parked worker threads and the kernel instantiations for this input
 * are loaded afterwards (CUDA loads kernels lazily), so that a timed call measures the algorithm.  `perman`
 * calls these between reading the matrix and starting its clock. */
int         sp_prepare_dense(const double *mat, int nov, int gpu_num);
int         sp_prepare_sparse(const double *mat, const int *cptrs, const int *rows, const double *cvals, int nov,
                              int skipper, int gpu_num);
int         sp_prepare_approx(const int *rptrs, const int *cols, const int *cptrs, const int *rows, int nov, int nnz,
                              int scaling, int scale_intervals, int scale_times, int gpu_num);
/* Calculation precision of the dense exact entry points (sp_dense_ryser, sp_dense_ryser_range, sp_dense_open and
 * the dense leaves of sp_permanent_compressed).  SP_PRECISION_DOUBLE: FP64 (default).  SP_PRECISION_QUAD: the
 * revised front-end's -q (flags.calculation_quad, revised_perman/flags.h:61-64, main.cpp:1298-1325) as
 * double-double arithmetic: X, the products and the sums are pairs of doubles (~106 bits), about 8 x the FP64
 * instruction count.  For permanents that are tiny against the Ryser terms they are the sum of (chesapeake: FP64 is
 * 2e-6 off, double-double exact to the printed digits).  The half / mixed precision flags -h -w -v have no
 * counterpart: FP64 is this library's lowest precision.  Process-wide. */
#define SP_PRECISION_DOUBLE 0
#define SP_PRECISION_QUAD   1
int         sp_set_precision(int precision);
/* first device the permanent entry points use (default 0); ids with gpu_num devices use
 * first .. first+gpu_num-1 (the revised front-end's -l flag) */
int         sp_set_first_device(int device);

/* ---------------------------------------------------------------------------------------------
 * Matrix input and preprocessing (host, C): what main.cu does before RunAlgo.
 * ------------------------------------------------------------------------------------------- */
#define SP_TYPE_INT    0
#define SP_TYPE_FLOAT  1
#define SP_TYPE_DOUBLE 2
#define SP_MAX_NOV     8192

typedef struct sp_matrix {
  int nov;          /* order */
  int nnz;          /* number of entries > 0 (what CRS/CCS hold), counted */
  int header_nnz;   /* the `nnz` field of the file header (main.cu:497), informational */
  int type;         /* SP_TYPE_* declared by the file header */
  double *mat;      /* row-major nov*nov, zero-initialised (fixes main.cu:501,531,564) */
  int *cptrs, *rows;   double *cvals;   /* CCS: column pointers [nov+1], row ids, values */
  int *rptrs, *cols;   double *rvals;   /* CRS: row pointers [nov+1], column ids, values */
} sp_matrix;

/* Header sniff + ReadMatrix<T> (main.cu:494-498, util.h:343-358): `nov nnz {int|float|double}`
 * then `i j val` lines, 0-based; unparsable lines are skipped; binary != 0 is the -b flag (every
 * listed entry becomes 1).  A file whose first line is a `%%MatrixMarket matrix coordinate ...`
 * banner is read as MatrixMarket instead (1-based, real / integer / pattern, general / symmetric:
 * the revised front-end's format, revised_perman/read_matrix.hpp:11-157). */
int  sp_matrix_read(const char *path, int binary, sp_matrix *out);
int  sp_matrix_from_dense(const double *mat, int nov, sp_matrix *out);
/* matrix2compressed (preprocessing 0, util.h:522-551), _sortOrder (1, util.h:553-619; rewrites
 * mat in the new column order) or _skipOrder (2, util.h:621-684; permutes rows and columns). */
int  sp_matrix_compress(sp_matrix *m, int preprocessing);
/* gridGraph2compressed (util.h:403-520): biadjacency matrix of the m x n grid, nov = m*n/2,
 * with CRS and CCS.  Fails when both dimensions are odd. */
int  sp_matrix_grid(int m, int n, sp_matrix *out);
/* Structural preprocessing of the revised front-end (exact; host/sp_reduce.c).  All of them work on
 * m->mat in place and drop CRS/CCS (call sp_matrix_compress afterwards).
 *
 * sp_matrix_min_degree: getMinNnz (revised_perman/util.h:1181).
 * sp_matrix_reduce_step: one d1compress (util.h:1200), else one d2compress (util.h:1260), picking
 *   the same row / column as upstream; returns 0 (nothing applied), 1 or 2.  The degree-1 entry is
 *   multiplied into *factor instead of into the first matrix row (util.h:1251-1253).
 * sp_matrix_reduce: the loop of compress_singleton_and_then_recurse (revised_perman/main.cpp:1058)
 *   -- steps until none applies; an empty row or column collapses the matrix to the 1x1 zero matrix
 *   with *factor = 0.  perm(original) = *factor * perm(reduced).  Returns the rows removed.
 * sp_matrix_split34: d34compress (util.h:1333) on the first row / column with min_deg (3 or 4)
 *   non-zeros: m becomes the first (nov-1) matrix, *second (caller frees) the other one, and
 *   perm(before) = perm(m) + perm(second).  Returns 1, or 0 when no row / column has that degree.
 * sp_matrix_scale: scalesk + scaleMatrix (util.h:1445-1593), rv / cv receive the nov row / column
 *   factors; perm(original) = perm(scaled) / prod(cv) / prod(rv).  Returns the number of sweeps.
 * sp_matrix_balance: Sinkhorn-Knopp to convergence (every column sum within 0.1 % of threshold after
 *   the row pass, <= 1000 sweeps); same outputs as sp_matrix_scale.  What sp_permanent_compressed
 *   uses: upstream's stopping rule (mean sums only) ends after one sweep, which leaves the matrices
 *   produced by the degree compression too unbalanced for a Ryser sum in FP64.
 * sp_matrix_dm: Dulmage-Mendelsohn fine decomposition (util.h:309): erases the entries that lie on no
 *   perfect matching (they cannot contribute to the permanent); *matching receives the size of a
 *   maximum matching (< nov: the permanent is 0, nothing erased).  Returns the entries erased. */
int  sp_matrix_min_degree(const sp_matrix *m);
int  sp_matrix_reduce_step(sp_matrix *m, double *factor);
int  sp_matrix_reduce(sp_matrix *m, double *factor);
int  sp_matrix_split34(sp_matrix *m, int min_deg, sp_matrix *second);
int  sp_matrix_scale(sp_matrix *m, double threshold, double *rv, double *cv);
int  sp_matrix_balance(sp_matrix *m, double threshold, double *rv, double *cv);
int  sp_matrix_dm(sp_matrix *m, int *matching);
/* compress_singleton_and_then_recurse + compress_and_calculate_recursive + scale_and_calculate
 * (revised_perman/main.cpp:993-1260) on the GPU engine: degree-1/2 compression, then -- while the
 * smallest degree is < 5 and nov > leaf_nov (0: upstream's 30; < 0: no compression at all, scaling
 * only) -- d1 / d2 steps and d34 splits; every leaf is reduced to its total support (sp_matrix_dm)
 * and Sinkhorn-balanced (sp_matrix_balance) to row / column sums scaling_threshold when that is > 0
 * (upstream's -u), to 1 when it is 0 and the compression changed the matrix (without it FP64 Ryser
 * on the merged matrices can be off by tens of percent), not at all when it is < 0 (upstream's
 * default); with scaling on, the input is balanced once before the first compression step as well
 * (upstream's -u before -o, main.cpp:1637) and all factors are divided out in long double.  Leaves
 * are computed with sp_dense_ryser (sparse == 0; algo_id 0-6) or, after sp_matrix_compress(preprocessing),
 * sp_sparse_ryser (algo_id 1-6) / sp_skipper (7, 8).  mat is the row-major nov x nov matrix as read
 * (not reordered).  stats: sums over the leaves, chunks = number of leaves. */
double sp_permanent_compressed(const double *mat, int nov, int sparse, int preprocessing, int algo_id,
                               int gpu_num, int threads, double scaling_threshold, int leaf_nov,
                               sp_stats *stats);
void sp_matrix_free(sp_matrix *m);

/* ---------------------------------------------------------------------------------------------
 * Dense exact (Ryser / Nijenhuis-Wilf), ids -p0..-p6 of main.cu:34-75.
 * mat is row-major nov x nov (mat[i*nov+j]), as produced by ReadMatrix (util.h:343-358).
 * ------------------------------------------------------------------------------------------- */

/* Replaces gpu_perman64_xglobal / _xlocal / _xshared / _xshared_coalescing /
 * _xshared_coalescing_mshared (gpu_exact_dense.cu:401,459,518,576,640; ids 0-4: one device),
 * _mshared_multigpu (gpu_exact_dense.cu:701; id 5: static split over gpu_num devices) and
 * _mshared_multigpucpu_chunks (gpu_exact_dense.cu:776; id 6: dynamic chunk queue over gpu_num
 * devices).  All ids run the same sm_100a kernel; the id selects the partition only.
 * use_cpu / threads are accepted for signature parity (id 6's "-c -tN" CPU helper thread,
 * gpu_exact_dense.cu:822-845) and ignored: this library has no CPU compute path.
 * Returns the permanent.  Single-device ids use device 0 (the reference hard-codes device 1,
 * gpu_exact_dense.cu:664). */
double sp_dense_ryser(const double *mat, int nov, int algo_id, int gpu_num, int use_cpu, int threads,
                      sp_stats *stats);

/* The kernel-level contract of kernel_xshared_coalescing_mshared(mat_t, x, p, nov, start, end)
 * (gpu_exact_dense.cu:329-399) and of cpu_perman64(mat_t, x, nov, start, end, threads)
 * (gpu_exact_dense.cu:6-69): the signed sum of the Ryser terms with Gray index in [start, end) on
 * one device, WITHOUT the base term and WITHOUT the final (4*(nov&1)-2) factor -- except that
 * start == 0 is allowed and then includes index 0, the base term prod_j x_j.  This is what a
 * multi-process launcher (one rank per GPU) calls with its own slice. */
double sp_dense_ryser_range(const double *mat, int nov, int device, long long start, long long end,
                            sp_stats *stats);

/* Resident variant: upload once, run ranges many times (no H2D in the run). */
typedef struct sp_dense_handle sp_dense_handle;
int    sp_dense_open(const double *mat, int nov, int device, sp_dense_handle **h);
double sp_dense_run(sp_dense_handle *h, long long start, long long end, sp_stats *stats);
void   sp_dense_close(sp_dense_handle *h);

/* ---------------------------------------------------------------------------------------------
 * Sparse exact: SpaRyser (ids -s -p1..-p6, main.cu:107-136) and SkipPer (ids -s -p7/-p8,
 * main.cu:137-146).  mat is the dense matrix AFTER the chosen preprocessing (sp_matrix_compress
 * rewrites it, as util.h:608-618 / 670-681 do), cptrs/rows/cvals its CCS, rptrs/cols its CRS.
 * ------------------------------------------------------------------------------------------- */

/* Replaces gpu_perman64_xlocal_sparse / _xshared_sparse / _xshared_coalescing_sparse /
 * _xshared_coalescing_mshared_sparse (gpu_exact_sparse.cu:672,732,792,853; ids 1-4: one device),
 * _mshared_multigpu_sparse (:916; id 5: static split) and _mshared_multigpucpu_chunks_sparse
 * (:995; id 6: dynamic chunks of 2^(nov-30)).  Returns the permanent. */
double sp_sparse_ryser(const double *mat, const int *cptrs, const int *rows, const double *cvals,
                       int nov, int algo_id, int gpu_num, int use_cpu, int threads, sp_stats *stats);

/* Replaces gpu_perman64_xshared_coalescing_mshared_skipper (gpu_exact_sparse.cu:1123; id 7: one
 * device) and _mshared_multigpucpu_chunks_skipper (:1192; id 8: dynamic chunks).  rptrs/cols are
 * accepted for signature parity: the skip test here needs only the column structure.
 * stats->visited reports how many Gray indices were actually evaluated.  A row can only reach X = 0
 * when all its entries are dyadic rationals; if no row qualifies (generic real values) the same engine
 * runs without the skip bookkeeping (identical sum, visited == units; SP_SKIP_ALWAYS=1 overrides). */
double sp_skipper(const double *mat, const int *rptrs, const int *cols, const int *cptrs,
                  const int *rows, const double *cvals, int nov, int algo_id, int gpu_num, int use_cpu,
                  int threads, sp_stats *stats);

/* Kernel-level contract (start, end) of the two sparse kernels (gpu_exact_sparse.cu:456, 556) and
 * of cpu_perman64_sparse / cpu_perman64_skipper (:7, :90): signed sum over Gray indices in
 * [start, end) on one device, no base term (unless start == 0), no final factor. */
double sp_sparse_ryser_range(const double *mat, const int *cptrs, const int *rows, const double *cvals,
                             int nov, int skipper, int device, long long start, long long end,
                             sp_stats *stats);

/* ---------------------------------------------------------------------------------------------
 * Approximations (-a): Rasmussen (-p1 / -p3) and the Sinkhorn-scaled estimator (-p2 / -p4),
 * main.cu:77-104 (dense), 157-184 (sparse), 250-323 (grid graphs).
 * trials is -x (number_of_times), scale_intervals -y, scale_times -z.  Exactly `trials` trials
 * are run and averaged (the reference rounds up to whole 1024-thread blocks, ids 1/2, or gives
 * the whole budget to the first GPU, ids 3/4 -- SURVEY.md Appendix C); with gpu_num > 1 the trial
 * indices are split evenly over the devices.  seed == 0 selects the library's fixed default
 * seed (or the environment variable SP_SEED): runs are reproducible, and a trial's value depends
 * only on (seed, trial index), not on the device count.
 * stats->std_error is the standard error of the returned mean, stats->units the trial count.
 * ------------------------------------------------------------------------------------------- */

/* gpu_perman64_rasmussen_sparse (gpu_approximation_sparse.cu:455; gpu_num = 1) and
 * gpu_perman64_rasmussen_multigpucpu_chunks_sparse (:497; gpu_num devices) */
double sp_rasmussen_sparse(const int *rptrs, const int *cols, const int *cptrs, const int *rows,
                           int nov, int nnz, long long trials, int gpu_num,
                           unsigned long long seed, sp_stats *stats);
/* gpu_perman64_approximation_sparse (gpu_approximation_sparse.cu:608) and
 * gpu_perman64_approximation_multigpucpu_chunks_sparse (:663) */
double sp_scaling_sparse(const int *cptrs, const int *rows, const int *rptrs, const int *cols,
                         int nov, int nnz, long long trials, int scale_intervals, int scale_times,
                         int gpu_num, unsigned long long seed, sp_stats *stats);
/* gpu_perman64_rasmussen (gpu_approximation_dense.cu:373) and
 * gpu_perman64_rasmussen_multigpucpu_chunks (:411): pattern = entries != 0, nov <= 64 as in the
 * reference's `long` bit masks is NOT required here (any nov whose pattern fits shared memory) */
double sp_rasmussen_dense(const double *mat, int nov, long long trials, int gpu_num,
                          unsigned long long seed, sp_stats *stats);
/* gpu_perman64_approximation (gpu_approximation_dense.cu:527) and
 * gpu_perman64_approximation_multigpucpu_chunks (:573): Sinkhorn sums weighted by the entries */
double sp_scaling_dense(const double *mat, int nov, long long trials, int scale_intervals,
                        int scale_times, int gpu_num, unsigned long long seed, sp_stats *stats);
/* One trial's estimate (for parity tests against the oracle): trial index `trial` of the stream
 * `seed`; scaling == 0 Rasmussen, else scaled with (y, z). */
double sp_approx_trial_sparse(const int *rptrs, const int *cols, const int *cptrs, const int *rows,
                              int nov, int nnz, int scaling, int scale_intervals, int scale_times,
                              unsigned long long seed, long long trial, int count, double *values,
                              sp_stats *stats);

double sp_approx_trial_dense(const double *mat, int nov, int scaling, int scale_intervals, int scale_times,
                             unsigned long long seed, long long trial, int count, double *values,
                             sp_stats *stats);
/* The same with each trial's record (count <= 65536, one launch): values[i] the estimate (0 for a dead
 * end), steps[i] the steps completed (== nov when the trial reached the last step) and partial[i] the
 * running product at that point.  On patterns where almost every trial dies (the 36 x 36 grid of BASELINE
 * config 5) the estimates alone compare 0 with 0; steps and partial products do not.  Any output may be NULL. */
double sp_approx_trace_sparse(const int *rptrs, const int *cols, const int *cptrs, const int *rows, int nov, int nnz,
                              int scaling, int scale_intervals, int scale_times, unsigned long long seed,
                              long long trial, int count, double *values, int *steps, double *partial,
                              sp_stats *stats);
double sp_approx_trace_dense(const double *mat, int nov, int scaling, int scale_intervals, int scale_times,
                             unsigned long long seed, long long trial, int count, double *values, int *steps,
                             double *partial, sp_stats *stats);

/* ---------------------------------------------------------------------------------------------
 * The reference's Python / MATLAB shim on the GPU engine (interface_connector.c:61-231,
 * matlab_calculate_return.h:4,12,20): same names and arguments; see superman_b200/host/sp_connector.c
 * for the algorithm numbering and the two defects fixed.  `connect()` is exported as sp_connect() here and
 * under its own name by libConnect.so (host/libconnect.c), the file the reference's bindings load.
 * ------------------------------------------------------------------------------------------- */
void   sp_connect(void);
double sp_read_calculate_return(char *filename, int algorithm, int nt, int x, int y, int z);
double sp_matlab_calculate_return_int(int *mat, int algorithm, int nt, int x, int y, int z, int nov, int nnz);
double sp_matlab_calculate_return_double(double *mat, int algorithm, int nt, int x, int y, int z, int nov, int nnz);
double read_calculate_return(char *filename, int algorithm, int nt, int x, int y, int z);
double matlab_calculate_return_int(int *mat, int algorithm, int nt, int x, int y, int z, int nov, int nnz);
double matlab_calculate_return_double(double *mat, int algorithm, int nt, int x, int y, int z, int nov, int nnz);

/* (4*(nov&1)-2): the factor every wrapper applies to base + sum (gpu_exact_dense.cu:698). */
double sp_nw_factor(int nov);

/* Measured FP64 instruction issue rate of a device (thread-level instr/s); roofline denominator. */
double sp_fp64_peak(int device, int millis);
/* Measured integer ALU issue rate (thread-level IADD3 / LOP3 instr/s): the estimators' roofline denominator. */
double sp_int_peak(int device, int millis);

#ifdef __cplusplus
}
#endif
#endif
