/* superman_b200_device.h -- the thin C-ABI layer between the C host code and the sm_100a CUDA
 * kernels (BASELINE.json north_star: "Host code in C ... calls CUDA through a thin C-ABI layer").
 *
 * Nothing here knows about files, flags or algorithm ids: a "plan" is one matrix made resident on
 * one device, and a "run" is one kernel pass over a range of Gray indices (exact paths) or trial
 * indices (approximations) that leaves ONE double on the host.  The reference has no such layer:
 * its host wrappers (gpu_perman64_*, e.g. gpu_exact_dense.cu:640-699) cudaMalloc / cudaMemcpy /
 * launch / copy back grid*block partial sums / cudaFree on every call.  The functions below are
 * what those wrappers' CUDA halves become.
 *
 * Threading: a plan belongs to one device; calls on different plans may come from different host
 * threads concurrently (one host thread per device is how the scheduler drives them).  Calls on the
 * same plan must be serialised by the caller.
 * Errors: every function returns 0 on success or a negative SPD_E* code; spd_last_error() returns
 * a thread-local message.  There is NO CPU fallback: without a usable CUDA device every entry
 * point fails with SPD_ENODEV.
 */
#ifndef SUPERMAN_B200_DEVICE_H
#define SUPERMAN_B200_DEVICE_H

#ifdef __cplusplus
extern "C" {
#endif

#define SPD_OK        0
#define SPD_ENODEV   -1   /* no CUDA device / driver */
#define SPD_EINVAL   -2   /* bad argument */
#define SPD_ECUDA    -3   /* CUDA runtime error (message in spd_last_error) */
#define SPD_ENOMEM   -4
#define SPD_ELIMIT   -5   /* size outside what the kernels support */

/* Per-run measurements, filled by every *_run / *_wait call. */
typedef struct spd_run_info {
  double kernel_ms;         /* CUDA-event time on the plan's stream around this run's launches */
  unsigned long long units; /* Gray indices covered (exact) or trials executed (approximations) */
  unsigned long long visited; /* Skipper: indices actually evaluated; approximations: trials that reached the
                               * last step (the others hit a dead end and estimate 0); otherwise == units */
  int launches;             /* kernels launched by this run */
  int path;                 /* SPD_PATH_* of the dominant kernel */
  int tile_log2;            /* log2 of the per-thread tile (exact register paths), else 0 */
  int reserved;
  double aux0, aux1;        /* approximations: sum of (estimate*aux1)^2, and the scale aux1;
                             * SpaRyser / SkipPer: aux1 = the host model's FP64 instructions per Gray index */
} spd_run_info;

#define SPD_PATH_DENSE_REG      1   /* X in registers, templated on n            */
#define SPD_PATH_DENSE_SMEM     2   /* X in shared memory, any n <= 64           */
#define SPD_PATH_SPARSE_REG     3
#define SPD_PATH_SPARSE_SMEM    4
#define SPD_PATH_SKIPPER        5
#define SPD_PATH_RASMUSSEN      6
#define SPD_PATH_SCALING        7
#define SPD_PATH_DENSE_DD       8   /* dense Ryser in double-double arithmetic (the -q precision mode) */

int         spd_device_count(void);               /* >= 0, or SPD_ENODEV */
const char *spd_last_error(void);
void        spd_shutdown(void);                   /* release pooled streams / buffers */
int         spd_device_name(int device, char *buf, int buflen);
int         spd_device_sm_count(int device);
int         spd_device_sm_clock_khz(int device);  /* max SM clock */

/* Measured FP64 issue rate of `device`: runs a register-only DFMA chain kernel for about
 * `millis` ms and returns warp-level FP64 instructions * 32 per second (thread-level FP64
 * instr/s), the denominator of the dense roofline.  Negative on error. */
double      spd_fp64_peak_instr_per_s(int device, int millis);
/* Measured integer ALU issue rate (IADD3 / LOP3 chains), thread-level instr/s: what the estimators'
 * instruction throughput is quoted against ("report against measured INT32 issue rate"). */
double      spd_int_peak_instr_per_s(int device, int millis);

/* ---- dense Ryser --------------------------------------------------------------------------- */
typedef struct spd_dense_plan spd_dense_plan;

/* Dense plans created after spd_set_quad(1) compute in double-double arithmetic (X, products and sums as
 * pairs of doubles, ~106 bits): the revised front-end's -q ("quad" calculation precision,
 * revised_perman/flags.h:61-64).  About 8 x the FP64 instruction count; process-wide switch. */
void spd_set_quad(int on);
int  spd_get_quad(void);

/* mat_t[k*nov + j] = A[j][k] (the transposed matrix the reference uploads, gpu_exact_dense.cu:657-674),
 * xbase[j] = A[j][nov-1] - rowsum_j/2 (gpu_exact_dense.cu:647-654).  2 <= nov <= 64. */
int  spd_dense_plan_create(int device, const double *mat_t, const double *xbase, int nov,
                           spd_dense_plan **plan);
void spd_dense_plan_destroy(spd_dense_plan *plan);
/* Signed sum of the Ryser terms with Gray index in [lo, hi), 0 <= lo <= hi <= 2^(nov-1); index 0
 * is the NW base term prod(xbase).  Synchronous: returns when *sum is valid. */
int  spd_dense_plan_run(spd_dense_plan *plan, unsigned long long lo, unsigned long long hi,
                        double *sum, spd_run_info *info);
/* Asynchronous pair (launch on the plan's stream / wait + fetch). */
int  spd_dense_plan_launch(spd_dense_plan *plan, unsigned long long lo, unsigned long long hi);
int  spd_dense_plan_wait(spd_dense_plan *plan, double *sum, spd_run_info *info);

/* ---- SpaRyser / SkipPer --------------------------------------------------------------------- */
typedef struct spd_sparse_plan spd_sparse_plan;

/* A sparse plan as the C host planned it (superman_b200/host/sp_level.c: sp_sparse_plan_open):
 *   mat_t[k*nov + j] = D[j][k], D the matrix the reference's sparse kernels iterate over -- the CCS (cptrs, rows,
 *     cvals) scattered back to dense form (gpu_exact_sparse.cu:467-476, 521-549) -- with the ROWS already ordered
 *     by level (the lowest flippable column holding a non-zero of the row) and the columns in the order the
 *     plan will walk them; xbase in the same row order (row sums over the non-zeros, gpu_exact_sparse.cu:861-871);
 *   level_sorted[j] the level of the row now at position j (ascending);
 *   img: the LevelRyser engine's configuration and packed images (level_reg.cuh), or img->B == 0 for the
 *     hot/cold register kernel (7 <= nov <= 48) / the shared-memory kernel.
 * skip != 0 selects SkipPer (zero products are skipped, gpu_exact_sparse.cu:634-666), else SpaRyser.
 * Nothing is decided on this side: it uploads, loads the kernel, launches and reduces. */
typedef struct spd_level_image {
  int B, S0, S, R;              /* low columns, slots of level 0 / of the other levels, register-cold rows */
  int NC, NCP, HSP;             /* cold rows, padded cold rows, padded register rows */
  const double *colT_hot;       /* [(nov-1) * HSP] */
  const double *lowR;           /* [(S0 + (B-1)*S) * (B rounded up to even)] */
  const double *dcold;          /* [(nov-1) * NCP] */
  const double *xb_hot;         /* [HSP] */
  const double *xb_cold;        /* [NCP] */
  const int *cold_start;        /* [nov - B + 2] */
  double instr_per_index;       /* FP64 instructions per Gray index of this packing (reported in spd_run_info.aux1) */
  int skip_long_tiles;          /* SkipPer: 1 when tiles twice the default length cost less (the host measured how
                                 * many more tiles the shorter ones let the tile filter drop), else 0 */
} spd_level_image;
int  spd_sparse_plan_create_packed(int device, const double *mat_t, const double *xbase, const int *level_sorted,
                                   int nov, int skip, const spd_level_image *img, spd_sparse_plan **plan);
void spd_sparse_plan_destroy(spd_sparse_plan *plan);
int  spd_sparse_plan_run(spd_sparse_plan *plan, unsigned long long lo, unsigned long long hi,
                         double *sum, spd_run_info *info);
int  spd_sparse_plan_launch(spd_sparse_plan *plan, unsigned long long lo, unsigned long long hi);
int  spd_sparse_plan_wait(spd_sparse_plan *plan, double *sum, spd_run_info *info);

/* ---- Rasmussen / scaling estimators ---------------------------------------------------------- */
typedef struct spd_approx_plan spd_approx_plan;

/* CRS (rptrs, cols) and CCS (cptrs, rows) of the 0/1 pattern, as gridGraph2compressed /
 * matrix2compressed produce them.  rvals / cvals (may be NULL) are entry weights used by the
 * dense scaling twin (gpu_approximation_dense.cu:286-313).  scaling == 0: Rasmussen
 * (gpu_approximation_sparse.cu:198-290); else the scaled estimator (:292-452) with
 * scale_intervals (-y) and scale_times (-z).  A run covers the trial indices [lo, hi); *sum is the
 * sum of the per-trial estimates; a trial's value depends only on (seed, trial index). */
int  spd_approx_plan_create(int device, const int *rptrs, const int *cols, const int *cptrs,
                            const int *rows, const double *rvals, const double *cvals, int nov, int nnz,
                            int scaling, int scale_intervals, int scale_times, unsigned long long seed,
                            spd_approx_plan **plan);
void spd_approx_plan_destroy(spd_approx_plan *plan);
int  spd_approx_plan_run(spd_approx_plan *plan, unsigned long long lo, unsigned long long hi,
                         double *sum, spd_run_info *info);
int  spd_approx_plan_launch(spd_approx_plan *plan, unsigned long long lo, unsigned long long hi);
int  spd_approx_plan_wait(spd_approx_plan *plan, double *sum, spd_run_info *info);
int  spd_approx_plan_trial(spd_approx_plan *plan, unsigned long long trial, double *value);
/* Trials [lo, hi) (at most 2^16) in one launch with their per-trial record: estimate[i] (0 for a dead end),
 * steps[i] completed (== nov when the trial reached the last step) and the running product at that point.
 * Any of the three output arrays may be NULL.  For parity tests against the oracle. */
int  spd_approx_plan_trace(spd_approx_plan *plan, unsigned long long lo, unsigned long long hi,
                           double *estimate, int *steps, double *partial);

#ifdef __cplusplus
}
#endif
#endif
