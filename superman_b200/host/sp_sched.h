/* Per-device chunk scheduler (C, pthreads): partitions a range of work units -- Gray indices for
 * the exact paths, trial indices for the approximations -- over gpu_num devices, one host thread
 * per device, and combines one double per chunk on the host in a fixed order.
 *
 * Replaces the OpenMP regions of the reference's multi-GPU wrappers:
 *   static  : gpu_exact_dense.cu:729-771, gpu_exact_sparse.cu:940-989   (-p5)
 *   dynamic : gpu_exact_dense.cu:786-901, gpu_exact_sparse.cu:1005-1118, 1202-1321,
 *             gpu_approximation_sparse.cu:497-606, 663-789             (-p6, -p8, -a -p3/-p4)
 */
#ifndef SP_SCHED_H
#define SP_SCHED_H

#include "superman_b200.h"
#include "superman_b200_device.h"

typedef struct sp_job_ops {
  /* open a plan for this job on `device`; returns SPD_OK and *plan */
  int (*open)(const void *job, int device, void **plan);
  /* asynchronous pair on the plan's own stream */
  int (*launch)(void *plan, unsigned long long lo, unsigned long long hi);
  int (*wait)(void *plan, double *sum, spd_run_info *info);
  void (*close)(void *plan);
} sp_job_ops;

#define SP_SCHED_STATIC  0
#define SP_SCHED_DYNAMIC 1
#define SP_SCHED_PREPARE 2   /* no work: every device opens and closes two plans (contexts, lanes, kernel loads) */

/* Runs [lo, hi) over `gpu_num` devices (devices 0..gpu_num-1, or first_device.. when gpu_num == 1).
 *  static : device g gets one contiguous slice, boundaries rounded down to 2^align_log2;
 *  dynamic: n_chunks equal chunks (boundaries rounded to 2^align_log2) pulled from a shared
 *           counter; every device keeps two chunks in flight (two plans, two streams) so its SMs
 *           never wait for the host between chunks.
 * *total = sum over chunks in chunk order (bit-reproducible whatever device took which chunk).
 * Fills stats (devices, chunks, per-device ms / partial / units, kernel_ms = max device ms). */
int sp_sched_run(const sp_job_ops *ops, const void *job, int mode, int gpu_num, int first_device,
                 unsigned long long lo, unsigned long long hi, int align_log2,
                 unsigned long long n_chunks, double *total, sp_stats *stats);

/* pure partition helpers (unit-tested on the CPU) */
unsigned long long sp_sched_boundary(unsigned long long lo, unsigned long long hi,
                                     unsigned long long parts, unsigned long long idx, int align_log2);

/* host-side planning of a sparse exact plan (sp_level.c): column choice, row order, engine and packed images,
 * then spd_sparse_plan_create_packed.  flags: SP_PLAN_WHOLE_SPACE -- the plan will be run over the whole index
 * space [0, 2^(nov-1)) (in whatever chunks, on whatever devices, all opened with the same flag) and may walk the
 * columns 0 .. nov-2 in an order of its own choosing; without it [lo, hi) means the caller's Gray indices. */
#define SP_PLAN_WHOLE_SPACE 1
int sp_sparse_plan_open(int device, const double *dmat_t, const double *xbase, int nov, int skip, int flags,
                        spd_sparse_plan **plan);
/* the planning half on its own (no device involved; unit-tested on the CPU): what sp_sparse_plan_open hands to
 * spd_sparse_plan_create_packed */
typedef struct sp_level_plan {
  int n, skip;
  double *mat_t;              /* [n*n] mat_t[k*n + j]: rows ordered by level, columns in plan order */
  double xbase[64];           /* in the same row order */
  int level_sorted[64];
  spd_level_image img;        /* img.B == 0: hot/cold or shared-memory kernel */
  void *owned[6];             /* the arrays img points to */
} sp_level_plan;
int  sp_level_plan_build(const double *dmat_t, const double *xbase, int nov, int skip, int flags, sp_level_plan *out);
void sp_level_plan_free(sp_level_plan *plan);

void sp_set_error(const char *fmt, ...);
int sp_first_device(void);            /* what sp_set_first_device stored (sp_api.c) */
double sp_now_ms(void);

#endif
