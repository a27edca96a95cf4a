/* Per-device chunk scheduler -- see sp_sched.h. */
#define _POSIX_C_SOURCE 200809L
#include "sp_sched.h"

#include <math.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdatomic.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static _Thread_local char g_sp_err[512];

void sp_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_sp_err, sizeof(g_sp_err), fmt, ap);
  va_end(ap);
}

const char *sp_last_error(void) { return g_sp_err; }

double sp_now_ms(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec * 1e3 + (double)ts.tv_nsec * 1e-6;
}

unsigned long long sp_sched_boundary(unsigned long long lo, unsigned long long hi,
                                     unsigned long long parts, unsigned long long idx, int align_log2) {
  if (idx == 0 || parts == 0) return lo;
  if (idx >= parts) return hi;
  const unsigned long long len = hi - lo;
  const unsigned long long q = len / parts, r = len % parts;
  unsigned long long b = lo + q * idx + (idx < r ? idx : r);
  if (align_log2 > 0 && align_log2 < 63) {
    const unsigned long long mask = (1ull << align_log2) - 1ull;
    b &= ~mask;
  }
  if (b < lo) b = lo;
  if (b > hi) b = hi;
  return b;
}

typedef struct sched_shared {
  const sp_job_ops *ops;
  const void *job;
  int mode;
  int gpu_num, first_device;
  unsigned long long lo, hi;
  int align_log2;
  unsigned long long n_chunks;
  atomic_ullong next_chunk;
  atomic_int failed;
  double *chunk_sum;          /* n_chunks doubles */
  pthread_mutex_t err_mu;
  int err_code;
  char err_msg[512];
} sched_shared;

typedef struct sched_worker {
  sched_shared *sh;
  int rank;                   /* 0 .. gpu_num-1 */
  double ms;
  double partial;
  unsigned long long units, visited;
  int launches, chunks, path, tile_log2;
  double aux0, aux1;          /* approximations: sum of scaled squares, the scale */
} sched_worker;

static void worker_fail(sched_shared *sh, int code) {
  pthread_mutex_lock(&sh->err_mu);
  if (sh->err_code == 0) {
    sh->err_code = code;
    snprintf(sh->err_msg, sizeof(sh->err_msg), "%s", spd_last_error());
  }
  pthread_mutex_unlock(&sh->err_mu);
  atomic_store(&sh->failed, 1);
}

static void worker_account(sched_worker *w, const spd_run_info *info, double sum) {
  /* dynamic mode keeps two chunks in flight on two streams: their event times overlap, so the
   * device's busy time is taken from the host clock around its whole loop instead */
  if (w->sh->mode == SP_SCHED_STATIC) w->ms += info->kernel_ms;
  w->partial += sum;
  w->units += info->units;
  w->visited += info->visited;
  w->launches += info->launches;
  w->chunks += 1;
  w->aux0 += info->aux0;
  w->aux1 = info->aux1;
  w->path = info->path;
  if (info->tile_log2 > w->tile_log2) w->tile_log2 = info->tile_log2;
}

static void *worker_main(void *arg) {
  sched_worker *w = (sched_worker *)arg;
  sched_shared *sh = w->sh;
  const sp_job_ops *ops = sh->ops;
  const int device = sh->first_device + w->rank;
  void *plan[2] = {NULL, NULL};
  int rc;

  if (sh->mode == SP_SCHED_PREPARE) {
    /* open and close two plans on this device: creates the context and two pooled lanes if they do not
     * exist yet, loads the kernels the job will launch (CUDA loads kernels lazily) and sets their
     * shared-memory opt-in -- everything a first timed call would otherwise pay for */
    for (int s = 0; s < 2; ++s)
      if ((rc = ops->open(sh->job, device, &plan[s])) != SPD_OK) { worker_fail(sh, rc); break; }
    for (int s = 0; s < 2; ++s)
      if (plan[s]) ops->close(plan[s]);
    return NULL;
  }

  if (sh->mode == SP_SCHED_STATIC) {
    const unsigned long long a = sp_sched_boundary(sh->lo, sh->hi, (unsigned long long)sh->gpu_num,
                                                   (unsigned long long)w->rank, sh->align_log2);
    const unsigned long long b = sp_sched_boundary(sh->lo, sh->hi, (unsigned long long)sh->gpu_num,
                                                   (unsigned long long)w->rank + 1, sh->align_log2);
    if ((rc = ops->open(sh->job, device, &plan[0])) != SPD_OK) { worker_fail(sh, rc); return NULL; }
    double sum = 0.0;
    spd_run_info info;
    if ((rc = ops->launch(plan[0], a, b)) != SPD_OK || (rc = ops->wait(plan[0], &sum, &info)) != SPD_OK) {
      worker_fail(sh, rc);
    } else {
      worker_account(w, &info, sum);
      sh->chunk_sum[w->rank] = sum;
    }
    ops->close(plan[0]);
    return NULL;
  }

  /* dynamic: two plans per device, each holding one in-flight chunk */
  unsigned long long inflight[2];
  int busy[2] = {0, 0};
  for (int s = 0; s < 2; ++s) {
    if ((rc = ops->open(sh->job, device, &plan[s])) != SPD_OK) {
      worker_fail(sh, rc);
      if (s == 1) ops->close(plan[0]);
      return NULL;
    }
  }
  int slot = 0;
  const double t0 = sp_now_ms();
  for (;;) {
    /* refill every idle slot */
    for (int s = 0; s < 2; ++s) {
      if (busy[s] || atomic_load(&sh->failed)) continue;
      const unsigned long long c = atomic_fetch_add(&sh->next_chunk, 1ull);
      if (c >= sh->n_chunks) continue;
      const unsigned long long a = sp_sched_boundary(sh->lo, sh->hi, sh->n_chunks, c, sh->align_log2);
      const unsigned long long b = sp_sched_boundary(sh->lo, sh->hi, sh->n_chunks, c + 1, sh->align_log2);
      if ((rc = ops->launch(plan[s], a, b)) != SPD_OK) { worker_fail(sh, rc); continue; }
      inflight[s] = c;
      busy[s] = 1;
    }
    if (!busy[0] && !busy[1]) break;
    if (!busy[slot]) slot ^= 1;     /* oldest in-flight chunk first */
    double sum = 0.0;
    spd_run_info info;
    rc = ops->wait(plan[slot], &sum, &info);
    busy[slot] = 0;
    if (rc != SPD_OK) {
      worker_fail(sh, rc);
    } else {
      worker_account(w, &info, sum);
      sh->chunk_sum[inflight[slot]] = sum;
    }
    slot ^= 1;
  }
  w->ms = sp_now_ms() - t0;
  ops->close(plan[0]);
  ops->close(plan[1]);
  return NULL;
}

/* ---- parked worker threads ------------------------------------------------------------------------
 * Ranks 1 .. gpu_num-1 of a multi-device call run on threads that are created once and then wait on
 * a condition variable between calls (rank 0 runs on the caller's thread): a call costs a signal and
 * a wake-up per device instead of a pthread_create / join pair.  One multi-device call uses the pool at
 * a time; a second caller arriving meanwhile falls back to creating its own threads. */
typedef struct pool_slot {
  pthread_t tid;
  int started;
  sched_worker *work;         /* set by the dispatcher, cleared by the worker when done */
} pool_slot;

static pthread_mutex_t g_pool_use = PTHREAD_MUTEX_INITIALIZER;   /* held for the duration of one call */
static pthread_mutex_t g_pool_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_pool_go = PTHREAD_COND_INITIALIZER, g_pool_done = PTHREAD_COND_INITIALIZER;
static pool_slot g_pool[SP_MAX_DEVICES];

static void *pool_main(void *arg) {
  pool_slot *slot = (pool_slot *)arg;
  for (;;) {
    pthread_mutex_lock(&g_pool_mu);
    while (slot->work == NULL) pthread_cond_wait(&g_pool_go, &g_pool_mu);
    sched_worker *w = slot->work;
    pthread_mutex_unlock(&g_pool_mu);
    worker_main(w);
    pthread_mutex_lock(&g_pool_mu);
    slot->work = NULL;
    pthread_cond_broadcast(&g_pool_done);
    pthread_mutex_unlock(&g_pool_mu);
  }
  return NULL;
}

/* hands workers[1 .. gpu_num-1] to the pool; returns 0, or -1 when the pool cannot be used */
static int pool_dispatch(sched_worker *workers, int gpu_num) {
  if (pthread_mutex_trylock(&g_pool_use) != 0) return -1;
  pthread_mutex_lock(&g_pool_mu);
  for (int g = 1; g < gpu_num; ++g) {
    if (!g_pool[g].started) {
      pthread_attr_t at;
      pthread_attr_init(&at);
      pthread_attr_setdetachstate(&at, PTHREAD_CREATE_DETACHED);
      const int err = pthread_create(&g_pool[g].tid, &at, pool_main, &g_pool[g]);
      pthread_attr_destroy(&at);
      if (err != 0) {
        pthread_mutex_unlock(&g_pool_mu);
        pthread_mutex_unlock(&g_pool_use);
        return -1;                         /* nothing dispatched yet: the caller creates its own threads */
      }
      g_pool[g].started = 1;
    }
  }
  for (int g = 1; g < gpu_num; ++g) g_pool[g].work = &workers[g];
  pthread_cond_broadcast(&g_pool_go);
  pthread_mutex_unlock(&g_pool_mu);
  return 0;
}

static void pool_wait(int gpu_num) {
  pthread_mutex_lock(&g_pool_mu);
  for (int g = 1; g < gpu_num; ++g)
    while (g_pool[g].work != NULL) pthread_cond_wait(&g_pool_done, &g_pool_mu);
  pthread_mutex_unlock(&g_pool_mu);
  pthread_mutex_unlock(&g_pool_use);
}

int sp_sched_run(const sp_job_ops *ops, const void *job, int mode, int gpu_num, int first_device,
                 unsigned long long lo, unsigned long long hi, int align_log2,
                 unsigned long long n_chunks, double *total, sp_stats *stats) {
  if (gpu_num < 1) gpu_num = 1;
  if (gpu_num > SP_MAX_DEVICES) {
    sp_set_error("at most %d devices are supported (asked for %d)", SP_MAX_DEVICES, gpu_num);
    return SP_ELIMIT;
  }
  const int have = spd_device_count();
  if (have <= 0) {
    sp_set_error("no CUDA device: %s (libsuperman_b200 has no CPU fallback)", spd_last_error());
    return SP_ENODEV;
  }
  if (first_device + gpu_num > have) {
    sp_set_error("asked for %d device(s) starting at %d but only %d visible", gpu_num, first_device, have);
    return SP_EINVAL;
  }
  if (mode == SP_SCHED_STATIC || mode == SP_SCHED_PREPARE) n_chunks = (unsigned long long)gpu_num;
  if (n_chunks < 1) n_chunks = 1;

  sched_shared sh;
  memset(&sh, 0, sizeof(sh));
  sh.ops = ops; sh.job = job; sh.mode = mode;
  sh.gpu_num = gpu_num; sh.first_device = first_device;
  sh.lo = lo; sh.hi = hi; sh.align_log2 = align_log2; sh.n_chunks = n_chunks;
  atomic_init(&sh.next_chunk, 0ull);
  atomic_init(&sh.failed, 0);
  sh.chunk_sum = (double *)calloc((size_t)n_chunks, sizeof(double));
  if (!sh.chunk_sum) { sp_set_error("out of memory"); return SP_ENOMEM; }
  pthread_mutex_init(&sh.err_mu, NULL);

  sched_worker workers[SP_MAX_DEVICES];
  pthread_t tids[SP_MAX_DEVICES];
  memset(workers, 0, sizeof(workers));
  for (int g = 0; g < gpu_num; ++g) { workers[g].sh = &sh; workers[g].rank = g; }
  /* rank 0 runs on the calling thread: a single-device call involves no other thread at all; the other
   * ranks go to the parked workers, or to threads of our own when the pool is busy */
  const int pooled = (gpu_num > 1) && (pool_dispatch(workers, gpu_num) == 0);
  int own = 0;
  if (gpu_num > 1 && !pooled) {
    for (int g = 1; g < gpu_num; ++g) {
      if (pthread_create(&tids[g], NULL, worker_main, &workers[g]) != 0) {
        atomic_store(&sh.failed, 1);
        sh.err_code = SP_ENOMEM;
        snprintf(sh.err_msg, sizeof(sh.err_msg), "pthread_create failed");
        break;
      }
      own = g;
    }
  }
  worker_main(&workers[0]);
  if (pooled) pool_wait(gpu_num);
  for (int g = 1; g <= own; ++g) pthread_join(tids[g], NULL);
  pthread_mutex_destroy(&sh.err_mu);

  int rc = SP_OK;
  if (sh.err_code != 0) {
    sp_set_error("%s", sh.err_msg);
    rc = sh.err_code;
    *total = NAN;
  } else {
    /* fixed order: chunk 0, 1, 2, ... (== device rank order for the static split,
     * gpu_exact_dense.cu:769-771) */
    double t = 0.0;
    for (unsigned long long c = 0; c < n_chunks; ++c) t += sh.chunk_sum[c];
    *total = t;
  }
  if (stats) {
    stats->devices = gpu_num;
    stats->chunks = (int)(n_chunks > 0x7fffffffULL ? 0x7fffffff : n_chunks);
    stats->kernel_ms = 0.0;
    stats->units = 0; stats->visited = 0; stats->launches = 0;
    stats->sumsq_scaled = 0.0; stats->sq_scale = 0.0;
    for (int g = 0; g < gpu_num; ++g) {
      stats->device_ms[g] = workers[g].ms;
      stats->device_partial[g] = workers[g].partial;
      stats->device_units[g] = workers[g].units;
      if (workers[g].ms > stats->kernel_ms) stats->kernel_ms = workers[g].ms;
      stats->units += workers[g].units;
      stats->visited += workers[g].visited;
      stats->launches += workers[g].launches;
      stats->sumsq_scaled += workers[g].aux0;
      if (workers[g].aux1 != 0.0) stats->sq_scale = workers[g].aux1;
      if (workers[g].path) stats->path = workers[g].path;
      if (workers[g].tile_log2 > stats->tile_log2) stats->tile_log2 = workers[g].tile_log2;
    }
    stats->error = rc;
  }
  free(sh.chunk_sum);
  return rc;
}
