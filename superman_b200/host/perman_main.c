/* perman -- command-line front end, drop-in for the reference's `perman` (main.cu:325-600).
 *
 * Same flags, defaults, algorithm-id tables and stdout lines:
 *   -f/--file  -b/--binary  -s/--sparse  -r/--preprocessing  -t/--threads  -g/--gpu  -d/--device
 *   -c/--cpu  -a/--approximation  -p/--perman  -x/--numOfTimes  -y/--scaleIntervals
 *   -z/--scaleTimes  -i/--grid  -m/--gridm  -n/--gridn
 * GPU is the default when neither -g nor -c is given (main.cu:482-484).  CPU-only runs (-c without
 * -g) are the reference's algo.h paths and are NOT provided here (BASELINE.json: "no CPU
 * fallback"): they exit with status 1 and a message.  With -g, -c only meant "a CPU thread also
 * pulls chunks" (main.cu:66); it is accepted and ignored.
 * Also accepted, from the revised front-end (revised_perman/main.cpp:1298-1325): -k <reps> repeats
 * the calculation, -l <device> picks the first GPU, -o (= --reduce) applies the degree compression
 * and the d34 recursion, -u <t> Sinkhorn-scales each matrix to row/column sums t before the kernel;
 * -q (quad calculation precision) selects the double-double dense kernel; -h -w -v -e (half / mixed
 * precision and launch-shape knobs) are accepted and ignored: FP64 is the lowest precision here.
 * Extra, off by default: `--dm` erases the entries that lie on no perfect matching (Dulmage-
 * Mendelsohn, dead code upstream) before the exact algorithms; the environment variable PERMAN_PRECISION=<digits> adds a second line
 * `Result17: <name> <value>` with that many significant digits (the reference prints 6).
 */
#define _POSIX_C_SOURCE 200809L
#include <getopt.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "superman_b200.h"

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + (double)ts.tv_nsec * 1e-9;
}

static void print_kernel_lines(const sp_stats *st) {
  /* the wrappers print "kernel in <t>" (single GPU, gpu_exact_dense.cu:683) or
   * "kernel<g> in <t>" (multi GPU, :755) with the launch+sync wall time in seconds */
  if (st->devices <= 1) {
    printf("kernel in %g\n", st->kernel_ms * 1e-3);
  } else {
    for (int g = 0; g < st->devices; ++g) printf("kernel%d in %g\n", g, st->device_ms[g] * 1e-3);
  }
}

static void extra_precision(const char *name, double v) {
  const char *e = getenv("PERMAN_PRECISION");
  if (!e || !*e) return;
  int digits = atoi(e);
  if (digits < 1) digits = 17;
  if (digits > 40) digits = 40;
  printf("Result17: %s %.*g\n", name, digits, v);
}

/* cout << "Result: name " << perman << " in " << secs  (6 significant digits, main.cu:58) */
static void result_cout(const char *name, double v, double secs) {
  printf("Result: %s %g in %g\n", name, v, secs);
  extra_precision(name, v);
}
/* printf("Result: name %2lf in %lf\n") followed by the cout line (approximations, main.cu:82-83) */
static void result_both(const char *name, const char *try_prefix, double v, double secs) {
  printf("Result: %s %2lf in %lf\n", name, v, secs);
  printf("%s: %s %g in %g\n", try_prefix, name, v, secs);
  extra_precision(name, v);
}

/* Everything is printed and the process owns nothing else: leave without the CUDA runtime's atexit
 * teardown.  Destroying the contexts costs ~0.15 s per device (1.3 s of a 3.0 s `-p5 -d 8` process on
 * an 8 x B200 box, profiles/r01_cli_configs_8xB200.log) and is not part of any algorithm.
 * PERMAN_SLOW_EXIT=1 keeps the ordinary return path. */
static int finish(int rc) {
  fflush(stdout);
  fflush(stderr);
  if (!getenv("PERMAN_SLOW_EXIT")) _exit(rc);
  return rc;
}

static int report_failure(void) {
  fprintf(stderr, "perman: %s\n", sp_last_error());
  return 1;
}

/* -o / --reduce (flags.compression) and -u <t> (flags.scaling_threshold) of the revised front-end:
 * the exact ids go through sp_permanent_compressed (degree compression, d34 splits, scaling) */
static int g_compress = 0;
static double g_threshold = 0.0;   /* 0: scale leaves only when the compression changed the matrix */
static int g_preprocessing = 0;
static int g_usable_devices = 1;   /* visible devices from the first one used (-l) on */

/* id 66 is the reference's hand-made 3:3:1:1 split over exactly four GPUs (gpu_exact_dense.cu:906-989,
 * gpu_exact_sparse.cu:1326-1407), a hack for one heterogeneous box; it is served by the even static split
 * over four devices, or over as many as the box has */
static int devices_for_66(void) {
  const int g = g_usable_devices < 4 ? g_usable_devices : 4;
  if (g < 4) fprintf(stderr, "perman: -p66 wants 4 devices, %d usable; using %d\n", g_usable_devices, g);
  return g < 1 ? 1 : g;
}

/* full-precision companion lines of an approximation (PERMAN_PRECISION=1): standard error of the mean and
 * how many trials reached the last step (the rest ran into an empty row and estimate 0) */
static void approx_precision(const char *name, const sp_stats *st) {
  if (getenv("PERMAN_PRECISION"))
    printf("StdError: %s %.6g trials %llu survived %llu\n", name, st->std_error, st->units, st->visited);
}

static void print_compressed(const sp_stats *st) {
  printf("Compressed: %d leaf matrix(es), %llu Gray indices\n", st->chunks, st->units);
}

static int run_matrix(const sp_matrix *m, int perman_algo, int gpu_num, int threads, int cpu, int dense,
                      int approximation, int number_of_times, int scale_intervals, int scale_times) {
  sp_stats st;
  double start, perman;
  const int nov = m->nov;
  if (dense) {
    if (!approximation) {
      static const char *names[7] = {
          "gpu_perman64_xlocal",  /* id 0 prints the xlocal label in the reference too (main.cu:38) */
          "gpu_perman64_xlocal", "gpu_perman64_xshared", "gpu_perman64_xshared_coalescing",
          "gpu_perman64_xshared_coalescing_mshared", "gpu_perman64_xshared_coalescing_mshared_multigpu",
          "gpu_perman64_xshared_coalescing_mshared_multigpucpu_chunks"};
      if (perman_algo == 66) {
        start = now_s();
        perman = sp_dense_ryser(m->mat, nov, 5, devices_for_66(), cpu, threads, &st);
        if (isnan(perman) && st.error) return report_failure();
        print_kernel_lines(&st);
        printf("Result: gpu_perman64_xshared_coalescing_mshared_multigpu_manual_distribution %2lf in %lf\n",
               perman, now_s() - start);
        return 0;
      }
      if (perman_algo < 0 || perman_algo > 6) { printf("Unknown Algorithm ID\n"); return 0; }
      start = now_s();
      if (g_compress || g_threshold > 0) {
        perman = sp_permanent_compressed(m->mat, nov, 0, 0, perman_algo, gpu_num, threads, g_threshold,
                                         g_compress ? 0 : -1, &st);
        if (isnan(perman) && st.error) return report_failure();
        print_compressed(&st);
      } else {
        perman = sp_dense_ryser(m->mat, nov, perman_algo, gpu_num, cpu, threads, &st);
        if (isnan(perman) && st.error) return report_failure();
      }
      print_kernel_lines(&st);
      result_cout(names[perman_algo], perman, now_s() - start);
    } else {
      const char *name;
      start = now_s();
      switch (perman_algo) {
        case 1: name = "gpu_perman64_rasmussen";
          perman = sp_rasmussen_dense(m->mat, nov, number_of_times, 1, 0, &st); break;
        case 2: name = "gpu_perman64_approximation";
          perman = sp_scaling_dense(m->mat, nov, number_of_times, scale_intervals, scale_times, 1, 0, &st); break;
        case 3: name = "gpu_perman64_rasmussen_multigpucpu_chunks";
          perman = sp_rasmussen_dense(m->mat, nov, number_of_times, gpu_num, 0, &st); break;
        case 4: name = "gpu_perman64_approximation_multigpucpu_chunks";
          perman = sp_scaling_dense(m->mat, nov, number_of_times, scale_intervals, scale_times, gpu_num, 0, &st); break;
        default: printf("Unknown Algorithm ID\n"); return 0;
      }
      if (isnan(perman) && st.error) return report_failure();
      print_kernel_lines(&st);
      result_both(name, "Result", perman, now_s() - start);
      approx_precision(name, &st);
    }
  } else {
    if (!approximation) {
      const char *name = NULL;
      start = now_s();
      switch (perman_algo) {
        case 1: name = "gpu_perman64_xlocal_sparse"; break;
        case 2: name = "gpu_perman64_xshared_sparse"; break;
        case 3: name = "gpu_perman64_xshared_coalescing_sparse"; break;
        case 4: name = "gpu_perman64_xshared_coalescing_mshared_sparse"; break;
        case 5: name = "gpu_perman64_xshared_coalescing_mshared_multigpu_sparse"; break;
        case 6: name = "gpu_perman64_xshared_coalescing_mshared_multigpucpu_chunks_sparse"; break;
        case 7: name = "gpu_perman64_xshared_coalescing_mshared_skipper"; break;
        case 8: name = "gpu_perman64_xshared_coalescing_mshared_multigpucpu_chunks_skipper"; break;
        case 66: name = "gpu_perman64_xshared_coalescing_mshared_multigpu_sparse_manual_distribution"; break;
        default: printf("Unknown Algorithm ID\n"); return 0;
      }
      if ((g_compress || g_threshold > 0) && perman_algo != 66) {
        perman = sp_permanent_compressed(m->mat, nov, 1, g_preprocessing, perman_algo, gpu_num, threads, g_threshold,
                                         g_compress ? 0 : -1, &st);
        if (!(isnan(perman) && st.error)) print_compressed(&st);
      } else if (perman_algo == 7 || perman_algo == 8)
        perman = sp_skipper(m->mat, m->rptrs, m->cols, m->cptrs, m->rows, m->cvals, nov, perman_algo, gpu_num,
                            cpu, threads, &st);
      else if (perman_algo == 66) {
        if (g_compress || g_threshold > 0) fprintf(stderr, "perman: -o / -u are not applied to -p66\n");
        perman = sp_sparse_ryser(m->mat, m->cptrs, m->rows, m->cvals, nov, 5, devices_for_66(), cpu, threads, &st);
      }
      else
        perman = sp_sparse_ryser(m->mat, m->cptrs, m->rows, m->cvals, nov, perman_algo, gpu_num, cpu, threads, &st);
      if (isnan(perman) && st.error) return report_failure();
      print_kernel_lines(&st);
      result_cout(name, perman, now_s() - start);
    } else {
      const char *name;
      start = now_s();
      switch (perman_algo) {
        case 1: name = "gpu_perman64_rasmussen_sparse";
          perman = sp_rasmussen_sparse(m->rptrs, m->cols, m->cptrs, m->rows, nov, m->nnz, number_of_times, 1, 0, &st); break;
        case 2: name = "gpu_perman64_approximation_sparse";
          perman = sp_scaling_sparse(m->cptrs, m->rows, m->rptrs, m->cols, nov, m->nnz, number_of_times,
                                     scale_intervals, scale_times, 1, 0, &st); break;
        case 3: name = "gpu_perman64_rasmussen_multigpucpu_chunks_sparse";
          perman = sp_rasmussen_sparse(m->rptrs, m->cols, m->cptrs, m->rows, nov, m->nnz, number_of_times, gpu_num, 0, &st); break;
        case 4: name = "gpu_perman64_approximation_multigpucpu_chunks_sparse";
          perman = sp_scaling_sparse(m->cptrs, m->rows, m->rptrs, m->cols, nov, m->nnz, number_of_times,
                                     scale_intervals, scale_times, gpu_num, 0, &st); break;
        default: printf("Unknown Algorithm ID\n"); return 0;
      }
      if (isnan(perman) && st.error) return report_failure();
      print_kernel_lines(&st);
      result_both(name, "Result", perman, now_s() - start);
      approx_precision(name, &st);
    }
  }
  return 0;
}

/* RunPermanForGridGraphs (main.cu:250-323) */
static int run_grid(int gm, int gn, int perman_algo, int gpu_num, int number_of_times, int scale_intervals,
                    int scale_times) {
  sp_matrix g;
  if (sp_matrix_grid(gm, gn, &g) != SP_OK) {
    /* the reference prints this (sic) and returns without computing, main.cu:404-407 / 253-260 */
    printf("one of the grid dimensions should be positive.");
    return 0;
  }
  sp_prepare_approx(g.rptrs, g.cols, g.cptrs, g.rows, g.nov, g.nnz, perman_algo == 2 || perman_algo == 4, scale_intervals,
                    scale_times, (perman_algo == 3 || perman_algo == 4) ? gpu_num : 1);
  sp_stats st;
  double perman;
  const char *name, *try_name;
  const double start = now_s();
  switch (perman_algo) {
    case 1: name = try_name = "gpu_perman64_rasmussen_sparse";
      perman = sp_rasmussen_sparse(g.rptrs, g.cols, g.cptrs, g.rows, g.nov, g.nnz, number_of_times, 1, 0, &st); break;
    case 2: name = try_name = "gpu_perman64_approximation_sparse";
      perman = sp_scaling_sparse(g.cptrs, g.rows, g.rptrs, g.cols, g.nov, g.nnz, number_of_times, scale_intervals,
                                 scale_times, 1, 0, &st); break;
    case 3: name = try_name = "gpu_perman64_rasmussen_multigpucpu_chunks";
      perman = sp_rasmussen_sparse(g.rptrs, g.cols, g.cptrs, g.rows, g.nov, g.nnz, number_of_times, gpu_num, 0, &st); break;
    case 4: name = try_name = "gpu_perman64_approximation_multigpucpu_chunks_sparse";
      perman = sp_scaling_sparse(g.cptrs, g.rows, g.rptrs, g.cols, g.nov, g.nnz, number_of_times, scale_intervals,
                                 scale_times, gpu_num, 0, &st); break;
    default:
      printf("Unknown Algorithm ID\n");
      sp_matrix_free(&g);
      return 0;
  }
  if (isnan(perman) && st.error) { sp_matrix_free(&g); return report_failure(); }
  const double secs = now_s() - start;
  print_kernel_lines(&st);
  printf("Result: %s %2lf in %lf\n", name, perman, secs);
  printf("Try: %s %g in %g\n", try_name, perman, secs);
  extra_precision(name, perman);
  approx_precision(name, &st);
  /* the reference dumps m rows x n columns of the nov x nov matrix here (main.cu:309-316) */
  printf("------------GRID--------------\n");
  for (int i = 0; i < gm; ++i) {
    for (int j = 0; j < gn; ++j) {
      const long idx = (long)i * gn + j;
      printf("%d ", idx < (long)g.nov * g.nov ? (int)g.mat[idx] : 0);
    }
    printf("\n");
  }
  printf("------------GRID--------------\n");
  sp_matrix_free(&g);
  return 0;
}

int main(int argc, char **argv) {
  int generic = 1, dense = 1, approximation = 0, gpu = 0, cpu = 0;
  int gpu_num = 2, threads = 16;                     /* main.cu:332-333 */
  const char *filename = "";
  int perman_algo = 1, preprocessing = 0;            /* main.cu:335-336 */
  int number_of_times = 100000, scale_intervals = 4, scale_times = 5;   /* main.cu:338-340 */
  int grid_graph = 0, gridm = 36, gridn = 36;        /* main.cu:342-344 */
  int reduce = 0;                                    /* --reduce: not a reference flag, off by default */

  /* the reference's option string (main.cu:347) plus the revised front-end's extra letters
   * (revised_perman/main.cpp:1298): -k reps, -l device id, -o degree compression, -u <threshold>
   * Sinkhorn scaling, -q quad (double-double) calculation precision, and the half / mixed precision and
   * launch-shape flags -h -w -v -e, which are accepted and ignored (FP64 is the lowest precision of this
   * engine, and it sizes its own launches) */
  static const char *short_options = "bsr:t:f:gd:cap:x:y:z:im:n:hwqk:e:ol:vu:";
  int reps = 1, first_device = 0, dm = 0, quad = 0;
  static const struct option long_options[] = {
      {"binary", 0, NULL, 'b'},        {"sparse", 0, NULL, 's'},       {"preprocessing", 1, NULL, 'r'},
      {"threads", 1, NULL, 't'},       {"file", 1, NULL, 'f'},         {"gpu", 0, NULL, 'g'},
      {"device", 1, NULL, 'd'},        {"cpu", 0, NULL, 'c'},          {"approximation", 0, NULL, 'a'},
      {"perman", 1, NULL, 'p'},        {"numOfTimes", 1, NULL, 'x'},   {"scaleIntervals", 1, NULL, 'y'},
      {"scaleTimes", 1, NULL, 'z'},    {"grid", 0, NULL, 'i'},         {"gridm", 1, NULL, 'm'},
      {"gridn", 1, NULL, 'n'},         {"reduce", 0, NULL, 1000},      {"dm", 0, NULL, 1001},
      {NULL, 0, NULL, 0}};

  int opt;
  while ((opt = getopt_long(argc, argv, short_options, long_options, NULL)) != -1) {
    /* every valued option refuses an argument that looks like another option (main.cu:381-384) */
    if (optarg && optarg[0] == '-' && opt < 256 && strchr("rtfdpxyzmnklu", opt)) {
      /* the reference's message names -t for -r as well (main.cu:382); we name the real option */
      fprintf(stderr, "Option -%c requires an argument.\n", opt);
      return 1;
    }
    switch (opt) {
      case 'b': generic = 0; break;
      case 's': dense = 0; break;
      case 'r': preprocessing = atoi(optarg); break;
      case 't': threads = atoi(optarg); break;
      case 'f': filename = optarg; break;
      case 'a': approximation = 1; break;
      case 'g': gpu = 1; break;
      case 'd': gpu_num = atoi(optarg); break;
      case 'c': cpu = 1; break;
      case 'p': perman_algo = atoi(optarg); break;
      case 'x': number_of_times = atoi(optarg); break;
      case 'y': scale_intervals = atoi(optarg); break;
      case 'z': scale_times = atoi(optarg); break;
      case 'i': grid_graph = 1; break;
      case 'm': gridm = atoi(optarg); break;
      case 'n': gridn = atoi(optarg); break;
      case 1000: reduce = 1; break;
      case 'o': reduce = 1; break;                      /* flags.compression */
      case 'k': reps = atoi(optarg); break;             /* flags.rep */
      case 'l': first_device = atoi(optarg); break;     /* flags.device_id */
      case 'u': g_threshold = (double)atoi(optarg); break;   /* flags.scaling_threshold (main.cpp:1466) */
      case 1001: dm = 1; break;
      case 'q': quad = 1; break;             /* flags.calculation_quad (revised main.cpp:1298-1325): double-double */
      case 'h': case 'w': case 'v': case 'e': break;
      case '?': return 1;
      default: abort();
    }
  }
  if (!grid_graph && filename[0] == '\0') {
    fprintf(stderr, "Option -f is a required argument.\n");
    return 1;
  }
  for (int index = optind; index < argc; index++) printf("Non-option argument %s\n", argv[index]);
  if (!cpu && !gpu) gpu = 1;
  if (!gpu) {
    fprintf(stderr,
            "perman: CPU-only algorithms (-c without -g: algo.h of the reference) are not part of this build; "
            "run without -c to use the GPU paths\n");
    return 1;
  }
  /* the reference's single-GPU ids pick device 1 and need two GPUs (gpu_exact_dense.cu:664); here
   * -d only bounds the multi-GPU ids, capped by what is visible */
  const double t_init0 = now_s();
  const int visible = sp_device_count();
  if (visible <= 0) return report_failure();
  const double t_init1 = now_s();
  if (first_device < 0 || first_device >= visible) {
    fprintf(stderr, "perman: -l %d but %d device(s) visible\n", first_device, visible);
    return 1;
  }
  sp_set_first_device(first_device);
  if (gpu_num > visible - first_device) {
    fprintf(stderr, "perman: -d %d but only %d device(s) usable; using %d\n", gpu_num, visible - first_device,
            visible - first_device);
    gpu_num = visible - first_device;
  }
  if (gpu_num < 1) gpu_num = 1;
  g_usable_devices = visible - first_device;
  /* CUDA context creation (~0.15 s per device) is not part of any algorithm: do it before the
   * timed wrapper calls.  Single-GPU ids only need device 0. */
  {
    int need = 1;
    if ((!approximation && (perman_algo == 5 || perman_algo == 6 || perman_algo == 8)) ||
        (approximation && (perman_algo == 3 || perman_algo == 4)))
      need = gpu_num;
    if (perman_algo == 66) need = (visible - first_device) < 4 ? (visible - first_device) : 4;
    if (sp_warmup(need) != SP_OK) return report_failure();
    if (getenv("PERMAN_TIMING"))
      fprintf(stderr, "perman: driver init %.3f s, %d context(s) + buffers %.3f s\n", t_init1 - t_init0, need,
              now_s() - t_init1);
  }

  if (reps < 1) reps = 1;
  if (grid_graph) {
    int rcg = 0;
    for (int r = 0; r < reps && rcg == 0; ++r)
      rcg = run_grid(gridm, gridn, perman_algo, gpu_num, number_of_times, scale_intervals, scale_times);
    return finish(rcg);
  }

  sp_matrix m;
  if (sp_matrix_read(filename, !generic, &m) != SP_OK) return report_failure();
  if (quad && !approximation) {
    /* -q: the exact ids compute in double-double arithmetic; that engine is the dense one, which returns the
     * same permanent as SpaRyser / SkipPer (they only skip terms that are exactly zero) */
    if (!dense) fprintf(stderr, "perman: -q computes through the dense double-double kernel (-s is set aside)\n");
    dense = 1;
    sp_set_precision(SP_PRECISION_QUAD);
  }
  if (dm && !approximation) {
    /* Dulmage-Mendelsohn: entries on no perfect matching cannot contribute (revised util.h:309) */
    int matching = 0;
    const int erased = sp_matrix_dm(&m, &matching);
    if (erased < 0) { sp_matrix_free(&m); return report_failure(); }
    printf("DM: maximum matching %d of %d, %d entries erased\n", matching, m.nov, erased);
  }
  g_compress = reduce && !approximation;
  g_preprocessing = preprocessing;
  if (approximation) g_threshold = 0.0;
  /* with -o / -u the exact ids hand the matrix AS READ to sp_permanent_compressed, which reduces it
   * and applies the -r ordering per leaf; keep an unordered copy for it */
  if (g_compress || g_threshold > 0) preprocessing = 0;
  if (sp_matrix_compress(&m, preprocessing) != SP_OK) { sp_matrix_free(&m); return report_failure(); }
  /* load what the timed call will launch (contexts and lanes exist since sp_warmup; kernels load lazily) */
  {
    const int multi = (!approximation && (perman_algo == 5 || perman_algo == 6 || perman_algo == 8)) ||
                      (approximation && (perman_algo == 3 || perman_algo == 4));
    const int devs = perman_algo == 66 ? (g_usable_devices < 4 ? g_usable_devices : 4) : (multi ? gpu_num : 1);
    if (!g_compress && !(g_threshold > 0)) {
      if (approximation) {
        if (!dense) sp_prepare_approx(m.rptrs, m.cols, m.cptrs, m.rows, m.nov, m.nnz, perman_algo == 2 || perman_algo == 4,
                                      scale_intervals, scale_times, devs);
      } else if (dense) {
        sp_prepare_dense(m.mat, m.nov, devs);
      } else {
        sp_prepare_sparse(m.mat, m.cptrs, m.rows, m.cvals, m.nov, perman_algo == 7 || perman_algo == 8, devs);
      }
    }
  }
  int rc = 0;
  for (int r = 0; r < reps && rc == 0; ++r)
    rc = run_matrix(&m, perman_algo, gpu_num, threads, cpu, dense, approximation, number_of_times,
                    scale_intervals, scale_times);
  sp_matrix_free(&m);
  return finish(rc);
}
