/* The reference's Python / MATLAB shim (interface_connector.c:61-231, superPython.py,
 * supermaTlab.m) re-pointed at the GPU engine -- SURVEY.md 8(f) rank 1.
 *
 * Same entry points and argument meaning:
 *   read_calculate_return(filename, algorithm, nt, x, y, z)
 *   matlab_calculate_return_int(mat, algorithm, nt, x, y, z, nov, nnz)
 *   matlab_calculate_return_double(mat, algorithm, nt, x, y, z, nov, nnz)
 * `algorithm` is the shim's own numbering (decide_and_call, interface_connector.c:18-59):
 *   0 rasmussen_sparse  1 rasmussen  2 approximation_perman64_sparse  3 approximation_perman64
 *   4 parallel_perman64_sparse  5 parallel_perman64  6 parallel_skip_perman64_w
 *   7 parallel_skip_perman64_w_balanced  8 perman64
 * with the preprocessing the shim picks (SortOrder for 0/2/4, SkipOrder for 6/7,
 * interface_connector.c:82-92).  x = trials, y = scale intervals, z = scale times; nt (threads) is
 * ignored.  Two defects of the shim are fixed: the result is returned as a double (the shim
 * truncates it through `int perman`, :22), and matrix values are honoured (the shim forces every
 * entry to 1, `generic = 0`, :76) -- set SP_CONNECT_BINARY=1 for the old 0/1 behaviour.
 * The shim's `connect()` is exported from this library as sp_connect(): a library that defines `connect`
 * and is linked into a process would shadow the socket call.  The file the reference's bindings actually
 * load, libConnect.so (superPython.py:6-7, supermaTlab.m), is built next to this library from
 * host/libconnect.c: it exports `connect` and the three entry points under their reference names and
 * forwards here; it is only ever dlopen'ed (ctypes / loadlibrary: RTLD_LOCAL), never linked.
 */
#define _POSIX_C_SOURCE 200809L
#include "superman_b200.h"
#include "sp_sched.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

void sp_connect(void) { printf("SUPerman Connected..\n"); }

static int connector_binary(void) {
  const char *e = getenv("SP_CONNECT_BINARY");
  return e && *e && strcmp(e, "0") != 0;
}

static int preprocessing_of(int algorithm) {
  if (algorithm == 0 || algorithm == 2 || algorithm == 4) return 1;
  if (algorithm == 6 || algorithm == 7) return 2;
  return 0;
}

static double decide_and_call(sp_matrix *m, int algorithm, int x, int y, int z) {
  sp_stats st;
  const int n = m->nov;
  switch (algorithm) {
    case 0: return sp_rasmussen_sparse(m->rptrs, m->cols, m->cptrs, m->rows, n, m->nnz, x, 1, 0, &st);
    case 1: return sp_rasmussen_dense(m->mat, n, x, 1, 0, &st);
    case 2: return sp_scaling_sparse(m->cptrs, m->rows, m->rptrs, m->cols, n, m->nnz, x, y, z, 1, 0, &st);
    case 3: return sp_scaling_dense(m->mat, n, x, y, z, 1, 0, &st);
    case 4: return sp_sparse_ryser(m->mat, m->cptrs, m->rows, m->cvals, n, 4, 1, 0, 0, &st);
    case 5: case 8: return sp_dense_ryser(m->mat, n, 4, 1, 0, 0, &st);
    case 6: case 7:
      return sp_skipper(m->mat, m->rptrs, m->cols, m->cptrs, m->rows, m->cvals, n, 7, 1, 0, 0, &st);
    default:
      sp_set_error("Algo unavailable");            /* interface_connector.c:53-56 */
      return NAN;
  }
}

double sp_read_calculate_return(char *filename, int algorithm, int nt, int x, int y, int z) {
  (void)nt;
  sp_matrix m;
  if (sp_matrix_read(filename, connector_binary(), &m) != SP_OK) return NAN;
  double r = NAN;
  if (sp_matrix_compress(&m, preprocessing_of(algorithm)) == SP_OK) r = decide_and_call(&m, algorithm, x, y, z);
  sp_matrix_free(&m);
  return r;
}

static double from_dense(const double *mat, int nov, int algorithm, int x, int y, int z) {
  sp_matrix m;
  if (sp_matrix_from_dense(mat, nov, &m) != SP_OK) return NAN;
  if (connector_binary())
    for (size_t e = 0; e < (size_t)nov * nov; ++e) m.mat[e] = (m.mat[e] != 0) ? 1.0 : 0.0;
  double r = NAN;
  if (sp_matrix_compress(&m, preprocessing_of(algorithm)) == SP_OK) r = decide_and_call(&m, algorithm, x, y, z);
  sp_matrix_free(&m);
  return r;
}

double sp_matlab_calculate_return_double(double *mat, int algorithm, int nt, int x, int y, int z, int nov, int nnz) {
  (void)nt; (void)nnz;
  if (!mat || nov < 1) { sp_set_error("bad matrix"); return NAN; }
  return from_dense(mat, nov, algorithm, x, y, z);
}

double sp_matlab_calculate_return_int(int *mat, int algorithm, int nt, int x, int y, int z, int nov, int nnz) {
  (void)nt; (void)nnz;
  if (!mat || nov < 1 || nov > SP_MAX_NOV) { sp_set_error("bad matrix"); return NAN; }
  double *d = (double *)malloc((size_t)nov * nov * sizeof(double));
  if (!d) { sp_set_error("out of memory"); return NAN; }
  for (size_t e = 0; e < (size_t)nov * nov; ++e) d[e] = (double)mat[e];
  const double r = from_dense(d, nov, algorithm, x, y, z);
  free(d);
  return r;
}

/* the reference's own names, kept in this library as well (libConnect.so forwards to the sp_ names above) */
double read_calculate_return(char *filename, int algorithm, int nt, int x, int y, int z) {
  return sp_read_calculate_return(filename, algorithm, nt, x, y, z);
}
double matlab_calculate_return_double(double *mat, int algorithm, int nt, int x, int y, int z, int nov, int nnz) {
  return sp_matlab_calculate_return_double(mat, algorithm, nt, x, y, z, nov, nnz);
}
double matlab_calculate_return_int(int *mat, int algorithm, int nt, int x, int y, int z, int nov, int nnz) {
  return sp_matlab_calculate_return_int(mat, algorithm, nt, x, y, z, nov, nnz);
}
