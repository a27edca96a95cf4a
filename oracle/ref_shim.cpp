// ref_shim.cpp -- exposes the UNMODIFIED reference (util.h / algo.h, compiled where they lie under
// /root/reference) through a C ABI so tests can pin the oracle against it and bench.py can time it
// as the CPU baseline.  TEST INFRASTRUCTURE ONLY; output goes to oracle/_ref/ (git-ignored).
// No reference source is copied into this repository: this file only #includes it at build time.
#include <cstdint>
#include "util.h"
#include "algo.h"

extern "C" {

double ref_perman64(const double* mat, int nov) {                      // algo.h:1031
  return perman64((double*)mat, nov);
}
double ref_parallel_perman64(const double* mat, int nov, int threads) { // algo.h:662 (float X!)
  return parallel_perman64((double*)mat, nov, threads);
}
double ref_parallel_perman64_int(const int* mat, int nov, int threads) {
  return parallel_perman64((int*)mat, nov, threads);
}
double ref_parallel_perman64_sparse(const double* mat, const int* cptrs, const int* rows,
                                    const double* cvals, int nov, int threads) {  // algo.h:568
  return parallel_perman64_sparse((double*)mat, (int*)cptrs, (int*)rows, (double*)cvals, nov, threads);
}
double ref_parallel_skip_perman64_w(const int* rptrs, const int* cols, const double* rvals,
                                    const int* cptrs, const int* rows, const double* cvals, int nov,
                                    int threads) {                               // algo.h:748
  return parallel_skip_perman64_w((int*)rptrs, (int*)cols, (double*)rvals, (int*)cptrs, (int*)rows,
                                  (double*)cvals, nov, threads);
}
double ref_parallel_skip_perman64_w_balanced(const int* rptrs, const int* cols, const double* rvals,
                                             const int* cptrs, const int* rows, const double* cvals,
                                             int nov, int threads) {              // algo.h:885
  return parallel_skip_perman64_w_balanced((int*)rptrs, (int*)cols, (double*)rvals, (int*)cptrs,
                                           (int*)rows, (double*)cvals, nov, threads);
}

// util.h:522 / 553 / 621 -- outputs copied into caller arrays of size nov+1 / nnz
static void copy_out(int nov, int nnz, int* cptrs, int* rows, double* cvals, int* rptrs, int* cols,
                     double* rvals, int* o_cptrs, int* o_rows, double* o_cvals, int* o_rptrs,
                     int* o_cols, double* o_rvals) {
  for (int i = 0; i <= nov; i++) { o_cptrs[i] = cptrs[i]; o_rptrs[i] = rptrs[i]; }
  for (int i = 0; i < nnz; i++) {
    o_rows[i] = rows[i]; o_cvals[i] = cvals[i]; o_cols[i] = cols[i]; o_rvals[i] = rvals[i];
  }
  delete[] cptrs; delete[] rows; delete[] cvals; delete[] rptrs; delete[] cols; delete[] rvals;
}
void ref_matrix2compressed(double* mat, int nov, int nnz, int preprocessing, int* o_cptrs, int* o_rows,
                           double* o_cvals, int* o_rptrs, int* o_cols, double* o_rvals) {
  int *cptrs, *rows, *rptrs, *cols;
  double *cvals, *rvals;
  if (preprocessing == 1) matrix2compressed_sortOrder(mat, cptrs, rows, cvals, rptrs, cols, rvals, nov, nnz);
  else if (preprocessing == 2) matrix2compressed_skipOrder(mat, cptrs, rows, cvals, rptrs, cols, rvals, nov, nnz);
  else matrix2compressed(mat, cptrs, rows, cvals, rptrs, cols, rvals, nov, nnz);
  copy_out(nov, nnz, cptrs, rows, cvals, rptrs, cols, rvals, o_cptrs, o_rows, o_cvals, o_rptrs, o_cols, o_rvals);
}
// util.h:403 ; o_mat must hold (m*n/2)^2 ints; returns nnz (or -1)
int ref_gridGraph2compressed(int m, int n, int* o_mat, int* o_cptrs, int* o_rows, int* o_rptrs, int* o_cols) {
  int *mat, *cptrs, *rows, *rptrs, *cols;
  if (m % 2 == 1 && n % 2 == 1) return -1;
  int nnz = gridGraph2compressed(m, n, mat, cptrs, rows, rptrs, cols);
  int nov = m * n / 2;
  for (int i = 0; i < nov * nov; i++) o_mat[i] = mat[i];
  for (int i = 0; i <= nov; i++) { o_cptrs[i] = cptrs[i]; o_rptrs[i] = rptrs[i]; }
  for (int i = 0; i < nnz; i++) { o_rows[i] = rows[i]; o_cols[i] = cols[i]; }
  delete[] mat; delete[] cptrs; delete[] rows; delete[] rptrs; delete[] cols;
  return nnz;
}
// ReadMatrix (util.h:343) + header sniff (main.cu:494-498); returns nov, fills type (0 int,1 float,2 double)
int ref_read_matrix(const char* filename, int generic, double* o_mat, int cap, int* o_nnz, int* o_type) {
  int nov, nnz;
  string type;
  ifstream inFile(filename);
  if (!inFile) return -1;
  string line;
  getline(inFile, line);
  istringstream iss(line);
  iss >> nov >> nnz >> type;
  if (nov * nov > cap) return -2;
  *o_nnz = nnz;
  if (type == "int") {
    int* mat = new int[nov * nov]();
    ReadMatrix(mat, inFile, nov, generic != 0);
    for (int i = 0; i < nov * nov; i++) o_mat[i] = mat[i];
    delete[] mat; *o_type = 0;
  } else if (type == "float") {
    float* mat = new float[nov * nov]();
    ReadMatrix(mat, inFile, nov, generic != 0);
    for (int i = 0; i < nov * nov; i++) o_mat[i] = mat[i];
    delete[] mat; *o_type = 1;
  } else {
    double* mat = new double[nov * nov]();
    ReadMatrix(mat, inFile, nov, generic != 0);
    for (int i = 0; i < nov * nov; i++) o_mat[i] = mat[i];
    delete[] mat; *o_type = 2;
  }
  return nov;
}
int ref_max_threads() { return omp_get_max_threads(); }

}  // extern "C"
