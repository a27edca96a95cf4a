// ref_shim_revised.cpp -- exposes the structural preprocessing of the reference's REVISED front-end
// (revised_perman/util.h, compiled where it lies under /root/reference) through a C ABI so the tests
// can pin sp_matrix_reduce / sp_matrix_split34 / sp_matrix_scale / sp_matrix_dm against it.
// TEST INFRASTRUCTURE ONLY; output goes to oracle/_ref/ (git-ignored).  No reference source is copied
// into this repository: this file only #includes it at build time.
#include <cstdint>
#include <cstring>
#include "util.h"     // revised_perman/util.h (include path set by the Makefile)

extern "C" {

int rev_min_nnz(const double* mat, int nov) { return getMinNnz((double*)mat, nov); }      // util.h:1181

// util.h:1200 / 1260 -- in place on a caller-owned nov*nov buffer; returns the new nov (unchanged
// when nothing applied)
int rev_d1compress(double* mat, int nov) { d1compress(mat, nov); return nov; }
int rev_d2compress(double* mat, int nov) { d2compress(mat, nov); return nov; }

// util.h:1333 -- mat (nov*nov) becomes the first (nov-1)^2 matrix, mat2_out receives the second
int rev_d34compress(double* mat, int nov, double* mat2_out, int minDeg) {
  double* mat2 = nullptr; int nov2 = 0;
  if (!d34compress(mat, nov, mat2, nov2, minDeg)) return nov;
  memcpy(mat2_out, mat2, sizeof(double) * nov2 * nov2);
  delete[] mat2;
  return nov;
}

// util.h:1445 -- Sinkhorn-Knopp companion vectors for a CRS+CCS matrix
void rev_scalesk(int nov, int nnz, int* cptrs, int* rows, double* cvals, int* rptrs, int* cols,
                 double* rvals, double threshold, double* rv_out, double* cv_out) {
  SparseMatrix<double> sm;
  sm.cptrs = cptrs; sm.rows = rows; sm.cvals = cvals; sm.rptrs = rptrs; sm.cols = cols; sm.rvals = rvals;
  sm.nov = nov; sm.nnz = nnz;
  flags f; f.scaling_threshold = threshold;
  ScaleCompanion<double>* sc = scalesk(&sm, f);
  for (int i = 0; i < nov; i++) { rv_out[i] = sc->r_v[i]; cv_out[i] = sc->c_v[i]; }
  sm.cptrs = sm.rows = sm.rptrs = sm.cols = nullptr; sm.cvals = sm.rvals = nullptr;
}

}  // extern "C"
