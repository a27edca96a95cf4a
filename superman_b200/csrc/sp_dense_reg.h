// Dispatch interface between sp_dense.cu and the instantiation units sp_dense_inst.cu.
#pragma once
#include <cuda_runtime.h>

#define SPB_REG_THREADS 128
#define SPB_NGROUPS 8
#define SPB_REG_NMIN 13
#define SPB_REG_NMAX 64
#define SPB_SPARSE_NMIN 7
#define SPB_SPARSE_NMAX 48

// mat_t == NULL: no launch, *blocks_out = resident blocks per SM of the instantiation (and the
// kernel is loaded into the current context)
extern "C" {
#define SPB_DECL(g)                                                                               \
  int spb_reg_launch_g##g(int n, int B, cudaStream_t st, const double* mat_t, const double* xbase, \
                          double* partials, unsigned long long group_first,                       \
                          unsigned int n_groups, unsigned int* queue, int c, int sm_count,        \
                          unsigned* blocks_out);
SPB_DECL(0) SPB_DECL(1) SPB_DECL(2) SPB_DECL(3) SPB_DECL(4) SPB_DECL(5) SPB_DECL(6) SPB_DECL(7)
#undef SPB_DECL
}
