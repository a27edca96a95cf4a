// Instantiations of the SpaRyser / SkipPer register kernel; compiled SPB_NGROUPS times with
// -DSPB_GROUP=g, unit g holding n % SPB_NGROUPS == g (see sp_dense_inst.cu).
#include "sp_internal.cuh"
#include "sparse_reg.cuh"
#include "sp_dense_reg.h"

#ifndef SPB_GROUP
#error "compile with -DSPB_GROUP=<0..SPB_NGROUPS-1>"
#endif

namespace spb {

template <int N, int B, bool SKIP>
static int launch_sparse(cudaStream_t st, const SparseArgs& a, unsigned blocks) {
  constexpr int MB = (B == 4) ? (N <= (SKIP ? 28 : 30) ? 4 : 3) : (N <= 36 ? 4 : 3);
  if (blocks == 0) {            // prepare only: load the (lazily loaded) kernel into the current context
    cudaFuncAttributes fa;
    (void)cudaFuncGetAttributes(&fa, sparse_reg_kernel<N, B, SPB_REG_THREADS, MB, SKIP>);
    (void)cudaGetLastError();
    return SPD_OK;
  }
  sparse_reg_kernel<N, B, SPB_REG_THREADS, MB, SKIP><<<blocks, SPB_REG_THREADS, 0, st>>>(a);
  return SPD_OK;
}

#define SPB_CASE(N)                                                            \
  case N:                                                                      \
    if (B == 3) return skip ? launch_sparse<N, 3, true>(st, *a, blocks) : launch_sparse<N, 3, false>(st, *a, blocks); \
    if (B == 4) return skip ? launch_sparse<N, 4, true>(st, *a, blocks) : launch_sparse<N, 4, false>(st, *a, blocks); \
    return SPD_ELIMIT;

#define SPB_GLUE2(a, b) a##b
#define SPB_GLUE(a, b) SPB_GLUE2(a, b)

extern "C" int SPB_GLUE(spb_sparse_launch_g, SPB_GROUP)(int n, int B, int skip, cudaStream_t st,
                                                        const SparseArgs* a, unsigned blocks) {
  switch (n) {
#if SPB_GROUP == 0
    SPB_CASE(8) SPB_CASE(16) SPB_CASE(24) SPB_CASE(32) SPB_CASE(40) SPB_CASE(48)
#elif SPB_GROUP == 1
    SPB_CASE(9) SPB_CASE(17) SPB_CASE(25) SPB_CASE(33) SPB_CASE(41)
#elif SPB_GROUP == 2
    SPB_CASE(10) SPB_CASE(18) SPB_CASE(26) SPB_CASE(34) SPB_CASE(42)
#elif SPB_GROUP == 3
    SPB_CASE(11) SPB_CASE(19) SPB_CASE(27) SPB_CASE(35) SPB_CASE(43)
#elif SPB_GROUP == 4
    SPB_CASE(12) SPB_CASE(20) SPB_CASE(28) SPB_CASE(36) SPB_CASE(44)
#elif SPB_GROUP == 5
    SPB_CASE(13) SPB_CASE(21) SPB_CASE(29) SPB_CASE(37) SPB_CASE(45)
#elif SPB_GROUP == 6
    SPB_CASE(14) SPB_CASE(22) SPB_CASE(30) SPB_CASE(38) SPB_CASE(46)
#elif SPB_GROUP == 7
    SPB_CASE(7) SPB_CASE(15) SPB_CASE(23) SPB_CASE(31) SPB_CASE(39) SPB_CASE(47)
#endif
    default:
      return SPD_ELIMIT;
  }
}

}  // namespace spb
