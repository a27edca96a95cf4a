// Instantiations of the register-resident dense Ryser kernel.  Compiled SPB_NGROUPS times with
// -DSPB_GROUP=g (g = 0 .. SPB_NGROUPS-1); unit g holds the kernels for n % SPB_NGROUPS == g so
// that the ~90 instantiations build in parallel.
#include "sp_internal.cuh"
#include "ryser_reg.cuh"
#include "sp_dense_reg.h"

#ifndef SPB_GROUP
#error "compile with -DSPB_GROUP=<0..SPB_NGROUPS-1>"
#endif

namespace spb {

template <int N, int B, int MB>
static int launch_one(cudaStream_t st, const double* mat_t, const double* xbase, double* partials,
                      unsigned long long group_first, unsigned long long n_groups, int gpb, int c,
                      unsigned* blocks_out) {
  const unsigned long long blocks = (n_groups + (unsigned)gpb - 1) / (unsigned)gpb;
  if (blocks == 0 || blocks > 0x7fffffffull) return SPD_EINVAL;
  ryser_reg_kernel<N, B, SPB_REG_THREADS, MB>
      <<<(unsigned)blocks, SPB_REG_THREADS, 0, st>>>(mat_t, xbase, partials, group_first, n_groups, gpb, c);
  *blocks_out = (unsigned)blocks;
  return SPD_OK;
}

// Registers: X is 2N, the 2^B running products 2^(B+1); 4 blocks of 128 threads per SM leave 128
// registers per thread, 3 blocks 168, 2 blocks 255 (ptxas -v: zero spills for every entry below).
template <int N, int B>
static int launch_nb(cudaStream_t st, const double* mat_t, const double* xbase, double* partials,
                     unsigned long long group_first, unsigned long long n_groups, int gpb, int c,
                     unsigned* blocks_out) {
  // blocks per SM from the ptxas -v survey of every (N, B, MB) and a timing of every order
  // (profiles/r01_ptxas_spill_survey_dense.txt, r01_dense_frac_vs_n.log): the largest occupancy without
  // spills, except where a spill of <= 60 bytes outside the inner product loop measured faster than the
  // next lower occupancy (n = 33, 40: 4 blocks; n = 55..59: 3 blocks)
  constexpr int MB = (B == 4) ? ((N <= 40) ? 4 : (N <= 59) ? 3 : 2) : ((N > 54) ? 2 : (N <= 43 ? 4 : 3));
  return launch_one<N, B, MB>(st, mat_t, xbase, partials, group_first, n_groups, gpb, c, blocks_out);
}

#define SPB_CASE(N)                                                                          \
  case N:                                                                                    \
    if (B == 3) return launch_nb<N, 3>(st, mat_t, xbase, partials, group_first, n_groups, gpb, c, blocks_out); \
    if (B == 4) return launch_nb<N, 4>(st, mat_t, xbase, partials, group_first, n_groups, gpb, c, blocks_out); \
    return SPD_ELIMIT;

#define SPB_GLUE2(a, b) a##b
#define SPB_GLUE(a, b) SPB_GLUE2(a, b)

extern "C" int SPB_GLUE(spb_reg_launch_g, SPB_GROUP)(
    int n, int B, cudaStream_t st, const double* mat_t, const double* xbase, double* partials,
    unsigned long long group_first, unsigned long long n_groups, int gpb, int c, unsigned* blocks_out) {
  switch (n) {
#if SPB_GROUP == 0
    SPB_CASE(16) SPB_CASE(24) SPB_CASE(32) SPB_CASE(40) SPB_CASE(48) SPB_CASE(56) SPB_CASE(64)
#elif SPB_GROUP == 1
    SPB_CASE(17) SPB_CASE(25) SPB_CASE(33) SPB_CASE(41) SPB_CASE(49) SPB_CASE(57)
#elif SPB_GROUP == 2
    SPB_CASE(18) SPB_CASE(26) SPB_CASE(34) SPB_CASE(42) SPB_CASE(50) SPB_CASE(58)
#elif SPB_GROUP == 3
    SPB_CASE(19) SPB_CASE(27) SPB_CASE(35) SPB_CASE(43) SPB_CASE(51) SPB_CASE(59)
#elif SPB_GROUP == 4
    SPB_CASE(20) SPB_CASE(28) SPB_CASE(36) SPB_CASE(44) SPB_CASE(52) SPB_CASE(60)
#elif SPB_GROUP == 5
    SPB_CASE(13) SPB_CASE(21) SPB_CASE(29) SPB_CASE(37) SPB_CASE(45) SPB_CASE(53) SPB_CASE(61)
#elif SPB_GROUP == 6
    SPB_CASE(14) SPB_CASE(22) SPB_CASE(30) SPB_CASE(38) SPB_CASE(46) SPB_CASE(54) SPB_CASE(62)
#elif SPB_GROUP == 7
    SPB_CASE(15) SPB_CASE(23) SPB_CASE(31) SPB_CASE(39) SPB_CASE(47) SPB_CASE(55) SPB_CASE(63)
#endif
    default:
      return SPD_ELIMIT;
  }
}

}  // namespace spb
