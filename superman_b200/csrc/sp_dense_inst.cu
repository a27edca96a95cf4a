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

// resident blocks per SM of one instantiation (occupancy query; also forces the lazily loaded
// kernel into the context, so the first timed launch does not pay the module load)
template <int N, int B, int MB>
static int blocks_per_sm() {
  int dev = 0, v = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return MB;
  static int cached[64];
  if (cached[dev] > 0) return cached[dev];
  using L = DenseLayout<N, B>;
  constexpr size_t dyn = L::DYN ? sizeof(double) * L::TOTAL : 0;
  if (L::DYN && cudaFuncSetAttribute(ryser_reg_kernel<N, B, SPB_REG_THREADS, MB>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, SPB_SMEM_OPTIN_BYTES) != cudaSuccess) {
    (void)cudaGetLastError();
    return MB;
  }
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, ryser_reg_kernel<N, B, SPB_REG_THREADS, MB>,
                                                    SPB_REG_THREADS, dyn) != cudaSuccess || v < 1) {
    (void)cudaGetLastError();
    return MB;
  }
  cached[dev] = v;
  return v;
}

template <int N, int B, int MB>
static int launch_one(cudaStream_t st, const double* mat_t, const double* xbase, double* partials,
                      unsigned long long group_first, unsigned int n_groups, unsigned int* queue, int c,
                      int sm_count, unsigned* blocks_out) {
  if (n_groups == 0) return SPD_EINVAL;
  if (!mat_t) { *blocks_out = (unsigned)blocks_per_sm<N, B, MB>(); return SPD_OK; }   // prepare only
  // persistent grid: resident blocks only, groups are pulled from queue[0]
  unsigned blocks = (unsigned)sm_count * (unsigned)blocks_per_sm<N, B, MB>();
  if (blocks > n_groups) blocks = n_groups;
  using L = DenseLayout<N, B>;
  ryser_reg_kernel<N, B, SPB_REG_THREADS, MB>
      <<<blocks, SPB_REG_THREADS, L::DYN ? sizeof(double) * L::TOTAL : 0, st>>>(mat_t, xbase, partials, group_first, n_groups, queue, c);
  *blocks_out = blocks;
  return SPD_OK;
}

// Registers: X is 2N, the 2^B running products 2^(B+1); 4 blocks of 128 threads per SM leave 128
// registers per thread, 3 blocks 168, 2 blocks 255.  The build regenerates
// profiles/ptxas_dense_resource_usage.txt from the shipped library (make resource-report).
template <int N, int B>
static int launch_nb(cudaStream_t st, const double* mat_t, const double* xbase, double* partials,
                     unsigned long long group_first, unsigned int n_groups, unsigned int* queue, int c,
                     int sm_count, unsigned* blocks_out) {
  constexpr int MB = (B == 4) ? ((N <= 40) ? 4 : (N <= 59) ? 3 : 2) : ((N <= 43) ? 4 : 3);
  return launch_one<N, B, MB>(st, mat_t, xbase, partials, group_first, n_groups, queue, c, sm_count, blocks_out);
}

#define SPB_CASE(N)                                                                          \
  case N:                                                                                    \
    if (B == 3) return launch_nb<N, 3>(st, mat_t, xbase, partials, group_first, n_groups, queue, c, sm_count, blocks_out); \
    if (B == 4) return launch_nb<N, 4>(st, mat_t, xbase, partials, group_first, n_groups, queue, c, sm_count, blocks_out); \
    return SPD_ELIMIT;

#define SPB_GLUE2(a, b) a##b
#define SPB_GLUE(a, b) SPB_GLUE2(a, b)

extern "C" int SPB_GLUE(spb_reg_launch_g, SPB_GROUP)(
    int n, int B, cudaStream_t st, const double* mat_t, const double* xbase, double* partials,
    unsigned long long group_first, unsigned int n_groups, unsigned int* queue, int c, int sm_count,
    unsigned* blocks_out) {
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
