// Instantiations of the LevelRyser kernel (level_reg.cuh); compiled four times with
// -DSPB_LV_B={3,4} -DSPB_LV_SKIP={0,1}.
#include "sp_internal.cuh"
#include "level_reg.cuh"
#include "sp_dense_reg.h"

#if !defined(SPB_LV_B) || !defined(SPB_LV_SKIP)
#error "compile with -DSPB_LV_B=<3|4> -DSPB_LV_SKIP=<0|1>"
#endif

namespace spb {

template <int B, int S0, int S, bool SKIP>
static int launch_level(cudaStream_t st, const LevelArgs& a, int sm_count, size_t smem, unsigned* blocks_out) {
  constexpr int R = spl_regcold(B, S0, S, SKIP), MB = spl_minblocks(B, S0, S, SKIP);
  auto kern = level_reg_kernel<B, S0, S, R, SPB_REG_THREADS, MB, SKIP>;
  // always the same constant (see SPB_SMEM_OPTIN_BYTES); also loads the lazily loaded kernel
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SPB_SMEM_OPTIN_BYTES);
  if (e != cudaSuccess) { set_error("smem opt-in (%zu B): %s", smem, cudaGetErrorString(e)); return SPD_ECUDA; }
  int bps = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, SPB_REG_THREADS, smem);
  if (e != cudaSuccess || bps < 1) {
    set_error("level kernel B=%d S0=%d S=%d does not fit an SM with %zu B of shared memory", B, S0, S, smem);
    (void)cudaGetLastError();
    return SPD_ELIMIT;
  }
  if (!a.partials) { *blocks_out = (unsigned)bps; return SPD_OK; }              // prepare only
  // persistent grid: resident blocks only, every warp pulls chunks of tiles from a.queue
  unsigned long long blocks = (unsigned long long)sm_count * (unsigned)bps;
  const unsigned long long need = ((unsigned long long)a.n_chunks + SPB_REG_THREADS / 32 - 1) / (SPB_REG_THREADS / 32);
  if (blocks > need) blocks = need;
  kern<<<(unsigned)blocks, SPB_REG_THREADS, smem, st>>>(a);
  *blocks_out = (unsigned)blocks;
  return SPD_OK;
}

#define SPB_GLUE3(a, b, c, d) a##b##c##d
#define SPB_GLUE(a, b, c, d) SPB_GLUE3(a, b, c, d)

// level 0 gets S0 = S, S-1 or S-2 slots (the first column of a sorted matrix is its sparsest)
#define SPB_LV_CASE(S_)                                                                                   \
  case S_:                                                                                                \
    if (S0 == S_) return launch_level<B, S_, S_, SKIP>(st, *a, sm_count, smem, blocks_out);               \
    if (S_ > 1 && S0 == S_ - 1) return launch_level<B, (S_ > 1 ? S_ - 1 : 1), S_, SKIP>(st, *a, sm_count, smem, blocks_out); \
    if (S_ > 2 && S0 == S_ - 2) return launch_level<B, (S_ > 2 ? S_ - 2 : 1), S_, SKIP>(st, *a, sm_count, smem, blocks_out); \
    return SPD_ELIMIT;

extern "C" int SPB_GLUE(spb_level_launch_b, SPB_LV_B, _s, SPB_LV_SKIP)(int S0, int S, cudaStream_t st, const LevelArgs* a,
                                                                      int sm_count, size_t smem, unsigned* blocks_out) {
  constexpr int B = SPB_LV_B;
  constexpr bool SKIP = SPB_LV_SKIP != 0;
  switch (S) {
    SPB_LV_CASE(1) SPB_LV_CASE(2) SPB_LV_CASE(3) SPB_LV_CASE(4) SPB_LV_CASE(6) SPB_LV_CASE(8)
    default: return SPD_ELIMIT;
  }
}

}  // namespace spb
