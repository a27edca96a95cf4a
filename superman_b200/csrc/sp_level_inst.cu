// Instantiations of the LevelRyser kernel (level_reg.cuh); compiled four times with
// -DSPB_LV_B={3,4} -DSPB_LV_SKIP={0,1}.
#include "sp_internal.cuh"
#include "level_reg.cuh"
#include "sp_dense_reg.h"

#if !defined(SPB_LV_B) || !defined(SPB_LV_SKIP)
#error "compile with -DSPB_LV_B=<3|4> -DSPB_LV_SKIP=<0|1>"
#endif

namespace spb {

template <int B, int S, int R, bool SKIP>
static int launch_level(cudaStream_t st, const LevelArgs& a, unsigned blocks, size_t smem) {
  // registers: X of the slots 2*(B*S + R), level products 2*(2^(B+1)-2): 128 registers up to 16 slots
  constexpr int MB = (B * S + R <= 16 && !(B == 4 && S == 4)) ? 4 : 3;
  auto kern = level_reg_kernel<B, S, R, SPB_REG_THREADS, MB, SKIP>;
  if (smem > 40 * 1024) {   // static shared memory (queue, partials) counts against the 48 KiB default too
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SPB_SMEM_OPTIN_BYTES);
    if (e != cudaSuccess) { set_error("smem opt-in (%zu B): %s", smem, cudaGetErrorString(e)); return SPD_ECUDA; }
  }
  kern<<<blocks, SPB_REG_THREADS, smem, st>>>(a);
  return SPD_OK;
}

#define SPB_GLUE3(a, b, c, d) a##b##c##d
#define SPB_GLUE(a, b, c, d) SPB_GLUE3(a, b, c, d)

#define SPB_LV_CASE(S_)                                                              \
  case S_:                                                                           \
    if (R == 0) return launch_level<B, S_, 0, SKIP>(st, *a, blocks, smem);           \
    if (R == 4) return launch_level<B, S_, 4, SKIP>(st, *a, blocks, smem);           \
    if (R == 8) return launch_level<B, S_, 8, SKIP>(st, *a, blocks, smem);           \
    return SPD_ELIMIT;

extern "C" int SPB_GLUE(spb_level_launch_b, SPB_LV_B, _s, SPB_LV_SKIP)(int S, int R, cudaStream_t st, const LevelArgs* a,
                                                                      unsigned blocks, size_t smem) {
  constexpr int B = SPB_LV_B;
  constexpr bool SKIP = SPB_LV_SKIP != 0;
  switch (S) {
    SPB_LV_CASE(1) SPB_LV_CASE(2) SPB_LV_CASE(3) SPB_LV_CASE(4) SPB_LV_CASE(6) SPB_LV_CASE(8)
    default: return SPD_ELIMIT;
  }
}

}  // namespace spb
