/* superman_b200_level.h -- the slot rules of the LevelRyser engine, shared by the C host code that packs a
 * matrix for it (superman_b200/host/sp_level.c) and the CUDA unit that instantiates its kernels
 * (superman_b200/csrc/level_reg.cuh): how many register-cold rows and how many resident blocks per SM go with
 * a slot configuration (S0 slots for level 0, S for each of the other B - 1 levels). */
#ifndef SUPERMAN_B200_LEVEL_H
#define SUPERMAN_B200_LEVEL_H

#ifdef __CUDACC__
#define SPL_FN __host__ __device__ constexpr
#elif defined(__cplusplus)
#define SPL_FN constexpr
#else
#define SPL_FN static inline
#endif

/* register rows of a configuration (hot slots) */
SPL_FN int spl_slots(int B, int S0, int S) { return S0 + (B - 1) * S; }

/* register-cold rows: 20 register rows fit 128 registers (4 blocks of 128 threads per SM), about 30 fit 168
 * (3 blocks); the SkipPer variant needs a few registers more (tile queue, votes) and gives up four
 * register-cold rows at the edge */
SPL_FN int spl_regcold(int B, int S0, int S, int skip) {
  return (spl_slots(B, S0, S) <= 12 ? 8 : spl_slots(B, S0, S) <= 16 ? 4 : spl_slots(B, S0, S) <= 24 ? 8 : 0) -
         ((skip && (spl_slots(B, S0, S) == 11 || spl_slots(B, S0, S) == 12 || spl_slots(B, S0, S) == 15 ||
                    spl_slots(B, S0, S) == 16)) ? 4 : 0) -
         ((!skip && ((B == 3 && spl_slots(B, S0, S) == 24) || spl_slots(B, S0, S) == 22)) ? 4 : 0);   /* 30+ register rows spill at 168 */
}

/* configurations with many hot slots read the low-column entries of their slots as uniform-register operands of
 * the FP64 instructions (kernel parameters) instead of loading two shared-memory images per block: measured
 * -3.5 % at 16 slots, +4 % at 11-12 (profiles/r02_level_uniform_operand_experiment.log); the 168-register
 * configurations (22+ slots) keep the images */
SPL_FN int spl_uniform_low(int hot_slots) { return hot_slots >= 13 && hot_slots <= 20; }

SPL_FN int spl_minblocks(int B, int S0, int S, int skip) {
  return spl_slots(B, S0, S) + spl_regcold(B, S0, S, skip) <= 20 ? 4
       : (spl_slots(B, S0, S) + spl_regcold(B, S0, S, skip) <= 30 || !skip) ? 3 : 2;
}

#endif
