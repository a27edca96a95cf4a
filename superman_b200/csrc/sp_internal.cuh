// Internal helpers shared by the CUDA translation units of libsuperman_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdint.h>
#include "superman_b200_device.h"

// Opt-in size for kernels that need more than the default 48 KiB of dynamic shared memory.  The
// attribute is per function and per context -- shared by every host thread that launches the kernel
// on that device -- so it is always set to the same constant (the 227 KiB opt-in maximum minus room
// for the kernels' static shared memory), never to the size one particular launch needs: two
// threads preparing launches of different sizes cannot lower it under each other's feet.
#define SPB_SMEM_OPTIN_BYTES (227 * 1024 - 4096)

namespace spb {

void set_error(const char* fmt, ...);

#define SPB_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      ::spb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return (e__ == cudaErrorMemoryAllocation) ? SPD_ENOMEM : SPD_ECUDA;                \
    }                                                                                    \
  } while (0)

// A stream, two events and a pinned result slot: what every plan needs to run on its device
// without touching cudaMalloc / cudaFree on the timed path (last row).
struct Lane {
  int device = -1;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double* h_result = nullptr;   // pinned, 8 doubles
  double* d_result = nullptr;   // 8 doubles
  double* d_partials = nullptr; // per-block partial sums
  size_t partials_cap = 0;      // in doubles
  unsigned long long* d_aux = nullptr;   // per-block integer counters (visited blocks)
  size_t aux_cap = 0;
  unsigned int* d_queue = nullptr;       // {next work item, blocks done} of the persistent kernels; 0 between launches
  char* d_arena = nullptr;      // plan inputs (matrix, CRS/CCS, ...) live here
  size_t arena_cap = 0, arena_used = 0;
  int sm_count = 0;
  bool busy = false;
};

// Lanes are pooled per device and live until spd_shutdown(): a plan borrows one, so opening a
// plan on a warm device costs one small H2D copy and no cudaMalloc / cudaMallocHost / stream or
// event creation.
int lane_acquire(int device, Lane** lane);
void lane_release(Lane* lane);
int lane_reserve_partials(Lane* lane, size_t count);
int lane_reserve_aux(Lane* lane, size_t count);
// 256-byte aligned sub-allocation from the lane's device arena (reset by lane_release)
int lane_arena_alloc(Lane* lane, size_t bytes, void** ptr);

// Deterministic reduction of `count` doubles to out[slot] (fixed summation tree, compensated).
int launch_reduce(const Lane& lane, const double* partials, size_t count, double* out, int slot,
                  bool accumulate);

// out[slot] (as u64) = (accumulate ? out[slot] : 0) + sum of counts
int launch_reduce_u64(const Lane& lane, const unsigned long long* counts, size_t count, double* out,
                      int slot, bool accumulate);

// estimators: out[0] = sum, out[1] = sum of (est * scale)^2, out[2] = number of non-zero estimates (u64 bits)
int launch_reduce_estimates(const Lane& lane, const double* est, size_t count, double scale, double* out, bool accumulate);

int check_device(int device);

// shared-memory-X dense kernel on [lo, hi) (sp_dense.cu); appends its blocks at partials[*pcount]
int enqueue_smem_range(Lane* L, const double* d_mat_t, const double* d_xbase, int n,
                       unsigned long long lo, unsigned long long hi, size_t* pcount, int* launches);
int smem_kernel_prepare(int n);
int env_int(const char* name, int dflt);
int ilog2_ull(unsigned long long v);

}  // namespace spb
