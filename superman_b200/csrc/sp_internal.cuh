// Internal helpers shared by the CUDA translation units of libsuperman_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdint.h>
#include "superman_b200_device.h"

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
// 256-byte aligned sub-allocation from the lane's device arena (reset by lane_release)
int lane_arena_alloc(Lane* lane, size_t bytes, void** ptr);

// Deterministic reduction of `count` doubles to out[slot] (fixed summation tree, compensated).
int launch_reduce(const Lane& lane, const double* partials, size_t count, double* out, int slot,
                  bool accumulate);

int check_device(int device);

}  // namespace spb
