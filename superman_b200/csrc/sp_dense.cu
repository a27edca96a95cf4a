// Dense Ryser on one device: plan (matrix resident in HBM), range planner and the
// shared-memory-X kernel used for ragged range ends and for n outside the register kernels.
//
// Reference path replaced: gpu_perman64_xshared_coalescing_mshared and its kernel
// (gpu_exact_dense.cu:329-399, 640-699); the [lo, hi) contract is that of the kernel's
// (start, end) arguments, which the multi-GPU wrappers slice (gpu_exact_dense.cu:729-752, 786-889).
#include "sp_internal.cuh"
#include "sp_dense_reg.h"
#include "ryser_dd.cuh"
#include <stdlib.h>
#include <string.h>
#include <new>

namespace spb {

#define SMEMK_THREADS 128

// X in shared memory as doubles, x_j of thread t at X[j*THREADS + t]: consecutive lanes touch
// consecutive 8-byte words, so every access is conflict-free (the reference's float X with the
// same layout, gpu_exact_dense.cu:336,389, is what north_star's "shared memory for large n,
// laid out free of bank conflicts" refers to).  Each thread owns the contiguous index range
// [lo + t*per_thread, lo + (t+1)*per_thread) ∩ [lo, hi) and initialises X explicitly at its start
// (gpu_exact_dense.cu:363-371).
__global__ void __launch_bounds__(SMEMK_THREADS)
ryser_smem_kernel(const double* __restrict__ mat_t, const double* __restrict__ xbase, int n,
                  unsigned long long lo, unsigned long long hi, unsigned long long per_thread,
                  double* __restrict__ partials) {
  extern __shared__ __align__(16) double dsm[];
  double* colT = dsm;                 // colT[k*n + j] = A[j][k]
  double* X = dsm + n * n;            // X[j*THREADS + t]
  __shared__ double warp_part[SMEMK_THREADS / 32];
  for (int e = threadIdx.x; e < n * n; e += SMEMK_THREADS) colT[e] = mat_t[e];
  double* x = X + threadIdx.x;
  for (int j = 0; j < n; ++j) x[j * SMEMK_THREADS] = xbase[j];
  __syncthreads();

  const unsigned long long gid = (unsigned long long)blockIdx.x * SMEMK_THREADS + threadIdx.x;
  // gid * per_thread cannot overflow: the host keeps grid*per_thread < 2^63
  unsigned long long i = lo + gid * per_thread;
  unsigned long long end = i + per_thread;
  if (end > hi) end = hi;
  double acc = 0.0;
  if (i < end) {
    unsigned long long g = 0;
    if (i == 0) {
      double p0 = 1.0, p1 = 1.0;
      for (int j = 0; j + 1 < n; j += 2) { p0 *= x[j * SMEMK_THREADS]; p1 *= x[(j + 1) * SMEMK_THREADS]; }
      if (n & 1) p0 *= x[(n - 1) * SMEMK_THREADS];
      acc = p0 * p1;                  // NW base term, index 0
      i = 1;
    } else {
      g = (i - 1) ^ ((i - 1) >> 1);
      for (int k = 0; k < n - 1; ++k) {
        if ((g >> k) & 1ull) {
          const double* col = colT + k * n;
          for (int j = 0; j < n; ++j) x[j * SMEMK_THREADS] += col[j];
        }
      }
    }
    for (; i < end; ++i) {
      const int k = __ffsll((long long)i) - 1;
      g ^= (1ull << k);
      const double s = ((g >> k) & 1ull) ? 1.0 : -1.0;
      const double* col = colT + k * n;
      double p0 = 1.0, p1 = 1.0, p2 = 1.0, p3 = 1.0;
      int j = 0;
      for (; j + 3 < n; j += 4) {
        const double a0 = fma(s, col[j], x[j * SMEMK_THREADS]);
        const double a1 = fma(s, col[j + 1], x[(j + 1) * SMEMK_THREADS]);
        const double a2 = fma(s, col[j + 2], x[(j + 2) * SMEMK_THREADS]);
        const double a3 = fma(s, col[j + 3], x[(j + 3) * SMEMK_THREADS]);
        x[j * SMEMK_THREADS] = a0; x[(j + 1) * SMEMK_THREADS] = a1;
        x[(j + 2) * SMEMK_THREADS] = a2; x[(j + 3) * SMEMK_THREADS] = a3;
        p0 *= a0; p1 *= a1; p2 *= a2; p3 *= a3;
      }
      for (; j < n; ++j) {
        const double a0 = fma(s, col[j], x[j * SMEMK_THREADS]);
        x[j * SMEMK_THREADS] = a0;
        p0 *= a0;
      }
      const double prod = (p0 * p1) * (p2 * p3);
      acc += (i & 1ull) ? -prod : prod;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int w = 0; w < SMEMK_THREADS / 32; ++w) v += warp_part[w];
    partials[blockIdx.x] = v;
  }
}

// low columns walked inside a block of 2^B indices: 16 running products (B = 4) measured faster than 8
// for every order (profiles/r02_dense_b3_b4.log); SP_DENSE_LOWCOLS=3 selects the other build
static int dense_lowcols(int n) {
  (void)n;
  const int B = env_int("SP_DENSE_LOWCOLS", 0);
  return (B == 3) ? 3 : 4;
}

static int reg_launch(int n, int B, cudaStream_t st, const double* mat_t, const double* xbase,
                      double* partials, unsigned long long tile_first, unsigned int n_tiles,
                      unsigned int* queue, int c, int sm_count, unsigned* blocks_out) {
  switch (n % SPB_NGROUPS) {
    case 0: return spb_reg_launch_g0(n, B, st, mat_t, xbase, partials, tile_first, n_tiles, queue, c, sm_count, blocks_out);
    case 1: return spb_reg_launch_g1(n, B, st, mat_t, xbase, partials, tile_first, n_tiles, queue, c, sm_count, blocks_out);
    case 2: return spb_reg_launch_g2(n, B, st, mat_t, xbase, partials, tile_first, n_tiles, queue, c, sm_count, blocks_out);
    case 3: return spb_reg_launch_g3(n, B, st, mat_t, xbase, partials, tile_first, n_tiles, queue, c, sm_count, blocks_out);
    case 4: return spb_reg_launch_g4(n, B, st, mat_t, xbase, partials, tile_first, n_tiles, queue, c, sm_count, blocks_out);
    case 5: return spb_reg_launch_g5(n, B, st, mat_t, xbase, partials, tile_first, n_tiles, queue, c, sm_count, blocks_out);
    case 6: return spb_reg_launch_g6(n, B, st, mat_t, xbase, partials, tile_first, n_tiles, queue, c, sm_count, blocks_out);
    default: return spb_reg_launch_g7(n, B, st, mat_t, xbase, partials, tile_first, n_tiles, queue, c, sm_count, blocks_out);
  }
}

// ragged piece [lo, hi) through the shared-memory kernel; appends its blocks to the partial array
int enqueue_smem_range(Lane* L, const double* d_mat_t, const double* d_xbase, int n,
                       unsigned long long lo, unsigned long long hi, size_t* pcount, int* launches) {
  if (hi <= lo) return SPD_OK;
  const size_t smem_bytes = ((size_t)n * n + (size_t)n * SMEMK_THREADS) * sizeof(double);
  const unsigned long long len = hi - lo;
  const unsigned long long want_threads = (unsigned long long)L->sm_count * 4 * SMEMK_THREADS;
  unsigned long long per_thread = (len + want_threads - 1) / want_threads;
  unsigned long long pt = 16;       // at least 16 indices per thread: amortises the explicit X start
  while (pt < per_thread) pt <<= 1; // power of two keeps the flipped column warp-uniform
  per_thread = pt;
  const unsigned long long threads = (len + per_thread - 1) / per_thread;
  const unsigned long long blocks = (threads + SMEMK_THREADS - 1) / SMEMK_THREADS;
  int rc = lane_reserve_partials(L, *pcount + (size_t)blocks);
  if (rc != SPD_OK) return rc;
  ryser_smem_kernel<<<(unsigned)blocks, SMEMK_THREADS, smem_bytes, L->stream>>>(
      d_mat_t, d_xbase, n, lo, hi, per_thread, L->d_partials + *pcount);
  SPB_CUDA(cudaGetLastError());
  *pcount += (size_t)blocks;
  *launches += 1;
  return SPD_OK;
}

// opt in to > 48 KiB of dynamic shared memory when n needs it (the reference never does and
// silently fails to launch)
int smem_kernel_prepare(int n) {
  const size_t smem_bytes = ((size_t)n * n + (size_t)n * SMEMK_THREADS) * sizeof(double);
  if (smem_bytes > 40 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(ryser_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SPB_SMEM_OPTIN_BYTES);
    if (e != cudaSuccess) { set_error("smem opt-in (%zu B): %s", smem_bytes, cudaGetErrorString(e)); return SPD_ECUDA; }
  }
  return SPD_OK;
}

}  // namespace spb

using namespace spb;

static int g_quad = 0;      // spd_set_quad: plans created from now on compute in double-double

struct spd_dense_plan {
  Lane* lanep = nullptr;
  int n = 0;
  bool quad = false;
  double* d_mat_t = nullptr;
  double* d_xbase = nullptr;
  double* d_xbase_lo = nullptr;
  bool pending = false;
  spd_run_info info;
};

static int enqueue_smem(spd_dense_plan* p, unsigned long long lo, unsigned long long hi,
                        size_t* pcount, int* launches) {
  return enqueue_smem_range(p->lanep, p->d_mat_t, p->d_xbase, p->n, lo, hi, pcount, launches);
}

static int dense_enqueue(spd_dense_plan* p, unsigned long long lo, unsigned long long hi) {
  const int n = p->n;
  const unsigned long long full = 1ull << (n - 1);
  if (lo > hi || hi > full) {
    set_error("dense range [%llu, %llu) outside [0, 2^%d]", lo, hi, n - 1);
    return SPD_EINVAL;
  }
  Lane& L = *p->lanep;
  SPB_CUDA(cudaSetDevice(L.device));
  memset(&p->info, 0, sizeof(p->info));
  p->info.units = hi - lo;
  p->info.visited = hi - lo;
  p->info.path = SPD_PATH_DENSE_SMEM;
  SPB_CUDA(cudaEventRecord(L.ev0, L.stream));
  int launches = 0;
  bool first_reduce = true;
  size_t pcount = 0;
  const unsigned long long len = hi - lo;

  if (p->quad) {
    // double-double mode (-q): the 16-aligned body of the range goes through the block kernel (rows outside, 16
    // running products), the ragged ends (< 16 indices each) through the loop kernel; X in shared memory in both.
    // The block sums come back as (high, low) pairs and all of them go through the compensated reduction.
    const size_t smem_bytes = ((size_t)n * n + 2 * (size_t)n * DDK_THREADS) * sizeof(double);
    if (smem_bytes > 40 * 1024) {
      SPB_CUDA(cudaFuncSetAttribute(ryser_dd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SPB_SMEM_OPTIN_BYTES));
      SPB_CUDA(cudaFuncSetAttribute(ryser_dd_blk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SPB_SMEM_OPTIN_BYTES));
    }
    int bps = 1;
    SPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, ryser_dd_blk_kernel, DDK_THREADS, smem_bytes));
    unsigned long long blo = (lo + 15ull) & ~15ull, bhi = hi & ~15ull;
    if (n < 6 || bhi <= blo || bhi - blo < 16ull * DDK_THREADS || env_int("SP_DD_LOOP_KERNEL", 0) != 0) blo = bhi = lo;   // no body
    struct Part { unsigned long long lo, hi, per_thread, blocks; bool blk; };
    Part parts[3];
    int np = 0;
    size_t total = 0;
    if (bhi > blo) {
      const unsigned long long blen = bhi - blo;
      unsigned long long blocks = (unsigned long long)L.sm_count * (unsigned)(bps > 0 ? bps : 1) * 8ull;   // 8 waves: tail below 1 %
      unsigned long long per_thread = (blen + blocks * DDK_THREADS - 1) / (blocks * DDK_THREADS);
      unsigned long long pt = 16;
      while (pt < per_thread) pt <<= 1;                // power of two: the flipped column stays warp-uniform
      if (pt > 4096) pt = 4096;                        // short chains (accuracy is the point of this mode)
      blocks = (blen + pt * DDK_THREADS - 1) / (pt * DDK_THREADS);
      if (blocks > (1ull << 22)) { set_error("range too long for one double-double launch"); return SPD_ELIMIT; }
      parts[np++] = Part{blo, bhi, pt, blocks, true};
      total += 2 * (size_t)blocks;
    }
    const unsigned long long edge[2][2] = {{lo, bhi > blo ? blo : hi}, {bhi > blo ? bhi : hi, hi}};
    for (int e = 0; e < 2; ++e) {
      const unsigned long long elo = edge[e][0], ehi = edge[e][1];
      if (ehi <= elo) continue;
      const unsigned long long elen = ehi - elo;
      unsigned long long blocks = (unsigned long long)L.sm_count * (unsigned)(bps > 0 ? bps : 1) * 8ull;
      unsigned long long per_thread = (elen + blocks * DDK_THREADS - 1) / (blocks * DDK_THREADS);
      unsigned long long pt = 16;
      while (pt < per_thread) pt <<= 1;
      if (pt > 4096) pt = 4096;
      blocks = (elen + pt * DDK_THREADS - 1) / (pt * DDK_THREADS);
      if (blocks > (1ull << 22)) { set_error("range too long for one double-double launch"); return SPD_ELIMIT; }
      parts[np++] = Part{elo, ehi, pt, blocks, false};
      total += 2 * (size_t)blocks;
    }
    int rc2 = lane_reserve_partials(&L, total + 2);
    if (rc2 != SPD_OK) return rc2;
    size_t off = 0;
    for (int i = 0; i < np; ++i) {
      if (parts[i].blk)
        ryser_dd_blk_kernel<<<(unsigned)parts[i].blocks, DDK_THREADS, smem_bytes, L.stream>>>(
            p->d_mat_t, p->d_xbase, p->d_xbase_lo, n, parts[i].lo, parts[i].hi, parts[i].per_thread, L.d_partials + off);
      else
        ryser_dd_kernel<<<(unsigned)parts[i].blocks, DDK_THREADS, smem_bytes, L.stream>>>(
            p->d_mat_t, p->d_xbase, p->d_xbase_lo, n, parts[i].lo, parts[i].hi, parts[i].per_thread, L.d_partials + off);
      SPB_CUDA(cudaGetLastError());
      off += 2 * (size_t)parts[i].blocks;
    }
    if (total == 0) { SPB_CUDA(cudaMemsetAsync(L.d_partials, 0, sizeof(double), L.stream)); total = 1; }
    if ((rc2 = launch_reduce(L, L.d_partials, total, L.d_result, 0, false)) != SPD_OK) return rc2;
    SPB_CUDA(cudaMemcpyAsync(L.h_result, L.d_result, sizeof(double), cudaMemcpyDeviceToHost, L.stream));
    SPB_CUDA(cudaEventRecord(L.ev1, L.stream));
    p->info.launches = np + 1;
    p->info.path = SPD_PATH_DENSE_DD;
    p->pending = true;
    return SPD_OK;
  }

  const int B = dense_lowcols(n);
  const bool reg_ok = (n >= SPB_REG_NMIN && n <= SPB_REG_NMAX && env_int("SP_DENSE_FORCE_SMEM", 0) == 0);
  unsigned long long body_lo = lo, body_hi = lo;   // empty body by default
  int c = 0;
  if (reg_ok && len >= (1ull << (B + 1 + 7))) {
    // tile = 2^c indices per thread, group = 128 tiles.  Short tiles keep X's rounding drift small
    // (ryser_reg.cuh); c = 9 costs < 1 % for the explicit X starts.  Small ranges use shorter tiles
    // so that there are still enough groups to fill the SMs.
    c = env_int("SP_DENSE_TILE_LOG2", 0);
    if (c <= 0) {
      c = ilog2_ull(len) - 7 - 12;
      if (c > 9) c = 9;
    }
    if (c < B + 1) c = B + 1;
    if (c > n - 8) c = n - 8;
    if (c >= B + 1) {
      const unsigned long long G = 1ull << (c + 7);
      body_lo = (lo + G - 1) & ~(G - 1);
      body_hi = hi & ~(G - 1);
      if (body_hi <= body_lo) { body_lo = body_hi = lo; }
    }
  }

  int rc;
  if (body_hi > body_lo) {
    p->info.path = SPD_PATH_DENSE_REG;
    p->info.tile_log2 = c;
    unsigned long long group = body_lo >> (c + 7);
    unsigned long long groups_left = (body_hi - body_lo) >> (c + 7);
    // one partial sum per group; a launch covers at most 2^21 groups (16 MiB of partials), pulled
    // from the lane's atomic queue by a grid of resident blocks
    const unsigned long long max_groups = 1ull << 21;
    SPB_CUDA(cudaMemsetAsync(L.d_queue, 0, 2 * sizeof(unsigned int), L.stream));
    while (groups_left) {
      const unsigned long long ng = groups_left < max_groups ? groups_left : max_groups;
      if (pcount + ng > max_groups) {   // flush what we have
        rc = launch_reduce(L, L.d_partials, pcount, L.d_result, 0, !first_reduce);
        if (rc != SPD_OK) return rc;
        first_reduce = false; pcount = 0; ++launches;
      }
      rc = lane_reserve_partials(&L, pcount + (size_t)ng);
      if (rc != SPD_OK) return rc;
      unsigned nb = 0;
      rc = reg_launch(n, B, L.stream, p->d_mat_t, p->d_xbase, L.d_partials + pcount, group, (unsigned)ng,
                      L.d_queue, c, L.sm_count, &nb);
      if (rc != SPD_OK) { set_error("no register kernel for n=%d B=%d", n, B); return rc; }
      SPB_CUDA(cudaGetLastError());
      pcount += (size_t)ng; ++launches;
      group += ng; groups_left -= ng;
    }
    rc = enqueue_smem(p, lo, body_lo, &pcount, &launches);
    if (rc != SPD_OK) return rc;
    rc = enqueue_smem(p, body_hi, hi, &pcount, &launches);
    if (rc != SPD_OK) return rc;
  } else {
    rc = enqueue_smem(p, lo, hi, &pcount, &launches);
    if (rc != SPD_OK) return rc;
  }
  if (pcount > 0 || first_reduce) {
    rc = lane_reserve_partials(&L, 1);
    if (rc != SPD_OK) return rc;
    rc = launch_reduce(L, L.d_partials, pcount, L.d_result, 0, !first_reduce);
    if (rc != SPD_OK) return rc;
    ++launches;
  }
  SPB_CUDA(cudaMemcpyAsync(L.h_result, L.d_result, sizeof(double), cudaMemcpyDeviceToHost, L.stream));
  SPB_CUDA(cudaEventRecord(L.ev1, L.stream));
  p->info.launches = launches;
  p->pending = true;
  return SPD_OK;
}

extern "C" {

void spd_set_quad(int on) { g_quad = on ? 1 : 0; }
int spd_get_quad(void) { return g_quad; }

int spd_dense_plan_create(int device, const double* mat_t, const double* xbase, int nov,
                          spd_dense_plan** out) {
  if (!mat_t || !xbase || !out) { set_error("null argument"); return SPD_EINVAL; }
  if (nov < 2 || nov > 64) { set_error("dense Ryser supports 2 <= n <= 64 (got %d)", nov); return SPD_ELIMIT; }
  spd_dense_plan* p = new (std::nothrow) spd_dense_plan();
  if (!p) return SPD_ENOMEM;
  int rc = lane_acquire(device, &p->lanep);
  if (rc != SPD_OK) { delete p; return rc; }
  Lane& L = *p->lanep;
  p->n = nov;
  auto fail = [&](int code) { spd_dense_plan_destroy(p); return code; };
  if ((rc = lane_arena_alloc(&L, (size_t)nov * nov * sizeof(double), (void**)&p->d_mat_t)) != SPD_OK) return fail(rc);
  if ((rc = lane_arena_alloc(&L, (size_t)nov * sizeof(double), (void**)&p->d_xbase)) != SPD_OK) return fail(rc);
  p->quad = g_quad != 0;
  double xlo[64];
  if (p->quad) {
    // low word of the NW start vector: a[j][n-1] - rowsum_j / 2 in long double, minus the double the caller computed
    for (int j = 0; j < nov; ++j) {
      long double rs = 0.0L;
      for (int k = 0; k < nov; ++k) rs += (long double)mat_t[(size_t)k * nov + j];
      const long double x = (long double)mat_t[(size_t)(nov - 1) * nov + j] - rs / 2.0L;
      xlo[j] = (double)(x - (long double)xbase[j]);
    }
    if ((rc = lane_arena_alloc(&L, (size_t)nov * sizeof(double), (void**)&p->d_xbase_lo)) != SPD_OK) return fail(rc);
  }
  cudaError_t e;
  // pageable sources: the copies are staged by the runtime before returning, so the caller's
  // arrays are not referenced after this function
  if ((e = cudaSetDevice(device)) != cudaSuccess ||
      (e = cudaMemcpyAsync(p->d_mat_t, mat_t, (size_t)nov * nov * sizeof(double), cudaMemcpyHostToDevice, L.stream)) != cudaSuccess ||
      (e = cudaMemcpyAsync(p->d_xbase, xbase, (size_t)nov * sizeof(double), cudaMemcpyHostToDevice, L.stream)) != cudaSuccess ||
      (p->quad && ((e = cudaMemcpyAsync(p->d_xbase_lo, xlo, (size_t)nov * sizeof(double), cudaMemcpyHostToDevice, L.stream)) != cudaSuccess ||
                   (e = cudaStreamSynchronize(L.stream)) != cudaSuccess))) {
    set_error("dense plan upload: %s", cudaGetErrorString(e));
    return fail(SPD_ECUDA);
  }
  if ((rc = smem_kernel_prepare(nov)) != SPD_OK) return fail(rc);
  if (nov >= SPB_REG_NMIN && nov <= SPB_REG_NMAX) {
    // loads the instantiation this plan will launch (CUDA loads kernels lazily) and sets its
    // shared-memory opt-in, so that the first run does not pay for either
    const int B = dense_lowcols(nov);
    unsigned bps = 0;
    (void)reg_launch(nov, B, L.stream, nullptr, nullptr, nullptr, 0ull, 1u, nullptr, 0, L.sm_count, &bps);
  }
  rc = lane_reserve_partials(&L, (1u << 21) + 4096);
  if (rc != SPD_OK) return fail(rc);
  *out = p;
  return SPD_OK;
}

void spd_dense_plan_destroy(spd_dense_plan* p) {
  if (!p) return;
  lane_release(p->lanep);
  delete p;
}

int spd_dense_plan_launch(spd_dense_plan* p, unsigned long long lo, unsigned long long hi) {
  if (!p) { set_error("null plan"); return SPD_EINVAL; }
  if (p->pending) { set_error("plan already has a pending run"); return SPD_EINVAL; }
  return dense_enqueue(p, lo, hi);
}

int spd_dense_plan_wait(spd_dense_plan* p, double* sum, spd_run_info* info) {
  if (!p || !p->pending) { set_error("no pending run"); return SPD_EINVAL; }
  p->pending = false;
  SPB_CUDA(cudaSetDevice(p->lanep->device));
  SPB_CUDA(cudaEventSynchronize(p->lanep->ev1));
  float ms = 0.f;
  SPB_CUDA(cudaEventElapsedTime(&ms, p->lanep->ev0, p->lanep->ev1));
  p->info.kernel_ms = ms;
  if (sum) *sum = p->lanep->h_result[0];
  if (info) *info = p->info;
  return SPD_OK;
}

int spd_dense_plan_run(spd_dense_plan* p, unsigned long long lo, unsigned long long hi, double* sum,
                       spd_run_info* info) {
  int rc = spd_dense_plan_launch(p, lo, hi);
  if (rc != SPD_OK) return rc;
  return spd_dense_plan_wait(p, sum, info);
}

}  // extern "C"
