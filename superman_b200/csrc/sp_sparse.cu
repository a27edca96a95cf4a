// SpaRyser / SkipPer on one device: plan (row-ordered matrix resident in HBM) and range planner.
//
// Reference paths replaced: gpu_perman64_xshared_coalescing_mshared_sparse + kernel
// (gpu_exact_sparse.cu:455-552, 853-914) and ..._mshared_skipper + kernel (:555-670, 1123-1190).
// The plan is given D, the matrix the reference's sparse kernels actually iterate over -- the CCS
// arrays scattered back to dense form (entries the CCS does not hold are 0) -- and the NW start
// vector; it orders the rows hot-first (see sparse_reg.cuh) once, at creation.
#include "sp_internal.cuh"
#include "sparse_reg.cuh"
#include "sp_dense_reg.h"
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <new>
#include <vector>

extern "C" {
#define SPB_DECL(g) \
  int spb_sparse_launch_g##g(int n, int B, int skip, cudaStream_t st, const spb::SparseArgs* a, unsigned blocks);
SPB_DECL(0) SPB_DECL(1) SPB_DECL(2) SPB_DECL(3) SPB_DECL(4) SPB_DECL(5) SPB_DECL(6) SPB_DECL(7)
#undef SPB_DECL
}

using namespace spb;

static int sparse_launch(int n, int B, int skip, cudaStream_t st, const SparseArgs* a, unsigned blocks) {
  switch (n % SPB_NGROUPS) {
    case 0: return spb_sparse_launch_g0(n, B, skip, st, a, blocks);
    case 1: return spb_sparse_launch_g1(n, B, skip, st, a, blocks);
    case 2: return spb_sparse_launch_g2(n, B, skip, st, a, blocks);
    case 3: return spb_sparse_launch_g3(n, B, skip, st, a, blocks);
    case 4: return spb_sparse_launch_g4(n, B, skip, st, a, blocks);
    case 5: return spb_sparse_launch_g5(n, B, skip, st, a, blocks);
    case 6: return spb_sparse_launch_g6(n, B, skip, st, a, blocks);
    default: return spb_sparse_launch_g7(n, B, skip, st, a, blocks);
  }
}

struct spd_sparse_plan {
  Lane* lanep = nullptr;
  int n = 0;
  int skip = 0;
  double* d_mat_t = nullptr;   // row-ordered, mat_t[k*n + j]
  double* d_xbase = nullptr;
  std::vector<int> level;      // sorted ascending: level[j] of the row now at position j
  bool pending = false;
  spd_run_info info;
};

static int count_level_below(const spd_sparse_plan* p, int bound) {
  return (int)(std::lower_bound(p->level.begin(), p->level.end(), bound) - p->level.begin());
}

static int sparse_enqueue(spd_sparse_plan* p, unsigned long long lo, unsigned long long hi) {
  const int n = p->n;
  const unsigned long long full = 1ull << (n - 1);
  if (lo > hi || hi > full) {
    set_error("sparse range [%llu, %llu) outside [0, 2^%d]", lo, hi, n - 1);
    return SPD_EINVAL;
  }
  Lane& L = *p->lanep;
  SPB_CUDA(cudaSetDevice(L.device));
  memset(&p->info, 0, sizeof(p->info));
  p->info.units = hi - lo;
  p->info.path = SPD_PATH_SPARSE_SMEM;
  SPB_CUDA(cudaEventRecord(L.ev0, L.stream));
  int launches = 0;
  size_t pcount = 0;
  const unsigned long long len = hi - lo;
  unsigned long long smem_indices = len;   // indices evaluated by the dense-semantics ragged kernel
  bool have_visited = false;

  const bool reg_ok = (n >= SPB_SPARSE_NMIN && n <= SPB_REG_NMAX && env_int("SP_SPARSE_FORCE_SMEM", 0) == 0);
  int rc;
  unsigned long long body_lo = lo, body_hi = lo;
  int c = 0, B = 3;
  if (reg_ok && len >= (1ull << 5)) {
    // low-column count: minimise FP64 instructions per Gray index,
    //   2*H_B (hot rows) + 2*(n - H_B)/2^B (cold rows) + ~8/2^B (block bookkeeping)
    double best = 1e300;
    for (int b = 3; b <= 4; ++b) {
      const int Hb = count_level_below(p, b);
      const double cost = 2.0 * Hb + (2.0 * (n - Hb) + 8.0) / (double)(1 << b);
      if (cost < best) { best = cost; B = b; }
    }
    const int forced = env_int("SP_SPARSE_LOWCOLS", 0);
    if (forced == 3 || forced == 4) B = forced;
    const int tiles_log2 = env_int("SP_SPARSE_TILES_LOG2", 21);
    c = env_int("SP_SPARSE_TILE_LOG2", 0);
    if (c <= 0) {
      c = ilog2_ull(len) - tiles_log2;
      if (c > 12) c = 12;
      if (c < 8) c = 8;
    }
    if (c < B + 1) c = B + 1;
    if (c > n - 1) c = n - 1;
    const unsigned long long T = 1ull << c;
    body_lo = (lo + T - 1) & ~(T - 1);
    body_hi = hi & ~(T - 1);
    if (body_hi <= body_lo) { body_lo = body_hi = lo; }
  }
  if (body_hi > body_lo) {
    p->info.path = p->skip ? SPD_PATH_SKIPPER : SPD_PATH_SPARSE_REG;
    p->info.tile_log2 = c;
    smem_indices = len - (body_hi - body_lo);
    const int H = count_level_below(p, B);
    const int TC = p->skip ? (n - count_level_below(p, c)) : 0;
    const int W = p->skip ? env_int("SP_SKIP_TILES_PER_LANE", 8) : 1;
    const int tiles_per_warp = 32 * (W < 1 ? 1 : W);
    const unsigned long long tiles_per_block = (unsigned long long)tiles_per_warp * (SPB_REG_THREADS / 32);
    unsigned long long tile = body_lo >> c;
    unsigned long long tiles_left = (body_hi - body_lo) >> c;
    const unsigned long long max_blocks = 1ull << 20;
    bool first = true;
    while (tiles_left) {
      unsigned long long blocks = (tiles_left + tiles_per_block - 1) / tiles_per_block;
      if (blocks > max_blocks) blocks = max_blocks;
      unsigned long long nt = blocks * tiles_per_block;
      if (nt > tiles_left) nt = tiles_left;
      if ((rc = lane_reserve_partials(&L, (size_t)blocks + 4096)) != SPD_OK) return rc;
      if ((rc = lane_reserve_aux(&L, (size_t)blocks)) != SPD_OK) return rc;
      SparseArgs a;
      a.mat_t = p->d_mat_t; a.xbase = p->d_xbase;
      a.partials = L.d_partials; a.visited = L.d_aux;
      a.tile_first = tile; a.n_tiles = nt; a.c = c; a.H = H; a.TC = TC; a.tiles_per_warp = tiles_per_warp;
      rc = sparse_launch(n, B, p->skip, L.stream, &a, (unsigned)blocks);
      if (rc != SPD_OK) { set_error("no sparse register kernel for n=%d B=%d", n, B); return rc; }
      SPB_CUDA(cudaGetLastError());
      ++launches;
      // fold this launch's per-block sums into result slots 0 (value) and 1 (visited blocks)
      if ((rc = launch_reduce(L, L.d_partials, (size_t)blocks, L.d_result, 0, !first)) != SPD_OK) return rc;
      if ((rc = launch_reduce_u64(L, L.d_aux, (size_t)blocks, L.d_result, 1, !first)) != SPD_OK) return rc;
      launches += 2;
      first = false;
      have_visited = true;
      tile += nt; tiles_left -= nt;
    }
    if ((rc = enqueue_smem_range(&L, p->d_mat_t, p->d_xbase, n, lo, body_lo, &pcount, &launches)) != SPD_OK) return rc;
    if ((rc = enqueue_smem_range(&L, p->d_mat_t, p->d_xbase, n, body_hi, hi, &pcount, &launches)) != SPD_OK) return rc;
    if (pcount > 0) {
      if ((rc = launch_reduce(L, L.d_partials, pcount, L.d_result, 0, true)) != SPD_OK) return rc;
      ++launches;
    }
    p->info.reserved = B;
  } else {
    if ((rc = enqueue_smem_range(&L, p->d_mat_t, p->d_xbase, n, lo, hi, &pcount, &launches)) != SPD_OK) return rc;
    if ((rc = lane_reserve_partials(&L, 1)) != SPD_OK) return rc;
    if ((rc = launch_reduce(L, L.d_partials, pcount, L.d_result, 0, false)) != SPD_OK) return rc;
    ++launches;
  }
  SPB_CUDA(cudaMemcpyAsync(L.h_result, L.d_result, 2 * sizeof(double), cudaMemcpyDeviceToHost, L.stream));
  SPB_CUDA(cudaEventRecord(L.ev1, L.stream));
  p->info.launches = launches;
  // visited is completed in wait(): blocks * 2^B + indices that went through the ragged kernel
  p->info.visited = smem_indices;
  p->info.reserved = have_visited ? B : 0;
  p->pending = true;
  return SPD_OK;
}

extern "C" {

int spd_sparse_plan_create(int device, const double* dmat_t, const double* xbase, int nov, int skip,
                           spd_sparse_plan** out) {
  if (!dmat_t || !xbase || !out) { set_error("null argument"); return SPD_EINVAL; }
  if (nov < 2 || nov > 64) { set_error("sparse Ryser supports 2 <= n <= 64 (got %d)", nov); return SPD_ELIMIT; }
  spd_sparse_plan* p = new (std::nothrow) spd_sparse_plan();
  if (!p) return SPD_ENOMEM;
  int rc = lane_acquire(device, &p->lanep);
  if (rc != SPD_OK) { delete p; return rc; }
  Lane& L = *p->lanep;
  p->n = nov;
  p->skip = skip ? 1 : 0;
  auto fail = [&](int code) { spd_sparse_plan_destroy(p); return code; };

  // row order: ascending by the lowest flippable column (0 .. n-2) holding a non-zero of the row;
  // rows touched by no such column come last.  Stable, so equal rows keep the caller's order.
  const int n = nov;
  std::vector<int> lvl(n), perm(n);
  for (int j = 0; j < n; ++j) {
    int l = n;
    for (int k = 0; k < n - 1; ++k)
      if (dmat_t[(size_t)k * n + j] != 0.0) { l = k; break; }
    lvl[j] = l;
    perm[j] = j;
  }
  std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return lvl[a] < lvl[b]; });
  std::vector<double> mt((size_t)n * n), xb(n);
  p->level.resize(n);
  for (int j = 0; j < n; ++j) {
    p->level[j] = lvl[perm[j]];
    xb[j] = xbase[perm[j]];
    for (int k = 0; k < n; ++k) mt[(size_t)k * n + j] = dmat_t[(size_t)k * n + perm[j]];
  }
  if ((rc = lane_arena_alloc(&L, (size_t)n * n * sizeof(double), (void**)&p->d_mat_t)) != SPD_OK) return fail(rc);
  if ((rc = lane_arena_alloc(&L, (size_t)n * sizeof(double), (void**)&p->d_xbase)) != SPD_OK) return fail(rc);
  cudaError_t e;
  if ((e = cudaSetDevice(device)) != cudaSuccess ||
      (e = cudaMemcpyAsync(p->d_mat_t, mt.data(), (size_t)n * n * sizeof(double), cudaMemcpyHostToDevice, L.stream)) != cudaSuccess ||
      (e = cudaMemcpyAsync(p->d_xbase, xb.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, L.stream)) != cudaSuccess ||
      (e = cudaStreamSynchronize(L.stream)) != cudaSuccess) {
    set_error("sparse plan upload: %s", cudaGetErrorString(e));
    return fail(SPD_ECUDA);
  }
  if ((rc = smem_kernel_prepare(nov)) != SPD_OK) return fail(rc);
  if ((rc = lane_reserve_partials(&L, (1u << 20) + 8192)) != SPD_OK) return fail(rc);
  if ((rc = lane_reserve_aux(&L, 1u << 20)) != SPD_OK) return fail(rc);
  *out = p;
  return SPD_OK;
}

void spd_sparse_plan_destroy(spd_sparse_plan* p) {
  if (!p) return;
  lane_release(p->lanep);
  delete p;
}

int spd_sparse_plan_launch(spd_sparse_plan* p, unsigned long long lo, unsigned long long hi) {
  if (!p) { set_error("null plan"); return SPD_EINVAL; }
  if (p->pending) { set_error("plan already has a pending run"); return SPD_EINVAL; }
  return sparse_enqueue(p, lo, hi);
}

int spd_sparse_plan_wait(spd_sparse_plan* p, double* sum, spd_run_info* info) {
  if (!p || !p->pending) { set_error("no pending run"); return SPD_EINVAL; }
  p->pending = false;
  SPB_CUDA(cudaSetDevice(p->lanep->device));
  SPB_CUDA(cudaEventSynchronize(p->lanep->ev1));
  float ms = 0.f;
  SPB_CUDA(cudaEventElapsedTime(&ms, p->lanep->ev0, p->lanep->ev1));
  p->info.kernel_ms = ms;
  if (p->info.reserved) {
    unsigned long long blocks = 0;
    memcpy(&blocks, &p->lanep->h_result[1], sizeof(blocks));
    p->info.visited += blocks << p->info.reserved;
  }
  if (sum) *sum = p->lanep->h_result[0];
  if (info) *info = p->info;
  return SPD_OK;
}

int spd_sparse_plan_run(spd_sparse_plan* p, unsigned long long lo, unsigned long long hi, double* sum,
                        spd_run_info* info) {
  int rc = spd_sparse_plan_launch(p, lo, hi);
  if (rc != SPD_OK) return rc;
  return spd_sparse_plan_wait(p, sum, info);
}

}  // extern "C"
