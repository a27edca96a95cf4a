// SpaRyser / SkipPer on one device: plan (row-ordered matrix resident in HBM) and range planner.
//
// Reference paths replaced: gpu_perman64_xshared_coalescing_mshared_sparse + kernel
// (gpu_exact_sparse.cu:455-552, 853-914) and ..._mshared_skipper + kernel (:555-670, 1123-1190).
// The plan is given D, the matrix the reference's sparse kernels actually iterate over -- the CCS
// arrays scattered back to dense form (entries the CCS does not hold are 0) -- and the NW start
// vector, with the rows already ordered by level and the level engine's images already packed by the C host
// (host/sp_level.c); this unit uploads, launches and reduces.
#include "sp_internal.cuh"
#include "sparse_reg.cuh"
#include "level_reg.cuh"
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

extern "C" {
// a->partials == NULL: no launch, *blocks_out = resident blocks per SM (and the kernel is loaded)
int spb_level_launch_b3_s0(int S0, int S, cudaStream_t st, const spb::LevelArgs* a, int sm_count, size_t smem, unsigned* blocks_out);
int spb_level_launch_b3_s1(int S0, int S, cudaStream_t st, const spb::LevelArgs* a, int sm_count, size_t smem, unsigned* blocks_out);
int spb_level_launch_b4_s0(int S0, int S, cudaStream_t st, const spb::LevelArgs* a, int sm_count, size_t smem, unsigned* blocks_out);
int spb_level_launch_b4_s1(int S0, int S, cudaStream_t st, const spb::LevelArgs* a, int sm_count, size_t smem, unsigned* blocks_out);
}

using namespace spb;

static int level_launch(int B, int S0, int S, int skip, cudaStream_t st, const LevelArgs* a, int sm_count, size_t smem,
                        unsigned* blocks_out) {
  if (B == 3) return skip ? spb_level_launch_b3_s1(S0, S, st, a, sm_count, smem, blocks_out) : spb_level_launch_b3_s0(S0, S, st, a, sm_count, smem, blocks_out);
  if (B == 4) return skip ? spb_level_launch_b4_s1(S0, S, st, a, sm_count, smem, blocks_out) : spb_level_launch_b4_s0(S0, S, st, a, sm_count, smem, blocks_out);
  return SPD_ELIMIT;
}

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
  // LevelRyser image (level_reg.cuh); lvB == 0 when the matrix does not fit its slots
  int lvB = 0, lvS0 = 0, lvS = 0, lvR = 0, NC = 0, NCP = 0, HSP = 0;
  int skip_long_tiles = 0;     // host's choice between 2^21 and 2^20 tiles per whole-space SkipPer run
  double lv_instr = 0.0;       // FP64 instructions per Gray index of the chosen engine (host model)
  double *d_colT_hot = nullptr, *d_lowR = nullptr, *d_dcold = nullptr, *d_xb_hot = nullptr, *d_xb_cold = nullptr;
  int cold_start[SPB_LV_MAXSEG] = {0};
  double low[SPB_LV_MAXLOW] = {0.0};
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

  const bool use_level = p->lvB != 0;
  const bool reg_ok = (use_level || (n >= SPB_SPARSE_NMIN && n <= SPB_SPARSE_NMAX)) && env_int("SP_SPARSE_FORCE_SMEM", 0) == 0;
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
    if (use_level) B = p->lvB;
    // tile size: SpaRyser is fastest with 2^12-index tiles at n = 33 (less start-up work per index),
    // SkipPer with 2^11 (its tile filter drops more when tiles are finer): 2^20 resp. 2^21 tiles per launch
    const int tiles_log2 = env_int("SP_SPARSE_TILES_LOG2", (p->skip && !p->skip_long_tiles) ? 21 : 20);
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
      if (use_level) {
        // persistent grid: every warp pulls chunks of tiles_per_warp tiles from the lane's atomic queue and
        // leaves one partial sum (and one visited count) per chunk; at most 2^20 chunks per launch.  The last
        // eighth of a SkipPer launch goes out in tiles a quarter as long, so that the warps finish together: its
        // chunks are 8 tiles per lane, and a chunk of full-length tiles is 1/14 of a warp's share at config-3 size.
        // (Not for SpaRyser, whose one-tile chunks lose less at the end than short tiles cost, and not below
        // 2^11-index tiles, where the short ones would be mostly prologue.)
        int c_small = p->skip ? c - env_int("SP_LEVEL_TAIL_SHIFT", 2) : c;
        if (c_small < B + 1) c_small = B + 1;
        unsigned long long nt = tiles_left;                          // tiles of 2^c taken by this launch
        unsigned long long big = nt, small = 0;
        if (c_small < c && c >= 11 && nt >= 8ull * (unsigned)tiles_per_warp) {
          big = (nt - nt / 8) / (unsigned)tiles_per_warp * (unsigned)tiles_per_warp;
          small = (nt - big) << (c - c_small);
        }
        unsigned long long chunks_big = (big + (unsigned)tiles_per_warp - 1) / (unsigned)tiles_per_warp;
        unsigned long long chunks_small = (small + (unsigned)tiles_per_warp - 1) / (unsigned)tiles_per_warp;
        if (chunks_big + chunks_small > max_blocks) {
          // too long for one launch: full-length tiles only, the rest comes with the next trip
          if (nt > max_blocks * (unsigned)tiles_per_warp) nt = max_blocks * (unsigned)tiles_per_warp;
          big = nt; small = 0;
          chunks_big = (big + (unsigned)tiles_per_warp - 1) / (unsigned)tiles_per_warp; chunks_small = 0;
        }
        const unsigned long long chunks = chunks_big + chunks_small;
        if ((rc = lane_reserve_partials(&L, (size_t)chunks + 4096)) != SPD_OK) return rc;
        if ((rc = lane_reserve_aux(&L, (size_t)chunks)) != SPD_OK) return rc;
        LevelArgs la;
        la.colT_hot = p->d_colT_hot; la.lowR = p->d_lowR; la.dcold = p->d_dcold;
        memcpy(la.low, p->low, sizeof(la.low));
        la.xb_hot = p->d_xb_hot; la.xb_cold = p->d_xb_cold;
        memcpy(la.cold_start, p->cold_start, sizeof(la.cold_start));
        la.partials = L.d_partials; la.visited = L.d_aux; la.queue = L.d_queue;
        la.tile_first = tile; la.n_tiles = big;
        la.tile_first_small = (tile + big) << (c - c_small); la.n_tiles_small = small;
        la.n_chunks = (unsigned)chunks; la.n_chunks_big = (unsigned)chunks_big;
        la.n = n; la.NC = p->NC; la.NCP = p->NCP; la.HSP = p->HSP; la.c = c; la.c_small = c_small;
        la.tiles_per_warp = tiles_per_warp;
        const int HS = p->lvS0 + (p->lvB - 1) * p->lvS, LBv = p->lvB + (p->lvB & 1);
        const size_t smem = level_smem_bytes(n, p->lvB, HS, p->HSP, LBv, p->NC, p->NCP, SPB_REG_THREADS);
        if (smem > 220 * 1024) { set_error("level engine needs %zu B of shared memory", smem); return SPD_ELIMIT; }
        SPB_CUDA(cudaMemsetAsync(L.d_queue, 0, 2 * sizeof(unsigned int), L.stream));
        unsigned nb = 0;
        rc = level_launch(p->lvB, p->lvS0, p->lvS, p->skip, L.stream, &la, L.sm_count, smem, &nb);
        if (rc != SPD_OK) { if (rc == SPD_ELIMIT) set_error("no level kernel for B=%d S0=%d S=%d", p->lvB, p->lvS0, p->lvS); return rc; }
        SPB_CUDA(cudaGetLastError());
        ++launches;
        if ((rc = launch_reduce(L, L.d_partials, (size_t)chunks, L.d_result, 0, !first)) != SPD_OK) return rc;
        if ((rc = launch_reduce_u64(L, L.d_aux, (size_t)chunks, L.d_result, 1, !first)) != SPD_OK) return rc;
        launches += 2;
        first = false;
        have_visited = true;
        tile += nt; tiles_left -= nt;
        continue;
      }
      unsigned long long blocks = (tiles_left + tiles_per_block - 1) / tiles_per_block;
      if (blocks > max_blocks) blocks = max_blocks;
      unsigned long long nt = blocks * tiles_per_block;
      if (nt > tiles_left) nt = tiles_left;
      if ((rc = lane_reserve_partials(&L, (size_t)blocks + 4096)) != SPD_OK) return rc;
      if ((rc = lane_reserve_aux(&L, (size_t)blocks)) != SPD_OK) return rc;
      {
        SparseArgs a;
        a.mat_t = p->d_mat_t; a.xbase = p->d_xbase;
        a.partials = L.d_partials; a.visited = L.d_aux;
        a.tile_first = tile; a.n_tiles = nt; a.c = c; a.H = H; a.TC = TC; a.tiles_per_warp = tiles_per_warp;
        rc = sparse_launch(n, B, p->skip, L.stream, &a, (unsigned)blocks);
        if (rc != SPD_OK) { set_error("no sparse register kernel for n=%d B=%d", n, B); return rc; }
      }
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

// The plan as the C host prepared it (host/sp_level.c): rows ordered by level, the engine chosen and its images
// packed.  This function uploads them and loads the kernel; nothing is decided here.
int spd_sparse_plan_create_packed(int device, const double* mat_t, const double* xbase, const int* level_sorted,
                                  int nov, int skip, const spd_level_image* img, spd_sparse_plan** out) {
  if (!mat_t || !xbase || !level_sorted || !img || !out) { set_error("null argument"); return SPD_EINVAL; }
  if (nov < 2 || nov > 64) { set_error("sparse Ryser supports 2 <= n <= 64 (got %d)", nov); return SPD_ELIMIT; }
  spd_sparse_plan* p = new (std::nothrow) spd_sparse_plan();
  if (!p) return SPD_ENOMEM;
  int rc = lane_acquire(device, &p->lanep);
  if (rc != SPD_OK) { delete p; return rc; }
  Lane& L = *p->lanep;
  const int n = nov;
  p->n = nov;
  p->skip = skip ? 1 : 0;
  p->level.assign(level_sorted, level_sorted + n);
  p->lv_instr = img->instr_per_index;
  p->skip_long_tiles = img->skip_long_tiles;
  auto fail = [&](int code) { spd_sparse_plan_destroy(p); return code; };
  auto up = [&](const void* src, size_t bytes, void** dst) -> int {
    int r = lane_arena_alloc(&L, bytes ? bytes : 8, dst);
    if (r != SPD_OK) return r;
    if (bytes && cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, L.stream) != cudaSuccess) {
      set_error("sparse plan upload failed");
      return SPD_ECUDA;
    }
    return SPD_OK;
  };
  if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice(%d) failed", device); return fail(SPD_ECUDA); }
  if ((rc = up(mat_t, (size_t)n * n * sizeof(double), (void**)&p->d_mat_t)) != SPD_OK) return fail(rc);
  if ((rc = up(xbase, (size_t)n * sizeof(double), (void**)&p->d_xbase)) != SPD_OK) return fail(rc);
  if (img->B) {
    const int HS = img->S0 + (img->B - 1) * img->S, LBv = img->B + (img->B & 1);
    p->lvB = img->B; p->lvS0 = img->S0; p->lvS = img->S; p->lvR = img->R;
    p->NC = img->NC; p->NCP = img->NCP; p->HSP = img->HSP;
    if ((rc = up(img->colT_hot, (size_t)(n - 1) * img->HSP * 8, (void**)&p->d_colT_hot)) != SPD_OK) return fail(rc);
    if ((rc = up(img->lowR, (size_t)HS * LBv * 8, (void**)&p->d_lowR)) != SPD_OK) return fail(rc);
    if (HS * LBv > SPB_LV_MAXLOW) { set_error("low-column image too long"); return fail(SPD_ELIMIT); }
    memcpy(p->low, img->lowR, (size_t)HS * LBv * sizeof(double));
    if ((rc = up(img->dcold, (size_t)(n - 1) * img->NCP * 8, (void**)&p->d_dcold)) != SPD_OK) return fail(rc);
    if ((rc = up(img->xb_hot, (size_t)img->HSP * 8, (void**)&p->d_xb_hot)) != SPD_OK) return fail(rc);
    if ((rc = up(img->xb_cold, (size_t)img->NCP * 8, (void**)&p->d_xb_cold)) != SPD_OK) return fail(rc);
    if (n - img->B + 2 > SPB_LV_MAXSEG) { set_error("level table too long"); return fail(SPD_ELIMIT); }
    memcpy(p->cold_start, img->cold_start, (size_t)(n - img->B + 2) * sizeof(int));
  }
  // the sources are the caller's: wait until the copies have left them
  if (cudaStreamSynchronize(L.stream) != cudaSuccess) { set_error("sparse plan upload failed"); return fail(SPD_ECUDA); }
  if ((rc = smem_kernel_prepare(nov)) != SPD_OK) return fail(rc);
  if ((rc = lane_reserve_partials(&L, (1u << 20) + 8192)) != SPD_OK) return fail(rc);
  if ((rc = lane_reserve_aux(&L, 1u << 20)) != SPD_OK) return fail(rc);
  // load the kernel this plan will launch (CUDA loads kernels lazily) and set its shared-memory opt-in now,
  // so that the first run does not pay for either
  if (p->lvB) {
    LevelArgs la;
    memset(&la, 0, sizeof(la));           // partials == nullptr: prepare only
    la.n_chunks = 1;
    const int HS = p->lvS0 + (p->lvB - 1) * p->lvS, LBv = p->lvB + (p->lvB & 1);
    const size_t smem = level_smem_bytes(n, p->lvB, HS, p->HSP, LBv, p->NC, p->NCP, SPB_REG_THREADS);
    unsigned bps = 0;
    (void)level_launch(p->lvB, p->lvS0, p->lvS, p->skip, L.stream, &la, L.sm_count, smem, &bps);
  } else if (n >= SPB_SPARSE_NMIN && n <= SPB_SPARSE_NMAX) {
    SparseArgs sa;
    memset(&sa, 0, sizeof(sa));
    for (int b = 3; b <= 4; ++b) (void)sparse_launch(n, b, p->skip, L.stream, &sa, 0u);
  }
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
  // the host model's FP64 instructions per index of the engine that ran (0 for the shared-memory kernel):
  // what a roofline fraction of the sparse paths is quoted with
  if (p->info.path == SPD_PATH_SPARSE_REG || p->info.path == SPD_PATH_SKIPPER)
    p->info.aux1 = p->lv_instr;
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
