// SpaRyser / SkipPer on one device: plan (row-ordered matrix resident in HBM) and range planner.
//
// Reference paths replaced: gpu_perman64_xshared_coalescing_mshared_sparse + kernel
// (gpu_exact_sparse.cu:455-552, 853-914) and ..._mshared_skipper + kernel (:555-670, 1123-1190).
// The plan is given D, the matrix the reference's sparse kernels actually iterate over -- the CCS
// arrays scattered back to dense form (entries the CCS does not hold are 0) -- and the NW start
// vector; it orders the rows hot-first (see sparse_reg.cuh) once, at creation.
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
  double lv_cost = 1e300, hc_cost = 1e300;
  double lv_instr = 0.0;       // FP64 instructions per Gray index of the chosen LevelRyser configuration (model)
  double *d_colT_hot = nullptr, *d_lowR = nullptr, *d_dcold = nullptr, *d_xb_hot = nullptr, *d_xb_cold = nullptr;
  int* d_cold_start = nullptr;
  bool pending = false;
  spd_run_info info;
};

static int count_level_below(const spd_sparse_plan* p, int bound) {
  return (int)(std::lower_bound(p->level.begin(), p->level.end(), bound) - p->level.begin());
}


// FP64-instruction cost per Gray index of the two sparse engines (used to pick one per matrix)
static double hotcold_cost(const spd_sparse_plan* p, int B) {
  const int n = p->n;
  int H = count_level_below(p, B);
  H = ((H + 3) / 4) * 4;
  if (H > n) H = n;
  return 2.0 * H + (2.0 * (n - H) + 8.0) / (double)(1 << B);
}

// Packs the rows into the LevelRyser layout for (B, S0, S): S0 register slots for level 0, S for each other
// level < B, the other rows cold, sorted by level.  lvl[] / dmat_t / xbase are in the ORIGINAL row order.  Returns
// false when some level has more rows than the slots at or below it can take.
static bool level_pack(int n, int B, int S0, int S, int R, const std::vector<int>& lvl, const double* dmat_t,
                       const double* xbase, std::vector<double>& colT_hot, std::vector<double>& lowR,
                       std::vector<double>& dcold, std::vector<double>& xb_hot, std::vector<double>& xb_cold,
                       std::vector<int>& cold_start, int* NC_out, double* cost_out) {
  const int HS = S0 + (B - 1) * S, HT = HS + R, HSP = HT + (HT & 1), LB = B + (B & 1);
  auto base = [&](int L) { return L == 0 ? 0 : S0 + (L - 1) * S; };
  auto count = [&](int L) { return L == 0 ? S0 : S; };
  std::vector<int> slot_row(HT, -1);
  std::vector<int> order(n);
  for (int j = 0; j < n; ++j) order[j] = j;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return lvl[a] < lvl[b]; });
  // hot rows by ascending level: own level's slots first, then any free slot of a lower level
  std::vector<int> cold;
  for (int idx = 0; idx < n; ++idx) {
    const int j = order[idx];
    if (lvl[j] >= B) { cold.push_back(j); continue; }
    int placed = -1;
    for (int L = lvl[j]; L >= 0 && placed < 0; --L)
      for (int t = 0; t < count(L); ++t)
        if (slot_row[base(L) + t] < 0) { placed = base(L) + t; break; }
    if (placed < 0) return false;
    slot_row[placed] = j;
  }
  // the R cold rows of lowest level (refreshed most often) stay in registers too
  const int nrc = std::min(R, (int)cold.size());
  for (int t = 0; t < nrc; ++t) slot_row[HS + t] = cold[t];
  cold.erase(cold.begin(), cold.begin() + nrc);
  const int NC = (int)cold.size(), NCP = NC + (NC & 1) + ((NC == 0) ? 2 : 0);
  colT_hot.assign((size_t)(n - 1) * HSP, 0.0);
  lowR.assign((size_t)HS * LB, 0.0);
  dcold.assign((size_t)(n - 1) * NCP, 0.0);
  xb_hot.assign(HSP, 1.0);
  xb_cold.assign(NCP, 1.0);
  for (int sl = 0; sl < HT; ++sl) {
    const int j = slot_row[sl];
    if (j < 0) continue;                          // neutral slot: x = 1, all entries 0
    xb_hot[sl] = xbase[j];
    for (int k = 0; k < n - 1; ++k) colT_hot[(size_t)k * HSP + sl] = dmat_t[(size_t)k * n + j];
    if (sl < HS)
      for (int q = 0; q < B; ++q) lowR[(size_t)sl * LB + q] = dmat_t[(size_t)q * n + j];
  }
  cold_start.assign(n - B + 2, NC);
  for (int jc = NC - 1; jc >= 0; --jc) {
    const int j = cold[jc];
    xb_cold[jc] = xbase[j];
    for (int k = 0; k < n - 1; ++k) dcold[(size_t)k * NCP + jc] = dmat_t[(size_t)k * n + j];
  }
  // cold_start[i] = first cold row with level >= B+i  (levels run up to n = "never touched")
  {
    int jc = 0;
    for (int i = 0; i <= n - B; ++i) {
      while (jc < NC && lvl[cold[jc]] < B + i) ++jc;
      cold_start[i] = jc;
    }
    cold_start[n - B + 1] = NC;
  }
  // cost per index: hot slots + recombination + expected cold refresh (x3: it runs from shared memory)
  double hot = 0.0;
  for (int L = 0; L < B; ++L) hot += (double)count(L) * 2.0 * (double)(1 << (B - L));
  double coldc = 0.0, w = 0.5;
  for (int z = 0; z < 16 && B + z <= n; ++z, w *= 0.5) coldc += w * 3.0 * (double)cold_start[(z + 1 <= n - B + 1) ? z + 1 : n - B + 1];
  const double per_block = hot + 2.0 * R + (double)((1 << B) + B) + coldc;
  *cost_out = per_block / (double)(1 << B);
  *NC_out = NC;
  return true;
}

// ---- choice of the low columns ------------------------------------------------------------------------
// A plan that will be run over the WHOLE index space (flag SPD_SPARSE_REORDER: the id entry points, whatever
// the split over chunks and devices) may walk the columns in any order: the Ryser sum runs over every subset
// of the columns 0 .. n-2.  The cost of the level engine is set by how many rows have their first non-zero
// in each of the B most frequently flipped columns (level populations, 2^(B-L) values per row of level L) and
// by how well they fit the slot configurations that exist (S0 slots for level 0, S for the others).  SortOrder
// puts the sparsest columns first but knows nothing about either: e.g. level populations (4, 5, 3, 3) need
// S = 6, while the same matrix with columns 1 and 2 exchanged has (4, 4, 4, 3) and fits S = 4 (15 instead of
// 18.5 FP64 instructions per index).  So: among the ordered B-tuples of the 9 sparsest columns take the one
// with the cheapest fitting slot configuration; the other columns keep their order.
static const int kSlotOpts[6] = {1, 2, 3, 4, 6, 8};

// hot-slot FP64 instructions per index of the cheapest (S0, S) that takes level populations pops[0..B), or < 0
static double fit_slots(const int* pops, int B) {
  for (int si = 0; si < 6; ++si) {
    const int S = kSlotOpts[si];
    for (int S0 = std::max(1, S - 2); S0 <= S; ++S0) {
      int free_[4] = {S0, S, S, S};
      bool ok = true;
      for (int L = 0; L < B && ok; ++L) {
        int need = pops[L];
        for (int LL = L; LL >= 0 && need > 0; --LL) { const int t = std::min(need, free_[LL]); free_[LL] -= t; need -= t; }
        ok = need == 0;
      }
      if (ok) {
        double c = 0.0;
        for (int L = 0; L < B; ++L) c += (double)(L == 0 ? S0 : S) * 2.0 * (double)(1 << (B - L));
        return c / (double)(1 << B);
      }
    }
  }
  return -1.0;
}

// perm[k'] = original column at position k' (a permutation of 0 .. n-2; column n-1 stays last)
static void choose_low_columns(int n, const double* dmat_t, std::vector<int>& perm) {
  perm.resize(n);
  for (int k = 0; k < n; ++k) perm[k] = k;
  if (n < 10) return;
  std::vector<std::pair<int, int>> cnt;
  for (int k = 0; k < n - 1; ++k) {
    int c = 0;
    for (int j = 0; j < n; ++j) c += dmat_t[(size_t)k * n + j] != 0.0;
    cnt.push_back({c, k});
  }
  std::stable_sort(cnt.begin(), cnt.end());
  const int K = std::min(9, n - 1);
  unsigned long long rows_of[9];
  int cand[9];
  for (int i = 0; i < K; ++i) {
    cand[i] = cnt[i].second;
    rows_of[i] = 0ull;
    for (int j = 0; j < n; ++j)
      if (dmat_t[(size_t)cand[i] * n + j] != 0.0) rows_of[i] |= 1ull << j;
  }
  // the caller's order as the incumbent (ties keep it)
  double best = 1e300;
  int best_t[4] = {0, 1, 2, 3}, best_B = 0;
  auto eval = [&](const int* t, int B, bool incumbent) {
    unsigned long long seen = 0ull;
    int pops[4];
    for (int L = 0; L < B; ++L) {
      unsigned long long r;
      if (incumbent) { r = 0ull; for (int j = 0; j < n; ++j) if (dmat_t[(size_t)t[L] * n + j] != 0.0) r |= 1ull << j; }
      else r = rows_of[t[L]];
      pops[L] = __builtin_popcountll(r & ~seen);
      seen |= r;
    }
    return fit_slots(pops, B);
  };
  for (int B = 3; B <= 4; ++B) {
    int t0[4] = {0, 1, 2, 3};
    const double c0 = eval(t0, B, true);
    if (c0 >= 0 && c0 < best) { best = c0; best_B = 0; }
  }
  for (int B = 3; B <= 4; ++B) {
    int t[4];
    for (t[0] = 0; t[0] < K; ++t[0])
      for (t[1] = 0; t[1] < K; ++t[1]) {
        if (t[1] == t[0]) continue;
        for (t[2] = 0; t[2] < K; ++t[2]) {
          if (t[2] == t[0] || t[2] == t[1]) continue;
          for (t[3] = 0; t[3] < (B == 4 ? K : 1); ++t[3]) {
            if (B == 4 && (t[3] == t[0] || t[3] == t[1] || t[3] == t[2])) continue;
            const double c = eval(t, B, false);
            if (c >= 0 && c < best - 0.26) {       // at least a quarter instruction per index better than the incumbent
              best = c; best_B = B;
              for (int L = 0; L < B; ++L) best_t[L] = cand[t[L]];
            }
          }
        }
      }
  }
  if (best_B == 0) return;
  std::vector<char> used(n, 0);
  int pos = 0;
  for (int L = 0; L < best_B; ++L) { perm[pos++] = best_t[L]; used[best_t[L]] = 1; }
  for (int k = 0; k < n - 1; ++k)
    if (!used[k]) perm[pos++] = k;
  perm[n - 1] = n - 1;
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
    const int tiles_log2 = env_int("SP_SPARSE_TILES_LOG2", p->skip ? 21 : 20);
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
        // leaves one partial sum (and one visited count) per chunk; at most 2^20 chunks per launch
        unsigned long long chunks = (tiles_left + (unsigned)tiles_per_warp - 1) / (unsigned)tiles_per_warp;
        if (chunks > max_blocks) chunks = max_blocks;
        unsigned long long nt = chunks * (unsigned)tiles_per_warp;
        if (nt > tiles_left) nt = tiles_left;
        if ((rc = lane_reserve_partials(&L, (size_t)chunks + 4096)) != SPD_OK) return rc;
        if ((rc = lane_reserve_aux(&L, (size_t)chunks)) != SPD_OK) return rc;
        LevelArgs la;
        la.colT_hot = p->d_colT_hot; la.lowR = p->d_lowR; la.dcold = p->d_dcold;
        la.xb_hot = p->d_xb_hot; la.xb_cold = p->d_xb_cold; la.cold_start = p->d_cold_start;
        la.partials = L.d_partials; la.visited = L.d_aux; la.queue = L.d_queue;
        la.tile_first = tile; la.n_tiles = nt; la.n_chunks = (unsigned)chunks;
        la.n = n; la.NC = p->NC; la.NCP = p->NCP; la.HSP = p->HSP; la.c = c; la.tiles_per_warp = tiles_per_warp;
        const int HS = p->lvS0 + (p->lvB - 1) * p->lvS, LBv = p->lvB + (p->lvB & 1);
        const size_t smem = level_smem_bytes(n, p->lvB, HS, p->HSP, LBv, p->NC, p->NCP, c, SPB_REG_THREADS);
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

int spd_sparse_plan_create(int device, const double* dmat_t, const double* xbase, int nov, int skip,
                           spd_sparse_plan** out) {
  return spd_sparse_plan_create_ex(device, dmat_t, xbase, nov, skip, 0, out);
}

int spd_sparse_plan_create_ex(int device, const double* dmat_in, const double* xbase, int nov, int skip, int flags,
                              spd_sparse_plan** out) {
  if (!dmat_in || !xbase || !out) { set_error("null argument"); return SPD_EINVAL; }
  if (nov < 2 || nov > 64) { set_error("sparse Ryser supports 2 <= n <= 64 (got %d)", nov); return SPD_ELIMIT; }
  // column order (see choose_low_columns): only for plans that will cover the whole index space
  std::vector<double> dperm;
  const double* dmat_t = dmat_in;
  if ((flags & SPD_SPARSE_REORDER) && env_int("SP_SPARSE_REORDER", 1) != 0) {
    std::vector<int> cperm;
    choose_low_columns(nov, dmat_in, cperm);
    bool moved = false;
    for (int k = 0; k < nov; ++k) moved = moved || cperm[k] != k;
    if (moved) {
      dperm.resize((size_t)nov * nov);
      for (int k = 0; k < nov; ++k)
        memcpy(&dperm[(size_t)k * nov], &dmat_in[(size_t)cperm[k] * nov], (size_t)nov * sizeof(double));
      dmat_t = dperm.data();
    }
  }
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

  // ---- choose the engine: LevelRyser (B, S) with the lowest cost, if the matrix fits its slots ----
  p->hc_cost = std::min(hotcold_cost(p, 3), hotcold_cost(p, 4)) * 1.45;   // measured: ~68 % of the pipe
  const int engine = env_int("SP_SPARSE_ENGINE", 0);                       // 0 auto, 1 hot/cold, 2 level
  if (engine != 1 && n >= 6) {
    static const int s_opts[6] = {1, 2, 3, 4, 6, 8};
    std::vector<double> bh, bl, bd, bxh, bxc;
    std::vector<int> bcs;
    const int forceB = env_int("SP_SPARSE_LOWCOLS", 0), forceS = env_int("SP_LEVEL_SLOTS", 0), forceS0 = env_int("SP_LEVEL_SLOTS0", 0);
    for (int B = 3; B <= 4; ++B) {
      if (B + 2 > n - 1) continue;
      if (forceB && forceB != B) continue;
      for (int si = 0; si < 6; ++si) {
        const int S = s_opts[si];
        if (forceS && forceS != S) continue;
        bool fits = false;
        // level 0 (2^B values per block, the most expensive level) may have up to two slots fewer
        for (int S0 = std::max(1, S - 2); S0 <= S; ++S0) {
          if (forceS0 && forceS0 != S0) continue;
          const int R = level_regcold(B, S0, S, p->skip != 0);
          std::vector<double> h, l, d, xh, xc;
          std::vector<int> cs;
          int NC = 0;
          double cost = 0;
          if (!level_pack(n, B, S0, S, R, lvl, dmat_t, xbase, h, l, d, xh, xc, cs, &NC, &cost)) continue;
          fits = true;
          const double raw = cost;
          cost *= 1.15;
          if (level_minblocks(B, S0, S, p->skip != 0) < 4) cost *= 1.1;      // 3 instead of 4 blocks per SM
          if (cost < p->lv_cost) {
            p->lv_cost = cost; p->lv_instr = raw; p->lvB = B; p->lvS0 = S0; p->lvS = S; p->lvR = R; p->NC = NC;
            bh.swap(h); bl.swap(l); bd.swap(d); bxh.swap(xh); bxc.swap(xc); bcs.swap(cs);
          }
          break;   // more level-0 slots for the same S only cost more
        }
        if (fits) break;   // a larger S for the same B only costs more
      }
    }
    if (p->lvB && (engine == 2 || p->lv_cost < p->hc_cost || n > SPB_SPARSE_NMAX)) {
      const int HT = p->lvS0 + (p->lvB - 1) * p->lvS + p->lvR;
      p->HSP = HT + (HT & 1);
      p->NCP = (int)bxc.size();
      auto up = [&](const void* src, size_t bytes, void** dst) -> int {
        int r = lane_arena_alloc(&L, bytes ? bytes : 8, dst);
        if (r != SPD_OK) return r;
        if (bytes && cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, L.stream) != cudaSuccess) {
          set_error("level image upload failed");
          return SPD_ECUDA;
        }
        return SPD_OK;
      };
      if ((rc = up(bh.data(), bh.size() * 8, (void**)&p->d_colT_hot)) != SPD_OK) return fail(rc);
      if ((rc = up(bl.data(), bl.size() * 8, (void**)&p->d_lowR)) != SPD_OK) return fail(rc);
      if ((rc = up(bd.data(), bd.size() * 8, (void**)&p->d_dcold)) != SPD_OK) return fail(rc);
      if ((rc = up(bxh.data(), bxh.size() * 8, (void**)&p->d_xb_hot)) != SPD_OK) return fail(rc);
      if ((rc = up(bxc.data(), bxc.size() * 8, (void**)&p->d_xb_cold)) != SPD_OK) return fail(rc);
      if ((rc = up(bcs.data(), bcs.size() * 4, (void**)&p->d_cold_start)) != SPD_OK) return fail(rc);
      if (cudaStreamSynchronize(L.stream) != cudaSuccess) { set_error("level image upload failed"); return fail(SPD_ECUDA); }
    } else {
      p->lvB = 0;
    }
  }
  // load the kernel this plan will launch (CUDA loads kernels lazily) and set its shared-memory opt-in now,
  // so that the first run does not pay for either
  if (p->lvB) {
    LevelArgs la;
    memset(&la, 0, sizeof(la));           // partials == nullptr: prepare only
    la.n_chunks = 1;
    const int HS = p->lvS0 + (p->lvB - 1) * p->lvS, LBv = p->lvB + (p->lvB & 1);
    int cc = n - 1 < 12 ? n - 1 : 12;
    if (cc < p->lvB + 1) cc = p->lvB + 1;
    const size_t smem = level_smem_bytes(n, p->lvB, HS, p->HSP, LBv, p->NC, p->NCP, cc, SPB_REG_THREADS);
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
    p->info.aux1 = p->lvB ? p->lv_instr : std::min(hotcold_cost(p, 3), hotcold_cost(p, 4));
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
