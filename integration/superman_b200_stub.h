// superman_b200_stub.h -- what a maintainer of the reference puts in place of
//     #include "gpu_exact_dense.cu"
//     #include "gpu_exact_sparse.cu"
//     #include "gpu_approximation_dense.cu"
//     #include "gpu_approximation_sparse.cu"
// at the top of main.cu (main.cu:12-15) to run RunAlgo / RunPermanForGridGraphs on libsuperman_b200.so:
// every gpu_perman64_* wrapper those two functions name (main.cu:36-177, 266-284), with the reference's
// own signatures, forwarding to the C-ABI of include/superman_b200.h.  Nothing else in main.cu, util.h
// or algo.h changes.  `make -C oracle integration` builds the reference's main.cu this way
// (oracle/_ref/perman_ref_stub) and tests/test_gpu_integration.py diffs its Result lines with `perman`.
//
// Matrix element types are widened to double at the call (int and float entries are exact in double).
// The launch geometry arguments (grid_dim, block_dim) are accepted and ignored: the library sizes its
// own persistent grids.  The "kernel in ..." lines the reference wrappers print are reproduced from the
// library's CUDA-event times.
#pragma once
extern "C" {
#include "superman_b200.h"
}
#include <cstdlib>
#include <iostream>
#include <vector>

namespace spb_stub {

template <class T>
inline std::vector<double> widen(const T* p, size_t n) { return std::vector<double>(p, p + n); }

inline void kernel_lines(const sp_stats& st) {
  if (st.devices <= 1) { std::cout << "kernel in " << st.kernel_ms * 1e-3 << std::endl; return; }
  for (int g = 0; g < st.devices; ++g) std::cout << "kernel" << g << " in " << st.device_ms[g] * 1e-3 << std::endl;
}

inline double checked(double v, const sp_stats& st) {
  if (st.error != SP_OK) { std::cerr << "superman_b200: " << sp_last_error() << std::endl; std::exit(1); }
  kernel_lines(st);
  return v;
}

template <class T>
inline double dense(T* mat, int nov, int id, int gpu_num, bool cpu, int threads) {
  sp_stats st;
  auto m = widen(mat, (size_t)nov * nov);
  return checked(sp_dense_ryser(m.data(), nov, id, gpu_num, cpu, threads, &st), st);
}
template <class T>
inline double sparse(T* mat, int* cptrs, int* rows, T* cvals, int nov, int id, int gpu_num, bool cpu, int threads) {
  sp_stats st;
  auto m = widen(mat, (size_t)nov * nov);
  auto v = widen(cvals, (size_t)cptrs[nov]);
  return checked(sp_sparse_ryser(m.data(), cptrs, rows, v.data(), nov, id, gpu_num, cpu, threads, &st), st);
}
template <class T>
inline double skipper(T* mat, int* rptrs, int* cols, int* cptrs, int* rows, T* cvals, int nov, int id, int gpu_num,
                      bool cpu, int threads) {
  sp_stats st;
  auto m = widen(mat, (size_t)nov * nov);
  auto v = widen(cvals, (size_t)cptrs[nov]);
  return checked(sp_skipper(m.data(), rptrs, cols, cptrs, rows, v.data(), nov, id, gpu_num, cpu, threads, &st), st);
}
// the reference's single-GPU sparse Rasmussen wrapper takes the CRS only (main.cu:159,266): the CCS the
// library also wants is the transposed pattern
inline void ccs_from_crs(const int* rptrs, const int* cols, int nov, std::vector<int>& cptrs, std::vector<int>& rows) {
  const int nnz = rptrs[nov];
  cptrs.assign(nov + 1, 0);
  rows.assign(nnz > 0 ? nnz : 1, 0);
  for (int t = 0; t < nnz; ++t) cptrs[cols[t] + 1]++;
  for (int j = 0; j < nov; ++j) cptrs[j + 1] += cptrs[j];
  std::vector<int> fill(cptrs.begin(), cptrs.end() - 1);
  for (int i = 0; i < nov; ++i)
    for (int t = rptrs[i]; t < rptrs[i + 1]; ++t) rows[fill[cols[t]]++] = i;
}

}  // namespace spb_stub

// ---- gpu_exact_dense.cu: ids 0-6 and 66 (main.cu:34-73) ---------------------------------------------
template <class T> double gpu_perman64_xglobal(T* mat, int nov, int, int) { return spb_stub::dense(mat, nov, 0, 1, false, 0); }
template <class T> double gpu_perman64_xlocal(T* mat, int nov, int, int) { return spb_stub::dense(mat, nov, 1, 1, false, 0); }
template <class T> double gpu_perman64_xshared(T* mat, int nov, int, int) { return spb_stub::dense(mat, nov, 2, 1, false, 0); }
template <class T> double gpu_perman64_xshared_coalescing(T* mat, int nov, int, int) { return spb_stub::dense(mat, nov, 3, 1, false, 0); }
template <class T> double gpu_perman64_xshared_coalescing_mshared(T* mat, int nov, int, int) { return spb_stub::dense(mat, nov, 4, 1, false, 0); }
template <class T> double gpu_perman64_xshared_coalescing_mshared_multigpu(T* mat, int nov, int gpu_num, int, int) {
  return spb_stub::dense(mat, nov, 5, gpu_num, false, 0);
}
template <class T> double gpu_perman64_xshared_coalescing_mshared_multigpucpu_chunks(T* mat, int nov, int gpu_num, bool cpu, int threads, int, int) {
  return spb_stub::dense(mat, nov, 6, gpu_num, cpu, threads);
}
template <class T> double gpu_perman64_xshared_coalescing_mshared_multigpu_manual_distribution(T* mat, int nov, int gpu_num, int, int) {
  const int have = sp_device_count();
  return spb_stub::dense(mat, nov, 5, gpu_num < have ? gpu_num : (have > 0 ? have : 1), false, 0);
}

// ---- gpu_exact_sparse.cu: ids 1-8 and 66 (main.cu:108-152) -------------------------------------------
template <class T> double gpu_perman64_xlocal_sparse(T* mat, int* cptrs, int* rows, T* cvals, int nov, int, int) { return spb_stub::sparse(mat, cptrs, rows, cvals, nov, 1, 1, false, 0); }
template <class T> double gpu_perman64_xshared_sparse(T* mat, int* cptrs, int* rows, T* cvals, int nov, int, int) { return spb_stub::sparse(mat, cptrs, rows, cvals, nov, 2, 1, false, 0); }
template <class T> double gpu_perman64_xshared_coalescing_sparse(T* mat, int* cptrs, int* rows, T* cvals, int nov, int, int) { return spb_stub::sparse(mat, cptrs, rows, cvals, nov, 3, 1, false, 0); }
template <class T> double gpu_perman64_xshared_coalescing_mshared_sparse(T* mat, int* cptrs, int* rows, T* cvals, int nov, int, int) { return spb_stub::sparse(mat, cptrs, rows, cvals, nov, 4, 1, false, 0); }
template <class T> double gpu_perman64_xshared_coalescing_mshared_multigpu_sparse(T* mat, int* cptrs, int* rows, T* cvals, int nov, int gpu_num, int, int) {
  return spb_stub::sparse(mat, cptrs, rows, cvals, nov, 5, gpu_num, false, 0);
}
template <class T> double gpu_perman64_xshared_coalescing_mshared_multigpucpu_chunks_sparse(T* mat, int* cptrs, int* rows, T* cvals, int nov, int gpu_num, bool cpu, int threads, int, int) {
  return spb_stub::sparse(mat, cptrs, rows, cvals, nov, 6, gpu_num, cpu, threads);
}
template <class T> double gpu_perman64_xshared_coalescing_mshared_skipper(T* mat, int* rptrs, int* cols, int* cptrs, int* rows, T* cvals, int nov, int, int) {
  return spb_stub::skipper(mat, rptrs, cols, cptrs, rows, cvals, nov, 7, 1, false, 0);
}
template <class T> double gpu_perman64_xshared_coalescing_mshared_multigpucpu_chunks_skipper(T* mat, int* rptrs, int* cols, int* cptrs, int* rows, T* cvals, int nov, int gpu_num, bool cpu, int threads, int, int) {
  return spb_stub::skipper(mat, rptrs, cols, cptrs, rows, cvals, nov, 8, gpu_num, cpu, threads);
}
template <class T> double gpu_perman64_xshared_coalescing_mshared_multigpu_sparse_manual_distribution(T* mat, int* cptrs, int* rows, T* cvals, int nov, int gpu_num, int, int) {
  const int have = sp_device_count();
  return spb_stub::sparse(mat, cptrs, rows, cvals, nov, 5, gpu_num < have ? gpu_num : (have > 0 ? have : 1), false, 0);
}

// ---- gpu_approximation_dense.cu (main.cu:78-100) ---------------------------------------------------------
template <class T> double gpu_perman64_rasmussen(T* mat, int nov, int number_of_times) {
  sp_stats st;
  auto m = spb_stub::widen(mat, (size_t)nov * nov);
  return spb_stub::checked(sp_rasmussen_dense(m.data(), nov, number_of_times, 1, 0, &st), st);
}
template <class T> double gpu_perman64_approximation(T* mat, int nov, int number_of_times, int scale_intervals, int scale_times) {
  sp_stats st;
  auto m = spb_stub::widen(mat, (size_t)nov * nov);
  return spb_stub::checked(sp_scaling_dense(m.data(), nov, number_of_times, scale_intervals, scale_times, 1, 0, &st), st);
}
template <class T> double gpu_perman64_rasmussen_multigpucpu_chunks(T* mat, int nov, int number_of_times, int gpu_num, bool, int) {
  sp_stats st;
  auto m = spb_stub::widen(mat, (size_t)nov * nov);
  return spb_stub::checked(sp_rasmussen_dense(m.data(), nov, number_of_times, gpu_num, 0, &st), st);
}
template <class T> double gpu_perman64_approximation_multigpucpu_chunks(T* mat, int nov, int number_of_times, int gpu_num, bool, int scale_intervals, int scale_times, int) {
  sp_stats st;
  auto m = spb_stub::widen(mat, (size_t)nov * nov);
  return spb_stub::checked(sp_scaling_dense(m.data(), nov, number_of_times, scale_intervals, scale_times, gpu_num, 0, &st), st);
}

// ---- gpu_approximation_sparse.cu (main.cu:157-179, 264-286) -------------------------------------------
inline double gpu_perman64_rasmussen_sparse(int* rptrs, int* cols, int nov, int nnz, int number_of_times, bool) {
  sp_stats st;
  std::vector<int> cptrs, rows;
  spb_stub::ccs_from_crs(rptrs, cols, nov, cptrs, rows);
  return spb_stub::checked(sp_rasmussen_sparse(rptrs, cols, cptrs.data(), rows.data(), nov, nnz, number_of_times, 1, 0, &st), st);
}
inline double gpu_perman64_approximation_sparse(int* cptrs, int* rows, int* rptrs, int* cols, int nov, int nnz, int number_of_times,
                                                int scale_intervals, int scale_times, bool) {
  sp_stats st;
  return spb_stub::checked(sp_scaling_sparse(cptrs, rows, rptrs, cols, nov, nnz, number_of_times, scale_intervals, scale_times, 1, 0, &st), st);
}
inline double gpu_perman64_rasmussen_multigpucpu_chunks_sparse(int* cptrs, int* rows, int* rptrs, int* cols, int nov, int nnz,
                                                               int number_of_times, int gpu_num, bool, int, bool) {
  sp_stats st;
  return spb_stub::checked(sp_rasmussen_sparse(rptrs, cols, cptrs, rows, nov, nnz, number_of_times, gpu_num, 0, &st), st);
}
inline double gpu_perman64_approximation_multigpucpu_chunks_sparse(int* cptrs, int* rows, int* rptrs, int* cols, int nov, int nnz,
                                                                   int number_of_times, int gpu_num, bool, int scale_intervals,
                                                                   int scale_times, int, bool) {
  sp_stats st;
  return spb_stub::checked(sp_scaling_sparse(cptrs, rows, rptrs, cols, nov, nnz, number_of_times, scale_intervals, scale_times, gpu_num, 0, &st), st);
}
