"""Host-side mirror of the reference's wrapper interface (same names, argument meaning and
return value as the templated gpu_perman64_* functions in gpu_exact_dense.cu etc.), implemented
as thin calls into the C-ABI of libsuperman_b200.so.  No arithmetic happens in Python."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _ffi
from ._ffi import SpStats, lib

__all__ = [
    "SupermanError", "device_count", "fp64_peak", "nw_factor", "Matrix",
    "SpStats", "dense_ryser", "dense_ryser_range", "DenseHandle", "permanent_compressed",
    "sparse_ryser", "skipper", "sparse_ryser_range",
    "rasmussen_sparse", "scaling_sparse", "rasmussen_dense", "scaling_dense", "approx_trials_sparse", "approx_trials_dense", "approx_trace_sparse", "approx_trace_dense", "int_peak", "set_precision",
    "gpu_perman64_rasmussen_sparse", "gpu_perman64_rasmussen_multigpucpu_chunks_sparse",
    "gpu_perman64_approximation_sparse", "gpu_perman64_approximation_multigpucpu_chunks_sparse",
    "gpu_perman64_rasmussen", "gpu_perman64_rasmussen_multigpucpu_chunks",
    "gpu_perman64_approximation", "gpu_perman64_approximation_multigpucpu_chunks",
    "gpu_perman64_xlocal_sparse", "gpu_perman64_xshared_sparse", "gpu_perman64_xshared_coalescing_sparse",
    "gpu_perman64_xshared_coalescing_mshared_sparse", "gpu_perman64_xshared_coalescing_mshared_multigpu_sparse",
    "gpu_perman64_xshared_coalescing_mshared_multigpucpu_chunks_sparse",
    "gpu_perman64_xshared_coalescing_mshared_skipper",
    "gpu_perman64_xshared_coalescing_mshared_multigpucpu_chunks_skipper",
    "gpu_perman64_xglobal", "gpu_perman64_xlocal", "gpu_perman64_xshared",
    "gpu_perman64_xshared_coalescing", "gpu_perman64_xshared_coalescing_mshared",
    "gpu_perman64_xshared_coalescing_mshared_multigpu",
    "gpu_perman64_xshared_coalescing_mshared_multigpucpu_chunks",
]


class SupermanError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"superman_b200 error {code}: {msg}")
        self.code = code


def _dmat(mat, nov=None) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(mat, dtype=np.float64))
    if nov is None:
        nov = a.shape[0]
    a = a.reshape(nov * nov)
    return a


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _check(value: float, st: SpStats) -> float:
    if st.error != 0 or (isinstance(value, float) and math.isnan(value) and st.error != 0):
        raise SupermanError(st.error, _ffi.last_error())
    return value


class Matrix:
    """Owner of an sp_matrix: dense row-major `mat` plus CRS/CCS once `compress()` ran.
    Mirrors what main.cu holds between ReadMatrix and RunAlgo (main.cu:500-528)."""

    def __init__(self):
        self._m = _ffi.SpMatrix()
        self._live = False

    # -- constructors -------------------------------------------------------------------------
    @classmethod
    def read(cls, path: str, binary: bool = False) -> "Matrix":
        self = cls()
        rc = lib.sp_matrix_read(str(path).encode(), int(binary), C.byref(self._m))
        if rc != 0:
            raise SupermanError(rc, _ffi.last_error())
        self._live = True
        return self

    @classmethod
    def from_dense(cls, mat, nov=None) -> "Matrix":
        a = _dmat(mat, nov)
        n = int(round(math.sqrt(a.size)))
        self = cls()
        rc = lib.sp_matrix_from_dense(_ptr(a), n, C.byref(self._m))
        if rc != 0:
            raise SupermanError(rc, _ffi.last_error())
        self._live = True
        return self

    @classmethod
    def grid(cls, m: int, n: int) -> "Matrix":
        self = cls()
        rc = lib.sp_matrix_grid(m, n, C.byref(self._m))
        if rc != 0:
            raise SupermanError(rc, _ffi.last_error())
        self._live = True
        return self

    def compress(self, preprocessing: int = 0) -> "Matrix":
        rc = lib.sp_matrix_compress(C.byref(self._m), preprocessing)
        if rc != 0:
            raise SupermanError(rc, _ffi.last_error())
        return self

    def reduce(self) -> float:
        """degree-0/1/2 compression in place; returns the factor with perm(original) = factor * perm(self)"""
        f = C.c_double(1.0)
        rc = lib.sp_matrix_reduce(C.byref(self._m), C.byref(f))
        if rc < 0:
            raise SupermanError(rc, _ffi.last_error())
        return f.value

    def reduce_step(self, factor: float = 1.0):
        """one d1compress-else-d2compress step; returns (kind 0/1/2, factor)"""
        f = C.c_double(factor)
        rc = lib.sp_matrix_reduce_step(C.byref(self._m), C.byref(f))
        if rc < 0:
            raise SupermanError(rc, _ffi.last_error())
        return rc, f.value

    def min_degree(self) -> int:
        rc = lib.sp_matrix_min_degree(C.byref(self._m))
        if rc < 0:
            raise SupermanError(rc, _ffi.last_error())
        return rc

    def split34(self, min_deg: int):
        """d34compress: self becomes the first matrix, returns the second (or None when nothing applies)"""
        other = Matrix()
        rc = lib.sp_matrix_split34(C.byref(self._m), min_deg, C.byref(other._m))
        if rc < 0:
            raise SupermanError(rc, _ffi.last_error())
        if rc == 0:
            return None
        other._live = True
        return other

    def scale(self, threshold: float, converge: bool = False):
        """scalesk + scaleMatrix (converge=True: Sinkhorn to convergence, sp_matrix_balance); returns (rv, cv, sweeps)"""
        n = self._m.nov
        rv = np.zeros(n, dtype=np.float64)
        cv = np.zeros(n, dtype=np.float64)
        rc = (lib.sp_matrix_balance if converge else lib.sp_matrix_scale)(C.byref(self._m), float(threshold), rv.ctypes.data_as(_ffi._dp),
                                 cv.ctypes.data_as(_ffi._dp))
        if rc < 0:
            raise SupermanError(rc, _ffi.last_error())
        return rv, cv, rc

    def dm(self):
        """Dulmage-Mendelsohn: erase entries on no perfect matching; returns (erased, matching size)"""
        k = C.c_int(0)
        rc = lib.sp_matrix_dm(C.byref(self._m), C.byref(k))
        if rc < 0:
            raise SupermanError(rc, _ffi.last_error())
        return rc, k.value

    # -- views (copies) ------------------------------------------------------------------------
    nov = property(lambda self: self._m.nov)
    nnz = property(lambda self: self._m.nnz)
    header_nnz = property(lambda self: self._m.header_nnz)
    type = property(lambda self: ("int", "float", "double")[self._m.type])

    def _arr(self, ptr, count, dtype):
        if not ptr:
            return None
        return np.ctypeslib.as_array(ptr, shape=(count,)).astype(dtype, copy=True)

    @property
    def mat(self):
        n = self._m.nov
        return self._arr(self._m.mat, n * n, np.float64).reshape(n, n)

    cptrs = property(lambda self: self._arr(self._m.cptrs, self._m.nov + 1, np.int32))
    rptrs = property(lambda self: self._arr(self._m.rptrs, self._m.nov + 1, np.int32))
    rows = property(lambda self: self._arr(self._m.rows, self._m.nnz, np.int32))
    cols = property(lambda self: self._arr(self._m.cols, self._m.nnz, np.int32))
    cvals = property(lambda self: self._arr(self._m.cvals, self._m.nnz, np.float64))
    rvals = property(lambda self: self._arr(self._m.rvals, self._m.nnz, np.float64))

    def free(self):
        if self._live:
            lib.sp_matrix_free(C.byref(self._m))
            self._live = False

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def device_count() -> int:
    return lib.sp_device_count()


def nw_factor(nov: int) -> float:
    return lib.sp_nw_factor(nov)


def fp64_peak(device: int = 0, millis: int = 200) -> float:
    r = lib.sp_fp64_peak(device, millis)
    if r < 0:
        raise SupermanError(-1, _ffi.last_error())
    return r


def set_precision(quad: bool) -> None:
    """dense exact paths in double-double arithmetic (the revised front-end's -q) when `quad`, else FP64"""
    lib.sp_set_precision(1 if quad else 0)


def int_peak(device: int = 0, millis: int = 200) -> float:
    r = lib.sp_int_peak(device, millis)
    if r < 0:
        raise SupermanError(-1, _ffi.last_error())
    return r


def dense_ryser(mat, nov=None, algo_id=4, gpu_num=1, cpu=False, threads=16, stats: SpStats | None = None) -> float:
    a = _dmat(mat, nov)
    nov = int(round(math.sqrt(a.size)))
    st = stats if stats is not None else SpStats()
    v = lib.sp_dense_ryser(_ptr(a), nov, algo_id, gpu_num, int(cpu), threads, C.byref(st))
    return _check(v, st)


def permanent_compressed(mat, nov=None, sparse=False, preprocessing=0, algo_id=4, gpu_num=1, threads=16,
                         scaling_threshold=0.0, leaf_nov=0, stats: SpStats | None = None) -> float:
    """the revised front-end's -o path: degree compression, d34 splits, optional -u scaling (sp_permanent_compressed)"""
    a = _dmat(mat, nov)
    nov = int(round(math.sqrt(a.size)))
    st = stats if stats is not None else SpStats()
    v = lib.sp_permanent_compressed(_ptr(a), nov, int(sparse), preprocessing, algo_id, gpu_num, threads,
                                    float(scaling_threshold), leaf_nov, C.byref(st))
    return _check(v, st)


def dense_ryser_range(mat, start, end, nov=None, device=0, stats: SpStats | None = None) -> float:
    a = _dmat(mat, nov)
    nov = int(round(math.sqrt(a.size)))
    st = stats if stats is not None else SpStats()
    v = lib.sp_dense_ryser_range(_ptr(a), nov, device, start, end, C.byref(st))
    return _check(v, st)


class DenseHandle:
    """Matrix resident on one device (sp_dense_open / sp_dense_run / sp_dense_close)."""

    def __init__(self, mat, nov=None, device=0):
        a = _dmat(mat, nov)
        self.nov = int(round(math.sqrt(a.size)))
        self._h = C.c_void_p()
        rc = lib.sp_dense_open(_ptr(a), self.nov, device, C.byref(self._h))
        if rc != 0:
            raise SupermanError(rc, _ffi.last_error())

    def run(self, start: int, end: int, stats: SpStats | None = None) -> float:
        st = stats if stats is not None else SpStats()
        v = lib.sp_dense_run(self._h, start, end, C.byref(st))
        return _check(v, st)

    def close(self):
        if self._h:
            lib.sp_dense_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _iarr(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def _iptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _darr(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def sparse_ryser(mat, cptrs, rows, cvals, nov=None, algo_id=4, gpu_num=1, cpu=False, threads=16,
                 stats: SpStats | None = None) -> float:
    a = _dmat(mat, nov)
    nov = int(round(math.sqrt(a.size)))
    cp, ro, cv = _iarr(cptrs), _iarr(rows), _darr(cvals)
    st = stats if stats is not None else SpStats()
    v = lib.sp_sparse_ryser(_ptr(a), _iptr(cp), _iptr(ro), _ptr(cv), nov, algo_id, gpu_num, int(cpu), threads, C.byref(st))
    return _check(v, st)


def skipper(mat, rptrs, cols, cptrs, rows, cvals, nov=None, algo_id=7, gpu_num=1, cpu=False, threads=16,
            stats: SpStats | None = None) -> float:
    a = _dmat(mat, nov)
    nov = int(round(math.sqrt(a.size)))
    rp, co, cp, ro, cv = _iarr(rptrs), _iarr(cols), _iarr(cptrs), _iarr(rows), _darr(cvals)
    st = stats if stats is not None else SpStats()
    v = lib.sp_skipper(_ptr(a), _iptr(rp), _iptr(co), _iptr(cp), _iptr(ro), _ptr(cv), nov, algo_id, gpu_num,
                       int(cpu), threads, C.byref(st))
    return _check(v, st)


def sparse_ryser_range(mat, cptrs, rows, cvals, start, end, nov=None, skipper=False, device=0,
                       stats: SpStats | None = None) -> float:
    a = _dmat(mat, nov)
    nov = int(round(math.sqrt(a.size)))
    cp, ro, cv = _iarr(cptrs), _iarr(rows), _darr(cvals)
    st = stats if stats is not None else SpStats()
    v = lib.sp_sparse_ryser_range(_ptr(a), _iptr(cp), _iptr(ro), _ptr(cv), nov, int(skipper), device, start, end, C.byref(st))
    return _check(v, st)


def rasmussen_sparse(rptrs, cols, cptrs, rows, nov, nnz, trials, gpu_num=1, seed=0, stats: SpStats | None = None) -> float:
    rp, co, cp, ro = _iarr(rptrs), _iarr(cols), _iarr(cptrs), _iarr(rows)
    st = stats if stats is not None else SpStats()
    v = lib.sp_rasmussen_sparse(_iptr(rp), _iptr(co), _iptr(cp), _iptr(ro), nov, nnz, trials, gpu_num, seed, C.byref(st))
    return _check(v, st)


def scaling_sparse(cptrs, rows, rptrs, cols, nov, nnz, trials, scale_intervals=4, scale_times=5, gpu_num=1, seed=0,
                   stats: SpStats | None = None) -> float:
    rp, co, cp, ro = _iarr(rptrs), _iarr(cols), _iarr(cptrs), _iarr(rows)
    st = stats if stats is not None else SpStats()
    v = lib.sp_scaling_sparse(_iptr(cp), _iptr(ro), _iptr(rp), _iptr(co), nov, nnz, trials, scale_intervals, scale_times,
                              gpu_num, seed, C.byref(st))
    return _check(v, st)


def rasmussen_dense(mat, nov, trials, gpu_num=1, seed=0, stats: SpStats | None = None) -> float:
    a = _dmat(mat, nov)
    st = stats if stats is not None else SpStats()
    v = lib.sp_rasmussen_dense(_ptr(a), nov, trials, gpu_num, seed, C.byref(st))
    return _check(v, st)


def scaling_dense(mat, nov, trials, scale_intervals=4, scale_times=5, gpu_num=1, seed=0, stats: SpStats | None = None) -> float:
    a = _dmat(mat, nov)
    st = stats if stats is not None else SpStats()
    v = lib.sp_scaling_dense(_ptr(a), nov, trials, scale_intervals, scale_times, gpu_num, seed, C.byref(st))
    return _check(v, st)


def approx_trials_sparse(rptrs, cols, cptrs, rows, nov, nnz, scaling=False, scale_intervals=4, scale_times=5, seed=1,
                         first=0, count=1):
    """Per-trial estimates of trials [first, first+count) (parity tests against the oracle)."""
    rp, co, cp, ro = _iarr(rptrs), _iarr(cols), _iarr(cptrs), _iarr(rows)
    out = np.zeros(count, dtype=np.float64)
    st = SpStats()
    v = lib.sp_approx_trial_sparse(_iptr(rp), _iptr(co), _iptr(cp), _iptr(ro), nov, nnz, int(scaling), scale_intervals,
                                   scale_times, seed, first, count, _ptr(out), C.byref(st))
    _check(v, st)
    return out


def approx_trials_dense(mat, nov, scaling=False, scale_intervals=4, scale_times=5, seed=1, first=0, count=1):
    """Per-trial estimates of the dense twins for trials [first, first+count)."""
    a = _dmat(mat, nov)
    out = np.zeros(count, dtype=np.float64)
    st = SpStats()
    v = lib.sp_approx_trial_dense(_ptr(a), nov, int(scaling), scale_intervals, scale_times, seed, first, count,
                                  _ptr(out), C.byref(st))
    _check(v, st)
    return out


def approx_trace_sparse(rptrs, cols, cptrs, rows, nov, nnz, scaling=False, scale_intervals=4, scale_times=5, seed=1,
                        first=0, count=1):
    """(estimates, steps completed, running product) of trials [first, first+count), one launch."""
    rp, co, cp, ro = _iarr(rptrs), _iarr(cols), _iarr(cptrs), _iarr(rows)
    out = np.zeros(count, dtype=np.float64)
    steps = np.zeros(count, dtype=np.int32)
    part = np.zeros(count, dtype=np.float64)
    st = SpStats()
    v = lib.sp_approx_trace_sparse(_iptr(rp), _iptr(co), _iptr(cp), _iptr(ro), nov, nnz, int(scaling), scale_intervals,
                                   scale_times, seed, first, count, _ptr(out), _iptr(steps), _ptr(part), C.byref(st))
    _check(v, st)
    return out, steps, part


def approx_trace_dense(mat, nov, scaling=False, scale_intervals=4, scale_times=5, seed=1, first=0, count=1):
    a = _dmat(mat, nov)
    out = np.zeros(count, dtype=np.float64)
    steps = np.zeros(count, dtype=np.int32)
    part = np.zeros(count, dtype=np.float64)
    st = SpStats()
    v = lib.sp_approx_trace_dense(_ptr(a), nov, int(scaling), scale_intervals, scale_times, seed, first, count,
                                  _ptr(out), _iptr(steps), _ptr(part), C.byref(st))
    _check(v, st)
    return out, steps, part


# ---- reference wrapper names (gpu_approximation_sparse.cu:455,497,608,663; _dense.cu:373,411,527,573)
# grid_graph is the reference's unused trailing flag.
def gpu_perman64_rasmussen_sparse(rptrs, cols, nov, nnz, number_of_times, grid_graph=False, *, cptrs, rows, seed=0):
    return rasmussen_sparse(rptrs, cols, cptrs, rows, nov, nnz, number_of_times, 1, seed)


def gpu_perman64_rasmussen_multigpucpu_chunks_sparse(cptrs, rows, rptrs, cols, nov, nnz, number_of_times, gpu_num,
                                                     cpu=False, threads=16, grid_graph=False, seed=0):
    return rasmussen_sparse(rptrs, cols, cptrs, rows, nov, nnz, number_of_times, gpu_num, seed)


def gpu_perman64_approximation_sparse(cptrs, rows, rptrs, cols, nov, nnz, number_of_times, scale_intervals, scale_times,
                                      grid_graph=False, seed=0):
    return scaling_sparse(cptrs, rows, rptrs, cols, nov, nnz, number_of_times, scale_intervals, scale_times, 1, seed)


def gpu_perman64_approximation_multigpucpu_chunks_sparse(cptrs, rows, rptrs, cols, nov, nnz, number_of_times, gpu_num,
                                                         cpu, scale_intervals, scale_times, threads=16,
                                                         grid_graph=False, seed=0):
    return scaling_sparse(cptrs, rows, rptrs, cols, nov, nnz, number_of_times, scale_intervals, scale_times, gpu_num, seed)


def gpu_perman64_rasmussen(mat, nov, number_of_times, seed=0):
    return rasmussen_dense(mat, nov, number_of_times, 1, seed)


def gpu_perman64_rasmussen_multigpucpu_chunks(mat, nov, number_of_times, gpu_num, cpu=False, threads=16, seed=0):
    return rasmussen_dense(mat, nov, number_of_times, gpu_num, seed)


def gpu_perman64_approximation(mat, nov, number_of_times, scale_intervals, scale_times, seed=0):
    return scaling_dense(mat, nov, number_of_times, scale_intervals, scale_times, 1, seed)


def gpu_perman64_approximation_multigpucpu_chunks(mat, nov, number_of_times, gpu_num, cpu, scale_intervals, scale_times,
                                                  threads=16, seed=0):
    return scaling_dense(mat, nov, number_of_times, scale_intervals, scale_times, gpu_num, seed)


# ---- reference wrapper names (gpu_exact_sparse.cu:672,732,792,853,916,995,1123,1192) ------------
def gpu_perman64_xlocal_sparse(mat, cptrs, rows, cvals, nov, grid_dim=2048, block_dim=256):
    return sparse_ryser(mat, cptrs, rows, cvals, nov, 1)


def gpu_perman64_xshared_sparse(mat, cptrs, rows, cvals, nov, grid_dim=2048, block_dim=256):
    return sparse_ryser(mat, cptrs, rows, cvals, nov, 2)


def gpu_perman64_xshared_coalescing_sparse(mat, cptrs, rows, cvals, nov, grid_dim=2048, block_dim=256):
    return sparse_ryser(mat, cptrs, rows, cvals, nov, 3)


def gpu_perman64_xshared_coalescing_mshared_sparse(mat, cptrs, rows, cvals, nov, grid_dim=2048, block_dim=256):
    return sparse_ryser(mat, cptrs, rows, cvals, nov, 4)


def gpu_perman64_xshared_coalescing_mshared_multigpu_sparse(mat, cptrs, rows, cvals, nov, gpu_num,
                                                            grid_dim=2048, block_dim=256):
    return sparse_ryser(mat, cptrs, rows, cvals, nov, 5, gpu_num)


def gpu_perman64_xshared_coalescing_mshared_multigpucpu_chunks_sparse(mat, cptrs, rows, cvals, nov, gpu_num,
                                                                      cpu=False, threads=16,
                                                                      grid_dim=2048, block_dim=256):
    return sparse_ryser(mat, cptrs, rows, cvals, nov, 6, gpu_num, cpu, threads)


def gpu_perman64_xshared_coalescing_mshared_skipper(mat, rptrs, cols, cptrs, rows, cvals, nov,
                                                    grid_dim=2048, block_dim=256):
    return skipper(mat, rptrs, cols, cptrs, rows, cvals, nov, 7)


def gpu_perman64_xshared_coalescing_mshared_multigpucpu_chunks_skipper(mat, rptrs, cols, cptrs, rows, cvals, nov,
                                                                       gpu_num, cpu=False, threads=16,
                                                                       grid_dim=2048, block_dim=256):
    return skipper(mat, rptrs, cols, cptrs, rows, cvals, nov, 8, gpu_num, cpu, threads)


# ---- reference wrapper names (gpu_exact_dense.cu:401,459,518,576,640,701,776) -------------------
# grid_dim / block_dim are accepted for signature parity and ignored: launch geometry is chosen by
# the library (RunAlgo hard-codes 2048 x {128,256}, main.cu:24-28).
def gpu_perman64_xglobal(mat, nov, grid_dim=2048, block_dim=256):
    return dense_ryser(mat, nov, 0)


def gpu_perman64_xlocal(mat, nov, grid_dim=2048, block_dim=256):
    return dense_ryser(mat, nov, 1)


def gpu_perman64_xshared(mat, nov, grid_dim=2048, block_dim=256):
    return dense_ryser(mat, nov, 2)


def gpu_perman64_xshared_coalescing(mat, nov, grid_dim=2048, block_dim=256):
    return dense_ryser(mat, nov, 3)


def gpu_perman64_xshared_coalescing_mshared(mat, nov, grid_dim=2048, block_dim=256):
    return dense_ryser(mat, nov, 4)


def gpu_perman64_xshared_coalescing_mshared_multigpu(mat, nov, gpu_num, grid_dim=2048, block_dim=256):
    return dense_ryser(mat, nov, 5, gpu_num)


def gpu_perman64_xshared_coalescing_mshared_multigpucpu_chunks(mat, nov, gpu_num, cpu=False, threads=16,
                                                               grid_dim=2048, block_dim=256):
    return dense_ryser(mat, nov, 6, gpu_num, cpu, threads)
