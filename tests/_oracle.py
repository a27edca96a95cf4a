"""ctypes bindings of the CPU oracle (oracle/liboracle.so) and, when present, of the reference shim
(oracle/_ref/libref.so).  Test infrastructure: imported only from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libref.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
u64 = C.c_ulonglong


def build_oracle() -> None:
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < max(
            os.path.getmtime(os.path.join(ORACLE_DIR, f)) for f in os.listdir(ORACLE_DIR) if f.endswith(".c")):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)


def _d(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _i(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def _pd(a):
    return a.ctypes.data_as(_dp)


def _pi(a):
    return a.ctypes.data_as(_ip)


class Oracle:
    def __init__(self):
        build_oracle()
        self.lib = L = C.CDLL(ORACLE_SO)
        L.orc_nw_factor.restype = C.c_double
        L.orc_nw_factor.argtypes = [C.c_int]
        L.orc_ryser_range_f64.restype = C.c_double
        L.orc_ryser_range_f64.argtypes = [_dp, C.c_int, u64, u64]
        L.orc_ryser_range_ld_as_double.restype = C.c_double
        L.orc_ryser_range_ld_as_double.argtypes = [_dp, C.c_int, u64, u64]
        L.orc_perm_ld_as_double.restype = C.c_double
        L.orc_perm_ld_as_double.argtypes = [_dp, C.c_int]
        L.orc_perm_f64.restype = C.c_double
        L.orc_perm_f64.argtypes = [_dp, C.c_int]
        L.orc_perm_i128.restype = None
        L.orc_perm_i128.argtypes = [_ip, C.c_int, C.POINTER(C.c_longlong), C.POINTER(u64)]
        L.orc_matrix2compressed.restype = C.c_int
        L.orc_matrix2compressed.argtypes = [_dp, C.c_int, _ip, _ip, _dp, _ip, _ip, _dp]
        L.orc_sort_order.restype = C.c_int
        L.orc_sort_order.argtypes = [_dp, C.c_int, _ip, _ip, _dp, _ip, _ip, _dp, _ip]
        L.orc_skip_order.restype = C.c_int
        L.orc_skip_order.argtypes = [_dp, C.c_int, _ip, _ip, _dp, _ip, _ip, _dp, _ip, _ip]
        L.orc_grid_graph.restype = C.c_int
        L.orc_grid_graph.argtypes = [C.c_int, C.c_int, _ip]
        L.orc_kasteleyn.restype = C.c_double
        L.orc_kasteleyn.argtypes = [C.c_int, C.c_int]
        L.orc_sparyser_range_f64.restype = C.c_double
        L.orc_sparyser_range_f64.argtypes = [_dp, _ip, _ip, _dp, C.c_int, u64, u64]
        L.orc_skipper_range_f64.restype = C.c_double
        L.orc_skipper_range_f64.argtypes = [_dp, _ip, _ip, _ip, _ip, _dp, C.c_int, u64, u64, C.POINTER(u64)]
        self._approx_sigs()

    # ---- approximations ----
    def _approx_sigs(self):
        L = self.lib
        L.orc_rasmussen_trial.restype = C.c_double
        L.orc_rasmussen_trial.argtypes = [_ip, _ip, C.c_int, u64, u64]
        L.orc_scaling_trial.restype = C.c_double
        L.orc_scaling_trial.argtypes = [_ip, _ip, _ip, _ip, _dp, _dp, C.c_int, C.c_int, C.c_int, u64, u64]
        L.orc_rasmussen_trace.restype = C.c_double
        L.orc_rasmussen_trace.argtypes = [_ip, _ip, C.c_int, u64, u64, C.POINTER(C.c_int), _dp]
        L.orc_scaling_trace.restype = C.c_double
        L.orc_scaling_trace.argtypes = [_ip, _ip, _ip, _ip, _dp, _dp, C.c_int, C.c_int, C.c_int, u64, u64, C.POINTER(C.c_int), _dp]
        L.orc_philox.restype = None
        L.orc_philox.argtypes = [u64, u64, C.c_uint, C.POINTER(C.c_uint * 4)]

    def philox(self, seed, trial, block):
        out = (C.c_uint * 4)()
        self.lib.orc_philox(seed, trial, block, C.byref(out))
        return list(out)

    def rasmussen_trial(self, rptrs, cols, nov, seed, trial) -> float:
        return self.lib.orc_rasmussen_trial(_pi(_i(rptrs)), _pi(_i(cols)), nov, seed, trial)

    def scaling_trial(self, rptrs, cols, cptrs, rows, nov, y, z, seed, trial, rvals=None, cvals=None) -> float:
        rv = _pd(_d(rvals)) if rvals is not None else None
        cv = _pd(_d(cvals)) if cvals is not None else None
        return self.lib.orc_scaling_trial(_pi(_i(rptrs)), _pi(_i(cols)), _pi(_i(cptrs)), _pi(_i(rows)), rv, cv, nov, y, z, seed, trial)

    def rasmussen_trace(self, rptrs, cols, nov, seed, trial):
        """(estimate, steps completed, running product at that point) of one trial"""
        st = C.c_int(); part = C.c_double()
        v = self.lib.orc_rasmussen_trace(_pi(_i(rptrs)), _pi(_i(cols)), nov, seed, trial, C.byref(st), C.byref(part))
        return v, st.value, part.value

    def scaling_trace(self, rptrs, cols, cptrs, rows, nov, y, z, seed, trial, rvals=None, cvals=None):
        rv = _pd(_d(rvals)) if rvals is not None else None
        cv = _pd(_d(cvals)) if cvals is not None else None
        st = C.c_int(); part = C.c_double()
        v = self.lib.orc_scaling_trace(_pi(_i(rptrs)), _pi(_i(cols)), _pi(_i(cptrs)), _pi(_i(rows)), rv, cv, nov, y, z, seed,
                                       trial, C.byref(st), C.byref(part))
        return v, st.value, part.value

    # ---- dense ----
    def perm_ld(self, mat) -> float:
        a = _d(mat); n = a.shape[0]
        return self.lib.orc_perm_ld_as_double(_pd(a), n)

    def perm_f64(self, mat) -> float:
        a = _d(mat); n = a.shape[0]
        return self.lib.orc_perm_f64(_pd(a), n)

    def ryser_range_f64(self, mat, lo, hi) -> float:
        a = _d(mat); n = a.shape[0]
        return self.lib.orc_ryser_range_f64(_pd(a), n, lo, hi)

    def ryser_range_ld(self, mat, lo, hi) -> float:
        a = _d(mat); n = a.shape[0]
        return self.lib.orc_ryser_range_ld_as_double(_pd(a), n, lo, hi)

    def perm_i128(self, mat) -> int:
        a = _i(mat); n = a.shape[0]
        hi = C.c_longlong(); lo = u64()
        self.lib.orc_perm_i128(_pi(a), n, C.byref(hi), C.byref(lo))
        return (hi.value << 64) | lo.value

    # ---- preprocessing ----
    def compress(self, mat, preprocessing=0):
        """returns dict(mat (possibly permuted), cptrs, rows, cvals, rptrs, cols, rvals, nnz, colperm, rowperm)"""
        a = _d(mat).copy(); n = a.shape[0]
        cap = n * n
        cptrs = np.zeros(n + 1, np.int32); rptrs = np.zeros(n + 1, np.int32)
        rows = np.zeros(cap, np.int32); cols = np.zeros(cap, np.int32)
        cvals = np.zeros(cap); rvals = np.zeros(cap)
        colperm = np.arange(n, dtype=np.int32); rowperm = np.arange(n, dtype=np.int32)
        if preprocessing == 1:
            nnz = self.lib.orc_sort_order(_pd(a), n, _pi(cptrs), _pi(rows), _pd(cvals), _pi(rptrs), _pi(cols), _pd(rvals), _pi(colperm))
        elif preprocessing == 2:
            nnz = self.lib.orc_skip_order(_pd(a), n, _pi(cptrs), _pi(rows), _pd(cvals), _pi(rptrs), _pi(cols), _pd(rvals), _pi(rowperm), _pi(colperm))
        else:
            nnz = self.lib.orc_matrix2compressed(_pd(a), n, _pi(cptrs), _pi(rows), _pd(cvals), _pi(rptrs), _pi(cols), _pd(rvals))
        return dict(mat=a, cptrs=cptrs, rows=rows[:nnz].copy(), cvals=cvals[:nnz].copy(), rptrs=rptrs,
                    cols=cols[:nnz].copy(), rvals=rvals[:nnz].copy(), nnz=nnz, colperm=colperm, rowperm=rowperm)

    def grid_graph(self, m, n):
        nov = m * n // 2
        mat = np.zeros((nov, nov), np.int32)
        nnz = self.lib.orc_grid_graph(m, n, _pi(mat))
        return mat, nnz

    def kasteleyn(self, m, n) -> float:
        return self.lib.orc_kasteleyn(m, n)

    # ---- sparse ----
    def sparyser_range(self, mat, cptrs, rows, cvals, lo, hi) -> float:
        a = _d(mat); n = a.shape[0]
        return self.lib.orc_sparyser_range_f64(_pd(a), _pi(_i(cptrs)), _pi(_i(rows)), _pd(_d(cvals)), n, lo, hi)

    def skipper_range(self, mat, rptrs, cols, cptrs, rows, cvals, lo, hi):
        a = _d(mat); n = a.shape[0]
        vis = u64()
        v = self.lib.orc_skipper_range_f64(_pd(a), _pi(_i(rptrs)), _pi(_i(cols)), _pi(_i(cptrs)), _pi(_i(rows)),
                                           _pd(_d(cvals)), n, lo, hi, C.byref(vis))
        return v, vis.value

    @staticmethod
    def nw_base_prod(mat) -> float:
        a = _d(mat); n = a.shape[0]
        x = a[:, n - 1] - a.sum(axis=1) / 2
        p = 1.0
        for v in x:
            p *= v
        return p


class Reference:
    """The unmodified reference (util.h / algo.h) behind oracle/_ref/libref.so."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        self.lib = L = C.CDLL(REF_SO)
        L.ref_perman64.restype = C.c_double
        L.ref_perman64.argtypes = [_dp, C.c_int]
        L.ref_parallel_perman64.restype = C.c_double
        L.ref_parallel_perman64.argtypes = [_dp, C.c_int, C.c_int]
        L.ref_parallel_perman64_int.restype = C.c_double
        L.ref_parallel_perman64_int.argtypes = [_ip, C.c_int, C.c_int]
        L.ref_parallel_perman64_sparse.restype = C.c_double
        L.ref_parallel_perman64_sparse.argtypes = [_dp, _ip, _ip, _dp, C.c_int, C.c_int]
        for name in ("ref_parallel_skip_perman64_w", "ref_parallel_skip_perman64_w_balanced"):
            f = getattr(L, name)
            f.restype = C.c_double
            f.argtypes = [_ip, _ip, _dp, _ip, _ip, _dp, C.c_int, C.c_int]
        L.ref_matrix2compressed.restype = None
        L.ref_matrix2compressed.argtypes = [_dp, C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, _ip, _ip, _dp]
        L.ref_gridGraph2compressed.restype = C.c_int
        L.ref_gridGraph2compressed.argtypes = [C.c_int, C.c_int, _ip, _ip, _ip, _ip, _ip]
        L.ref_read_matrix.restype = C.c_int
        L.ref_read_matrix.argtypes = [C.c_char_p, C.c_int, _dp, C.c_int, _ip, _ip]
        L.ref_max_threads.restype = C.c_int

    def perman64(self, mat) -> float:
        a = _d(mat)
        return self.lib.ref_perman64(_pd(a), a.shape[0])

    def parallel_perman64(self, mat, threads) -> float:
        a = _d(mat)
        return self.lib.ref_parallel_perman64(_pd(a), a.shape[0], threads)

    def parallel_perman64_int(self, mat, threads) -> float:
        a = _i(mat)
        return self.lib.ref_parallel_perman64_int(_pi(a), a.shape[0], threads)

    def compress(self, mat, preprocessing=0):
        a = _d(mat).copy(); n = a.shape[0]
        nnz = int((a > 0).sum())
        cptrs = np.zeros(n + 1, np.int32); rptrs = np.zeros(n + 1, np.int32)
        rows = np.zeros(nnz, np.int32); cols = np.zeros(nnz, np.int32)
        cvals = np.zeros(nnz); rvals = np.zeros(nnz)
        self.lib.ref_matrix2compressed(_pd(a), n, nnz, preprocessing, _pi(cptrs), _pi(rows), _pd(cvals), _pi(rptrs), _pi(cols), _pd(rvals))
        return dict(mat=a, cptrs=cptrs, rows=rows, cvals=cvals, rptrs=rptrs, cols=cols, rvals=rvals, nnz=nnz)

    def sparse(self, c, threads=1) -> float:
        n = c["mat"].shape[0]
        return self.lib.ref_parallel_perman64_sparse(_pd(c["mat"]), _pi(c["cptrs"]), _pi(c["rows"]), _pd(c["cvals"]), n, threads)

    def skipper(self, c, threads=1, balanced=True) -> float:
        n = c["mat"].shape[0]
        f = self.lib.ref_parallel_skip_perman64_w_balanced if balanced else self.lib.ref_parallel_skip_perman64_w
        return f(_pi(c["rptrs"]), _pi(c["cols"]), _pd(c["rvals"]), _pi(c["cptrs"]), _pi(c["rows"]), _pd(c["cvals"]), n, threads)

    def grid_graph(self, m, n):
        nov = m * n // 2
        mat = np.zeros((nov, nov), np.int32)
        cap = nov * 4 + 8
        cptrs = np.zeros(nov + 1, np.int32); rptrs = np.zeros(nov + 1, np.int32)
        rows = np.zeros(cap, np.int32); cols = np.zeros(cap, np.int32)
        nnz = self.lib.ref_gridGraph2compressed(m, n, _pi(mat), _pi(cptrs), _pi(rows), _pi(rptrs), _pi(cols))
        return mat, nnz, cptrs, rows[:max(nnz, 0)], rptrs, cols[:max(nnz, 0)]

    def read_matrix(self, path, generic=True):
        cap = 128 * 128
        buf = np.zeros(cap)
        nnz = C.c_int(); typ = C.c_int()
        nov = self.lib.ref_read_matrix(path.encode(), int(generic), _pd(buf), cap, C.byref(nnz), C.byref(typ))
        if nov < 0:
            raise IOError(path)
        return buf[: nov * nov].reshape(nov, nov).copy(), nnz.value, ("int", "float", "double")[typ.value]


def have_reference() -> bool:
    return os.path.exists(REF_SO)


REF_REVISED_SO = os.path.join(ORACLE_DIR, "_ref", "libref_revised.so")


def have_revised_reference() -> bool:
    return os.path.exists(REF_REVISED_SO)


class RevisedReference:
    """The structural preprocessing of the reference's revised front-end (revised_perman/util.h)
    behind oracle/_ref/libref_revised.so."""

    def __init__(self):
        if not os.path.exists(REF_REVISED_SO):
            raise FileNotFoundError(REF_REVISED_SO)
        self.lib = L = C.CDLL(REF_REVISED_SO)
        L.rev_min_nnz.restype = C.c_int
        L.rev_min_nnz.argtypes = [_dp, C.c_int]
        for name in ("rev_d1compress", "rev_d2compress"):
            f = getattr(L, name)
            f.restype = C.c_int
            f.argtypes = [_dp, C.c_int]
        L.rev_d34compress.restype = C.c_int
        L.rev_d34compress.argtypes = [_dp, C.c_int, _dp, C.c_int]
        L.rev_scalesk.restype = None
        L.rev_scalesk.argtypes = [C.c_int, C.c_int, _ip, _ip, _dp, _ip, _ip, _dp, C.c_double, _dp, _dp]

    def min_nnz(self, mat) -> int:
        a = _d(mat)
        return self.lib.rev_min_nnz(_pd(a), a.shape[0])

    def _step(self, fn, mat):
        a = _d(mat).copy(); n = a.shape[0]
        flat = a.reshape(-1).copy()
        k = fn(_pd(flat), n)
        return flat[:k * k].reshape(k, k).copy()

    def d1compress(self, mat):
        return self._step(self.lib.rev_d1compress, mat)

    def d2compress(self, mat):
        return self._step(self.lib.rev_d2compress, mat)

    def d34compress(self, mat, min_deg):
        a = _d(mat).copy(); n = a.shape[0]
        flat = a.reshape(-1).copy()
        other = np.zeros(n * n, dtype=np.float64)
        k = self.lib.rev_d34compress(_pd(flat), n, _pd(other), min_deg)
        if k == n:
            return None
        return flat[:k * k].reshape(k, k).copy(), other[:k * k].reshape(k, k).copy()

    def scalesk(self, mat, threshold):
        """(rv, cv) from the CRS/CCS of mat in natural order"""
        a = _d(mat); n = a.shape[0]
        rptrs = [0]; cols = []; rvals = []
        for i in range(n):
            for j in range(n):
                if a[i, j] > 0:
                    cols.append(j); rvals.append(a[i, j])
            rptrs.append(len(cols))
        cptrs = [0]; rows = []; cvals = []
        for j in range(n):
            for i in range(n):
                if a[i, j] > 0:
                    rows.append(i); cvals.append(a[i, j])
            cptrs.append(len(rows))
        cptrs, rows, rptrs, cols = _i(cptrs), _i(rows), _i(rptrs), _i(cols)
        cvals, rvals = _d(cvals), _d(rvals)
        rv = np.zeros(n); cv = np.zeros(n)
        self.lib.rev_scalesk(n, len(rows), _pi(cptrs), _pi(rows), _pd(cvals), _pi(rptrs), _pi(cols), _pd(rvals),
                             float(threshold), _pd(rv), _pd(cv))
        return rv, cv
