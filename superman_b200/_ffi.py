"""ctypes binding of libsuperman_b200.so (the C-ABI declared in include/superman_b200.h).

The library is built in-tree by `make` (or `__graft_entry__.build()`); there is no Python or CPU
fallback -- if the shared object is missing the import fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SUPERMAN_B200_LIB: development only (tools/level_variants.py compares kernel variants built side by side)
LIB_PATH = os.environ.get("SUPERMAN_B200_LIB") or os.path.join(_HERE, "libsuperman_b200.so")

SP_MAX_DEVICES = 16


class SpStats(C.Structure):
    _fields_ = [
        ("kernel_ms", C.c_double),
        ("wall_ms", C.c_double),
        ("device_ms", C.c_double * SP_MAX_DEVICES),
        ("device_partial", C.c_double * SP_MAX_DEVICES),
        ("device_units", C.c_ulonglong * SP_MAX_DEVICES),
        ("units", C.c_ulonglong),
        ("visited", C.c_ulonglong),
        ("std_error", C.c_double),
        ("devices", C.c_int),
        ("chunks", C.c_int),
        ("launches", C.c_int),
        ("path", C.c_int),
        ("tile_log2", C.c_int),
        ("error", C.c_int),
        ("sumsq_scaled", C.c_double),
        ("sq_scale", C.c_double),
    ]

    def as_dict(self) -> dict:
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = list(v)[: max(1, self.devices)] if hasattr(v, "__len__") else v
        return d


class SpMatrix(C.Structure):
    """Mirror of sp_matrix (include/superman_b200.h)."""
    _fields_ = [
        ("nov", C.c_int),
        ("nnz", C.c_int),
        ("header_nnz", C.c_int),
        ("type", C.c_int),
        ("mat", C.POINTER(C.c_double)),
        ("cptrs", C.POINTER(C.c_int)),
        ("rows", C.POINTER(C.c_int)),
        ("cvals", C.POINTER(C.c_double)),
        ("rptrs", C.POINTER(C.c_int)),
        ("cols", C.POINTER(C.c_int)),
        ("rvals", C.POINTER(C.c_double)),
    ]


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -j8` (or __graft_entry__.build()); "
            "superman_b200 has no Python/CPU fallback")
    return C.CDLL(LIB_PATH)


lib = _load()

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_sp = C.POINTER(SpStats)


def _sig(name, restype, argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = argtypes
    return fn


_sig("sp_last_error", C.c_char_p, [])
_sig("sp_version", C.c_char_p, [])
_sig("sp_device_count", C.c_int, [])
_sig("sp_warmup", C.c_int, [C.c_int])
_sig("sp_set_first_device", C.c_int, [C.c_int])
_sig("sp_nw_factor", C.c_double, [C.c_int])
_sig("sp_fp64_peak", C.c_double, [C.c_int, C.c_int])
_sig("sp_int_peak", C.c_double, [C.c_int, C.c_int])
_sig("sp_set_precision", C.c_int, [C.c_int])
_sig("sp_dense_ryser", C.c_double, [_dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _sp])
_sig("sp_dense_ryser_range", C.c_double, [_dp, C.c_int, C.c_int, C.c_longlong, C.c_longlong, _sp])
_sig("sp_dense_open", C.c_int, [_dp, C.c_int, C.c_int, C.POINTER(C.c_void_p)])
_sig("sp_dense_run", C.c_double, [C.c_void_p, C.c_longlong, C.c_longlong, _sp])
_sig("sp_dense_close", None, [C.c_void_p])


_sig("sp_sparse_ryser", C.c_double, [_dp, _ip, _ip, _dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _sp])
_sig("sp_skipper", C.c_double, [_dp, _ip, _ip, _ip, _ip, _dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _sp])
_sig("sp_sparse_ryser_range", C.c_double, [_dp, _ip, _ip, _dp, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_longlong, _sp])

_sig("sp_rasmussen_sparse", C.c_double, [_ip, _ip, _ip, _ip, C.c_int, C.c_int, C.c_longlong, C.c_int, C.c_ulonglong, _sp])
_sig("sp_scaling_sparse", C.c_double, [_ip, _ip, _ip, _ip, C.c_int, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_ulonglong, _sp])
_sig("sp_rasmussen_dense", C.c_double, [_dp, C.c_int, C.c_longlong, C.c_int, C.c_ulonglong, _sp])
_sig("sp_scaling_dense", C.c_double, [_dp, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_ulonglong, _sp])
_sig("sp_approx_trial_sparse", C.c_double, [_ip, _ip, _ip, _ip, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_ulonglong, C.c_longlong, C.c_int, _dp, _sp])

_sig("sp_approx_trial_dense", C.c_double, [_dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_ulonglong, C.c_longlong, C.c_int, _dp, _sp])
_sig("sp_approx_trace_sparse", C.c_double, [_ip, _ip, _ip, _ip, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_ulonglong, C.c_longlong, C.c_int, _dp, _ip, _dp, _sp])
_sig("sp_approx_trace_dense", C.c_double, [_dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_ulonglong, C.c_longlong, C.c_int, _dp, _ip, _dp, _sp])
_sig("sp_connect", None, [])
_sig("read_calculate_return", C.c_double, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int])
_sig("matlab_calculate_return_int", C.c_double, [_ip, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int])
_sig("matlab_calculate_return_double", C.c_double, [_dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int])

_mp = C.POINTER(SpMatrix)
_sig("sp_matrix_read", C.c_int, [C.c_char_p, C.c_int, _mp])
_sig("sp_matrix_from_dense", C.c_int, [_dp, C.c_int, _mp])
_sig("sp_matrix_compress", C.c_int, [_mp, C.c_int])
_sig("sp_matrix_grid", C.c_int, [C.c_int, C.c_int, _mp])
_sig("sp_matrix_reduce", C.c_int, [_mp, _dp])
_sig("sp_matrix_reduce_step", C.c_int, [_mp, _dp])
_sig("sp_matrix_min_degree", C.c_int, [_mp])
_sig("sp_matrix_split34", C.c_int, [_mp, C.c_int, _mp])
_sig("sp_matrix_scale", C.c_int, [_mp, C.c_double, _dp, _dp])
_sig("sp_matrix_balance", C.c_int, [_mp, C.c_double, _dp, _dp])
_sig("sp_matrix_dm", C.c_int, [_mp, C.POINTER(C.c_int)])
_sig("sp_permanent_compressed", C.c_double,
     [_dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.POINTER(SpStats)])
_sig("sp_matrix_free", None, [_mp])


def last_error() -> str:
    return lib.sp_last_error().decode("utf-8", "replace")
