"""GPU: the reference's Python / MATLAB shim entry points (interface_connector.c:61-231) on the GPU engine."""
import ctypes as C

import numpy as np
import pytest

import _golden

pytestmark = pytest.mark.gpu


def test_read_calculate_return_and_matlab_entry_points(sp, tmp_path):
    from superman_b200 import _ffi
    lib = _ffi.lib
    e = _golden.small()[3]          # n = 12, int
    p = tmp_path / "m.txt"
    _golden.write_matrix_file(e, p)
    A = _golden.dense_from(e)
    n = e["n"]
    for algo in (4, 5, 6, 7, 8):    # exact ids of decide_and_call (interface_connector.c:37-51)
        v = lib.read_calculate_return(str(p).encode(), algo, 4, 1000, 4, 5)
        assert v == pytest.approx(e["ld"], rel=1e-9), algo       # a double, not truncated through int
    ai = np.ascontiguousarray(A.astype(np.int32))
    ad = np.ascontiguousarray(A)
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    assert lib.matlab_calculate_return_int(ai.ctypes.data_as(ip), 5, 1, 0, 0, 0, n, int((A != 0).sum())) == pytest.approx(e["ld"], rel=1e-9)
    assert lib.matlab_calculate_return_double(ad.ctypes.data_as(dp), 7, 1, 0, 0, 0, n, int((A != 0).sum())) == pytest.approx(e["ld"], rel=1e-9)
    exact_bin = float(e["i128_binary"])
    for algo in (0, 1, 2, 3):       # estimators see the pattern only
        v = lib.read_calculate_return(str(p).encode(), algo, 4, 60000, 4, 5)
        assert v == pytest.approx(exact_bin, rel=0.2), algo
    assert np.isnan(lib.read_calculate_return(str(p).encode(), 42, 1, 1, 1, 1))
    assert np.isnan(lib.read_calculate_return(b"/nonexistent", 5, 1, 1, 1, 1))
