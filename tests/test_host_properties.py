"""CPU, property-based (hypothesis): invariants of the C host layer that hold for ANY input."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

SET = dict(max_examples=60, deadline=None)


def _matrix(draw, nmin=2, nmax=9, zero_ok=True):
    n = draw(st.integers(nmin, nmax))
    vals = draw(st.lists(st.integers(-1 if zero_ok else 1, 4), min_size=n * n, max_size=n * n))
    return np.array([max(v, 0) for v in vals], float).reshape(n, n)


@settings(**SET)
@given(st.data())
def test_crs_ccs_describe_the_positive_entries(sp, data):
    A = _matrix(data.draw)
    n = A.shape[0]
    m = sp.Matrix.from_dense(A).compress(0)
    assert m.nnz == int((A > 0).sum())
    B = np.zeros_like(A); Cm = np.zeros_like(A)
    for i in range(n):
        cs = m.cols[m.rptrs[i]:m.rptrs[i + 1]]
        assert list(cs) == sorted(cs)
        B[i, cs] = m.rvals[m.rptrs[i]:m.rptrs[i + 1]]
        rs = m.rows[m.cptrs[i]:m.cptrs[i + 1]]
        assert list(rs) == sorted(rs)
        Cm[rs, i] = m.cvals[m.cptrs[i]:m.cptrs[i + 1]]
    assert np.array_equal(B, A) and np.array_equal(Cm, A)


@settings(**SET)
@given(st.data())
def test_orderings_are_permutations_with_the_documented_shape(sp, data):
    A = _matrix(data.draw, zero_ok=True)
    n = A.shape[0]
    m1 = sp.Matrix.from_dense(A).compress(1)          # SortOrder: column counts ascending
    counts = np.diff(m1.cptrs)
    assert list(counts) == sorted(counts)
    assert sorted(map(tuple, m1.mat.T.tolist())) == sorted(map(tuple, A.T.tolist()))   # same columns, reordered
    if (A.sum(axis=1) > 0).all():                     # SkipOrder is defined when no row is empty
        m2 = sp.Matrix.from_dense(A).compress(2)
        assert sorted(m2.mat.reshape(-1).tolist()) == sorted(A.reshape(-1).tolist())
        assert sorted(m2.mat.sum(axis=0).tolist()) == sorted(A.sum(axis=0).tolist())
        assert sorted(m2.mat.sum(axis=1).tolist()) == sorted(A.sum(axis=1).tolist())


@settings(**SET)
@given(st.data())
def test_reduce_keeps_the_permanent(sp, oracle, data):
    A = _matrix(data.draw, nmin=3, nmax=8)
    want = oracle.perm_ld(A)
    m = sp.Matrix.from_dense(A)
    f = m.reduce()
    got = f * (oracle.perm_ld(m.mat) if m.nov > 1 else m.mat[0, 0])
    assert got == pytest.approx(want, rel=1e-12, abs=1e-9)


@settings(**SET)
@given(lo=st.integers(0, 1 << 40), length=st.integers(0, 1 << 40), parts=st.integers(1, 64), align=st.integers(0, 20))
def test_partition_covers_the_range_exactly_once(sp, lo, length, parts, align):
    from superman_b200 import _ffi
    f = _ffi.lib.sp_sched_boundary
    f.restype = C.c_ulonglong
    f.argtypes = [C.c_ulonglong, C.c_ulonglong, C.c_ulonglong, C.c_ulonglong, C.c_int]
    hi = lo + length
    b = [f(lo, hi, parts, i, align) for i in range(parts + 1)]
    assert b[0] == lo and b[-1] == hi
    assert all(x <= y for x, y in zip(b, b[1:]))
    assert sum(y - x for x, y in zip(b, b[1:])) == length
