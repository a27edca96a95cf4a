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


@settings(**SET)
@given(st.data())
def test_dm_split_and_balance_keep_the_permanent(sp, oracle, data):
    """Dulmage-Mendelsohn erasure, the d34 split and Sinkhorn balancing are exact for ANY matrix"""
    A = _matrix(data.draw, nmin=5, nmax=8)
    n = A.shape[0]
    want = oracle.perm_ld(A)
    # DM: never changes the permanent; with a perfect matching the result has total support
    m = sp.Matrix.from_dense(A)
    erased, matching = m.dm()
    assert 0 <= matching <= n and erased >= 0
    if matching < n:
        assert want == pytest.approx(0.0, abs=1e-9) and np.array_equal(m.mat, A)
    else:
        assert oracle.perm_ld(m.mat) == pytest.approx(want, rel=1e-12, abs=1e-9)
        assert (m.mat <= A).all() and erased == int((A != 0).sum() - (m.mat != 0).sum())
        # balancing the total support: row sums exact, permanent recovered from the factors
        rv, cv, sweeps = m.scale(1.0, converge=True)
        assert sweeps >= 1 and np.allclose(m.mat.sum(axis=1), 1.0, rtol=1e-12)
        p = np.longdouble(oracle.perm_ld(m.mat))
        for i in range(n):
            p /= np.longdouble(cv[i]); p /= np.longdouble(rv[i])
        assert float(p) == pytest.approx(want, rel=1e-10, abs=1e-9)
    # d34 split on the smallest degree, when it is 3 or 4
    m = sp.Matrix.from_dense(A)
    d = m.min_degree()
    if d in (3, 4):
        other = m.split34(d)
        assert other is not None and m.nov == other.nov == n - 1
        assert oracle.perm_ld(m.mat) + oracle.perm_ld(other.mat) == pytest.approx(want, rel=1e-12, abs=1e-9)
