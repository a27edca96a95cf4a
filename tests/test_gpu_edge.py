"""GPU: edge cases across every exact path -- empty rows / columns, all-zero and negative entries,
fully dense input through the sparse engines, n above the 48-row template range."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REL = 1e-9


def _all_paths(sp, A, pre=0):
    n = A.shape[0]
    m = sp.Matrix.from_dense(A).compress(pre)
    return (sp.dense_ryser(m.mat, n, 4),
            sp.sparse_ryser(m.mat, m.cptrs, m.rows, m.cvals, n, 4),
            sp.skipper(m.mat, m.rptrs, m.cols, m.cptrs, m.rows, m.cvals, n, 7))


def test_structurally_singular(sp):
    rng = np.random.default_rng(1)
    for n in (2, 5, 9, 14, 20):
        A = rng.integers(1, 4, (n, n)).astype(float)
        A0 = A.copy(); A0[n // 2, :] = 0            # empty row
        A1 = A.copy(); A1[:, 0] = 0                 # empty first (most flipped) column
        A2 = A.copy(); A2[:, n - 1] = 0             # empty last (never flipped) column
        for M in (A0, A1, A2, np.zeros((n, n))):
            sc = float(np.prod(np.abs(A).sum(axis=1)))
            for v in _all_paths(sp, M):
                assert abs(v) <= 1e-12 * sc, (n, v)


def test_identity_and_permutation_matrices(sp):
    rng = np.random.default_rng(2)
    for n in (2, 3, 8, 16, 25, 31):
        P = np.eye(n)[rng.permutation(n)]
        for pre in (0, 1, 2):
            for v in _all_paths(sp, P, pre):
                assert v == pytest.approx(1.0, abs=1e-9)
        D = np.diag(rng.integers(1, 5, n).astype(float))
        for v in _all_paths(sp, D):
            assert v == pytest.approx(float(np.prod(np.diag(D))), rel=REL)


def test_negative_entries_dense(sp, oracle):
    """the dense path takes any real matrix (real/d_ss.mtxzero has negatives); the sparse paths keep
    the reference's `> 0` CRS/CCS rule (util.h:537) and are only defined for non-negative input"""
    rng = np.random.default_rng(3)
    for n in (6, 13, 18):
        A = np.round(rng.uniform(-2, 3, (n, n)), 4)
        assert sp.dense_ryser(A, n, 4) == pytest.approx(oracle.perm_ld(A), rel=1e-8, abs=1e-12 * float(np.prod(np.abs(A).sum(axis=1))))


def test_fully_dense_through_sparse_engines(sp, oracle, monkeypatch):
    rng = np.random.default_rng(4)
    for n in (6, 8, 11, 15):
        A = rng.integers(1, 4, (n, n)).astype(float)
        want = oracle.perm_ld(A)
        for engine in ("0", "1", "2"):
            monkeypatch.setenv("SP_SPARSE_ENGINE", engine)
            for v in _all_paths(sp, A):
                assert v == pytest.approx(want, rel=REL), (n, engine)
    monkeypatch.delenv("SP_SPARSE_ENGINE")


def test_level_engine_slot_choices(sp, oracle, monkeypatch):
    rng = np.random.default_rng(5)
    n = 18
    pat = rng.random((n, n)) < 0.22
    pat[np.arange(n), rng.permutation(n)] = True
    A = pat * rng.integers(1, 5, (n, n)).astype(float)
    want = oracle.perm_ld(A)
    m = sp.Matrix.from_dense(A).compress(1)
    monkeypatch.setenv("SP_SPARSE_ENGINE", "2")
    hits = 0
    for lowcols in ("3", "4"):
        for slots in ("1", "2", "3", "4", "6", "8"):
            monkeypatch.setenv("SP_SPARSE_LOWCOLS", lowcols)
            monkeypatch.setenv("SP_LEVEL_SLOTS", slots)
            # a slot count the matrix does not fit silently falls back to the hot/cold kernel
            assert sp.sparse_ryser(m.mat, m.cptrs, m.rows, m.cvals, n, 4) == pytest.approx(want, rel=REL)
            assert sp.skipper(m.mat, m.rptrs, m.cols, m.cptrs, m.rows, m.cvals, n, 7) == pytest.approx(want, rel=REL)
            hits += 1
    assert hits == 12


def test_sparse_above_48_rows(sp, oracle):
    """n = 52 / 60 sparse: only the level engine has kernels there (N is not a template parameter);
    leading ranges against the oracle's SpaRyser restatement"""
    rng = np.random.default_rng(6)
    for n in (52, 60):
        pat = rng.random((n, n)) < 0.06
        pat[np.arange(n), rng.permutation(n)] = True
        pat[np.arange(n), np.arange(n)] = True
        A = pat * rng.integers(1, 4, (n, n)).astype(float)
        m = sp.Matrix.from_dense(A).compress(1)
        lo, hi = 1, 1 + (1 << 20)
        want = oracle.sparyser_range(m.mat, m.cptrs, m.rows, m.cvals, lo, hi)
        sc = float(np.prod(np.abs(A).sum(axis=1)))
        for skip in (False, True):
            got = sp.sparse_ryser_range(m.mat, m.cptrs, m.rows, m.cvals, lo, hi, n, skipper=skip)
            assert got == pytest.approx(want, rel=1e-9, abs=1e-13 * sc), (n, skip)


def test_every_order_30_to_50_short_range(sp, oracle):
    """sweeps n (and with it every shared-memory footprint of the sparse engines, including the ones
    just under the 48 KiB default limit) on a short range of Gray indices"""
    rng = np.random.default_rng(8)
    for n in range(30, 51):
        pat = rng.random((n, n)) < 4.0 / n
        pat[np.arange(n), rng.permutation(n)] = True
        A = pat * rng.integers(1, 4, (n, n)).astype(float)
        sc = float(np.prod(np.maximum(np.abs(A).sum(axis=1), 1)))
        for pre in (1, 2):
            m = sp.Matrix.from_dense(A).compress(pre)
            lo, hi = 1 << 15, (1 << 15) + (1 << 16)
            want = oracle.sparyser_range(m.mat, m.cptrs, m.rows, m.cvals, lo, hi)
            for skip in (False, True):
                got = sp.sparse_ryser_range(m.mat, m.cptrs, m.rows, m.cvals, lo, hi, n, skipper=skip)
                assert got == pytest.approx(want, rel=1e-9, abs=1e-13 * sc), (n, pre, skip)
            wantd = oracle.ryser_range_ld(A, lo, hi)
            assert sp.dense_ryser_range(A, lo, hi, n) == pytest.approx(wantd, rel=1e-9, abs=1e-13 * sc), n
