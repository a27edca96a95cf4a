"""Structural preprocessing of the revised front-end (host/sp_reduce.c) against
the reference's own d1compress / d2compress / d34compress / scalesk (revised_perman/util.h): the
committed fixtures of tests/golden/revised.json, and -- where oracle/_ref/libref_revised.so exists --
the unmodified reference live on seeded random matrices.  Bit-exact: same row / column choice, same
expressions.  CPU only."""
import itertools
import json
import os

import numpy as np
import pytest

from _compressed import banded, oracle_compressed

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "revised.json")


def cases(op):
    with open(GOLDEN) as f:
        return [c for c in json.load(f)["cases"] if c["op"] == op]


def mat_of(c, key="mat", n=None):
    n = c["nov"] if n is None else n
    return np.array(c[key], dtype=np.float64).reshape(n, n)


def sparse_matrix(rng, n, lo, hi, weights):
    a = np.zeros((n, n))
    for i in range(n):
        k = int(rng.integers(lo, hi + 1))
        cols = set(rng.choice(n, size=min(k, n), replace=False).tolist()) | {i}
        for j in cols:
            a[i, j] = float(rng.integers(1, 6)) if weights == "int" else round(float(rng.uniform(0.1, 5.0)), 6)
    return a


def check_step(sp, a, want, kind):
    """one reduce step of ours == the reference's, except that d1 keeps the entry in `factor`
    instead of multiplying it into row 0 (util.h:1251-1253)"""
    m = sp.Matrix.from_dense(a)
    got_kind, f = m.reduce_step()
    assert got_kind == kind
    ours = m.mat
    if kind == 1:
        ours = ours.copy()
        ours[0, :] *= f
    assert ours.shape == want.shape
    assert np.array_equal(ours, want)


def test_d1_d2_steps_match_golden(sp):
    for c in cases("d1"):
        check_step(sp, mat_of(c), mat_of(c, "out", c["nov"] - 1), 1)
    for c in cases("d2"):
        check_step(sp, mat_of(c), mat_of(c, "out", c["nov"] - 1), 2)


def test_d34_split_matches_golden(sp, oracle):
    for c in cases("d34"):
        a = mat_of(c)
        m = sp.Matrix.from_dense(a)
        assert m.min_degree() == c["min_deg"]
        second = m.split34(c["min_deg"])
        assert second is not None and m.nov == second.nov == c["nov"] - 1
        assert np.array_equal(m.mat, mat_of(c, "out", c["nov"] - 1))
        assert np.array_equal(second.mat, mat_of(c, "out2", c["nov"] - 1))
        # perm(A) = perm(A1) + perm(A2)
        assert oracle.perm_ld(m.mat) + oracle.perm_ld(second.mat) == pytest.approx(oracle.perm_ld(a), rel=1e-13)


def test_scalesk_matches_golden(sp, oracle):
    for c in cases("scalesk"):
        a = mat_of(c)
        m = sp.Matrix.from_dense(a)
        rv, cv, sweeps = m.scale(c["threshold"])
        assert sweeps >= 1
        assert np.array_equal(rv, np.array(c["rv"])) and np.array_equal(cv, np.array(c["cv"]))
        assert np.array_equal(m.mat, (a * rv[:, None]) * cv[None, :])
        # row sums of the scaled matrix are the threshold (the last sweep normalised the rows)
        assert np.allclose(m.mat.sum(axis=1), c["threshold"], rtol=1e-12)
        back = oracle.perm_ld(m.mat)
        for v in cv:
            back /= v
        for v in rv:
            back /= v
        assert back == pytest.approx(oracle.perm_ld(a), rel=1e-12)


def test_steps_match_reference_live(sp, revised):
    rng = np.random.default_rng(5)
    seen = {1: 0, 2: 0, 3: 0, 4: 0}
    for trial in range(600):
        n = int(rng.integers(6, 16))
        lo = int(rng.integers(0, 6))
        a = sparse_matrix(rng, n, lo, lo + 1, "int" if trial % 2 else "real")
        if trial % 3 == 0:
            a = a.T.copy()
        d = revised.min_nnz(a)
        m = sp.Matrix.from_dense(a)
        assert m.min_degree() == d
        if d == 1:
            check_step(sp, a, revised.d1compress(a), 1)
        elif d == 2:
            check_step(sp, a, revised.d2compress(a), 2)
        elif d in (3, 4):
            first, second = revised.d34compress(a, d)
            other = m.split34(d)
            assert np.array_equal(m.mat, first) and np.array_equal(other.mat, second)
        else:
            continue
        seen[d] += 1
    assert min(seen.values()) >= 10, seen


def test_scalesk_matches_reference_live(sp, revised):
    rng = np.random.default_rng(6)
    for trial in range(40):
        n = int(rng.integers(4, 30))
        a = sparse_matrix(rng, n, 1, 6, "real")
        thr = float(rng.choice([1.0, 2.0, 5.0, 30.0]))
        rv_ref, cv_ref = revised.scalesk(a, thr)
        rv, cv, _ = sp.Matrix.from_dense(a).scale(thr)
        assert np.array_equal(rv, rv_ref) and np.array_equal(cv, cv_ref), trial


def test_reduce_then_split_keeps_the_permanent(sp, oracle):
    """the whole recursion of sp_permanent_compressed with the CPU oracle at the leaves (reduce, then
    split while the smallest degree is 3 or 4) reproduces the permanent of the original matrix --
    provided the leaves are Sinkhorn-balanced to convergence.  Merged columns carry products of
    entries: without balancing, or with the single sweep upstream's stopping rule amounts to, the Ryser
    sum cancels catastrophically even in long double."""
    rng = np.random.default_rng(7)
    for trial in range(25):
        n = int(rng.integers(8, 15))
        a = sparse_matrix(rng, n, 2, 4, "int" if trial % 2 else "real")
        assert oracle_compressed(sp, oracle, a, leaf_nov=6) == pytest.approx(oracle.perm_ld(a), rel=1e-13), trial
    worst = {"none": 0.0, "one": 0.0, "full": 0.0}
    for n, seed in ((22, 2), (24, 0), (24, 8)):
        a = banded(np.random.default_rng(seed * 100 + n), n, "real" if seed % 2 else "int")
        want = oracle.perm_ld(a)
        for leaf in (6, 8, 10):
            for mode in worst:
                got = oracle_compressed(sp, oracle, a, leaf_nov=leaf, mode=mode)
                worst[mode] = max(worst[mode], abs(got / want - 1))
    assert worst["full"] < 1e-13
    assert worst["one"] > 1e-5 and worst["none"] > 1.0        # the reason sp_matrix_balance exists


def test_balance_converges_and_keeps_the_permanent(sp, oracle):
    rng = np.random.default_rng(12)
    for trial in range(20):
        n = int(rng.integers(4, 13))
        base = sparse_matrix(rng, n, 1, 4, "real")
        rs, cs = np.exp(rng.uniform(-12, 12, n)), np.exp(rng.uniform(-12, 12, n))
        a = base * rs[:, None] * cs[None, :]          # badly scaled: a direct Ryser sum loses ~1e-5 here
        want = oracle.perm_ld(base) * np.prod(rs) * np.prod(cs)
        m = sp.Matrix.from_dense(a)
        m.dm()
        support = m.mat.copy()
        thr = float(rng.choice([1.0, 3.0]))
        rv, cv, sweeps = m.scale(thr, converge=True)
        assert 1 <= sweeps < 1000
        b = m.mat
        assert np.allclose(b.sum(axis=1), thr, rtol=1e-12) and np.allclose(b.sum(axis=0), thr, rtol=2e-3)
        assert np.allclose(b, support * rv[:, None] * cv[None, :], rtol=1e-15)
        p = oracle.perm_ld(b)
        for i in range(n):
            p /= cv[i]
            p /= rv[i]
        assert p == pytest.approx(want, rel=1e-12), trial


def test_dm_erases_exactly_the_entries_on_no_perfect_matching(sp, oracle):
    rng = np.random.default_rng(8)
    erased_total = 0
    for trial in range(60):
        n = int(rng.integers(3, 9))
        pat = (rng.random((n, n)) < rng.choice([0.2, 0.3, 0.45])).astype(float)
        if trial % 4:
            pat[np.arange(n), rng.permutation(n)] = 1.0      # most cases have a perfect matching
        a = pat * rng.integers(1, 5, (n, n))
        m = sp.Matrix.from_dense(a)
        erased, matching = m.dm()
        # brute force over permutations of the 0/1 pattern
        on_some = np.zeros((n, n), dtype=bool)
        best = 0
        for p in itertools.permutations(range(n)):
            hits = sum(pat[i, p[i]] != 0 for i in range(n))
            best = max(best, hits)
            if hits == n:
                on_some[np.arange(n), list(p)] = True
        assert matching == best
        if best < n:
            assert erased == 0 and np.array_equal(m.mat, a)
            continue
        want = np.where(on_some, a, 0.0)
        assert np.array_equal(m.mat, want), trial
        assert erased == int((a != 0).sum() - (want != 0).sum())
        assert oracle.perm_ld(m.mat) == pytest.approx(oracle.perm_ld(a), rel=1e-13)
        erased_total += erased
    assert erased_total > 20


def test_argument_errors(sp):
    m = sp.Matrix.from_dense(np.ones((6, 6)))
    with pytest.raises(sp.SupermanError):
        m.split34(5)
    with pytest.raises(sp.SupermanError):
        m.scale(0.0)
    assert m.split34(3) is None                     # no row or column with three non-zeros
    small = sp.Matrix.from_dense(np.ones((4, 4)))
    with pytest.raises(sp.SupermanError):
        small.split34(4)
