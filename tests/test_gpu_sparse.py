"""GPU parity tests of SpaRyser / SkipPer (ids -s -p1..-p8) through the C-ABI."""
import numpy as np
import pytest

import _golden

pytestmark = pytest.mark.gpu
REL = 1e-9


def _rand(rng, n, p, kind):
    pat = rng.random((n, n)) < p
    pat[np.arange(n), rng.permutation(n)] = True
    if kind == "bin":
        return pat.astype(float)
    if kind == "int":
        return pat * rng.integers(1, 6, (n, n)).astype(float)
    return pat * np.round(rng.uniform(0.01, 5, (n, n)), 6)


def _scale(A):
    return float(np.prod(np.abs(A).sum(axis=1)))


@pytest.mark.parametrize("n", [2, 3, 5, 7, 8, 10, 13, 16, 18, 21])
def test_parity_vs_oracle(sp, oracle, n):
    rng = np.random.default_rng(500 + n)
    for p in (0.2, 0.4):
        for kind in ("bin", "int", "dbl"):
            A = _rand(rng, n, p, kind)
            want = oracle.perm_ld(A)
            tol = dict(rel=REL, abs=1e-13 * _scale(A))
            for pre in (0, 1, 2):
                m = sp.Matrix.from_dense(A).compress(pre)
                for algo in (1, 2, 3, 4, 5, 6):
                    assert sp.sparse_ryser(m.mat, m.cptrs, m.rows, m.cvals, n, algo) == pytest.approx(want, **tol), (n, p, kind, pre, algo)
                st = sp._ffi.SpStats()
                for algo in (7, 8):
                    got = sp.skipper(m.mat, m.rptrs, m.cols, m.cptrs, m.rows, m.cvals, n, algo, stats=st)
                    assert got == pytest.approx(want, **tol), (n, p, kind, pre, algo)
                    assert 0 < st.visited <= st.units == 1 << (n - 1)


def test_matches_reference_restatement_bitwise_class(sp, oracle):
    """against the oracle's restatement of the reference's own sparse arithmetic
    (incremental product with divide; zero-skipping), same CCS / CRS inputs"""
    rng = np.random.default_rng(77)
    n = 17
    A = _rand(rng, n, 0.3, "int")
    full = 1 << (n - 1)
    for pre in (1, 2):
        m = sp.Matrix.from_dense(A).compress(pre)
        base = oracle.ryser_range_f64(m.mat, 0, 1)
        spa = (base + oracle.sparyser_range(m.mat, m.cptrs, m.rows, m.cvals, 1, full)) * sp.nw_factor(n)
        skp, vis = oracle.skipper_range(m.mat, m.rptrs, m.cols, m.cptrs, m.rows, m.cvals, 1, full)
        skp = (base + skp) * sp.nw_factor(n)
        assert sp.sparse_ryser(m.mat, m.cptrs, m.rows, m.cvals, n, 4) == pytest.approx(spa, rel=REL)
        assert sp.skipper(m.mat, m.rptrs, m.cols, m.cptrs, m.rows, m.cvals, n, 7) == pytest.approx(skp, rel=REL)


def test_ranges(sp, oracle):
    rng = np.random.default_rng(31)
    n = 19
    A = _rand(rng, n, 0.3, "int")
    m = sp.Matrix.from_dense(A).compress(2)
    full = 1 << (n - 1)
    want = oracle.perm_ld(A)
    cuts = [0, 1, 300, 5000, full // 3 + 1, full // 2, full - 2, full]
    for skip in (False, True):
        tot = sum(sp.sparse_ryser_range(m.mat, m.cptrs, m.rows, m.cvals, cuts[i], cuts[i + 1], n, skipper=skip)
                  for i in range(len(cuts) - 1))
        assert tot * sp.nw_factor(n) == pytest.approx(want, rel=REL)
        assert sp.sparse_ryser_range(m.mat, m.cptrs, m.rows, m.cvals, 9, 9, n, skipper=skip) == 0.0


def test_skipper_really_skips_and_paths_agree(sp, monkeypatch):
    rng = np.random.default_rng(2)
    n = 24
    A = _rand(rng, n, 0.18, "bin")
    m = sp.Matrix.from_dense(A).compress(2)
    st = sp._ffi.SpStats()
    a = sp.sparse_ryser(m.mat, m.cptrs, m.rows, m.cvals, n, 4)
    b = sp.skipper(m.mat, m.rptrs, m.cols, m.cptrs, m.rows, m.cvals, n, 7, stats=st)
    assert round(a) == round(b)
    assert st.visited < st.units // 2 and st.path == 5
    for lowcols in ("3", "4"):
        monkeypatch.setenv("SP_SPARSE_LOWCOLS", lowcols)
        assert round(sp.skipper(m.mat, m.rptrs, m.cols, m.cptrs, m.rows, m.cvals, n, 7)) == round(a)
    monkeypatch.delenv("SP_SPARSE_LOWCOLS")
    monkeypatch.setenv("SP_SPARSE_FORCE_SMEM", "1")
    assert round(sp.sparse_ryser(m.mat, m.cptrs, m.rows, m.cvals, n, 4)) == round(a)


def test_golden_corpus_config3(sp):
    """BASELINE.json configs[2]: SpaRyser + SortOrder (-s -p4 -r1) and SkipPer + SkipOrder (-s -p7 -r2)
    on the corpus files, against the stored long-double permanents"""
    c = _golden.corpus()
    for name, e in sorted(c.items()):
        A = _golden.dense_from(e)
        n = e["n"]
        m1 = sp.Matrix.from_dense(A).compress(1)
        assert np.diff(m1.cptrs).tolist() == e["colcount_sort"], name
        assert sp.sparse_ryser(m1.mat, m1.cptrs, m1.rows, m1.cvals, n, 4) == pytest.approx(e["ld"], rel=REL), name
        m2 = sp.Matrix.from_dense(A).compress(2)
        assert np.diff(m2.cptrs).tolist() == e["colcount_skip"], name
        assert sp.skipper(m2.mat, m2.rptrs, m2.cols, m2.cptrs, m2.rows, m2.cvals, n, 7) == pytest.approx(e["ld"], rel=REL), name
        if e.get("ld_binary") is not None:
            mb = sp.Matrix.from_dense((A != 0).astype(float)).compress(2)
            got = sp.skipper(mb.mat, mb.rptrs, mb.cols, mb.cptrs, mb.rows, mb.cvals, n, 7)
            assert got == pytest.approx(e["ld_binary"], rel=REL), name


def test_errors(sp):
    m = sp.Matrix.from_dense(np.ones((5, 5))).compress(0)
    with pytest.raises(sp.SupermanError):
        sp.sparse_ryser(m.mat, m.cptrs, m.rows, m.cvals, 5, 9)
    with pytest.raises(sp.SupermanError):
        sp.skipper(m.mat, m.rptrs, m.cols, m.cptrs, m.rows, m.cvals, 5, 4)
    bad_rows = m.rows.copy(); bad_rows[0] = 99
    with pytest.raises(sp.SupermanError):
        sp.sparse_ryser(m.mat, m.cptrs, bad_rows, m.cvals, 5, 4)


def test_reference_real_matrices(sp):
    """the reference's real/ inputs that the exact paths can handle: ibm32 (n = 32, 0/1) against the
    long-double oracle value -- an exact integer -- and cage5_c2 (n = 37, real-valued): the dense,
    SpaRyser+SortOrder and SkipPer+SkipOrder paths walk different Gray sequences over differently
    permuted matrices and must agree"""
    r = _golden.real()
    if not r:
        pytest.skip("tests/golden/real.json not generated")
    e = r["real/ibm32.mtxzero"]
    A = _golden.dense_from(e)
    n = e["n"]
    assert e["ld"] == round(e["ld"])
    m1 = sp.Matrix.from_dense(A).compress(1)
    m2 = sp.Matrix.from_dense(A).compress(2)
    assert np.diff(m1.cptrs).tolist() == e["colcount_sort"] and np.diff(m2.cptrs).tolist() == e["colcount_skip"]
    assert round(sp.dense_ryser(A, n, 4)) == e["ld"]
    assert round(sp.sparse_ryser(m1.mat, m1.cptrs, m1.rows, m1.cvals, n, 4)) == e["ld"]
    st = sp._ffi.SpStats()
    assert round(sp.skipper(m2.mat, m2.rptrs, m2.cols, m2.cptrs, m2.rows, m2.cvals, n, 7, stats=st)) == e["ld"]
    assert st.visited < st.units
    e = r["real/cage5_c2.mtxzero"]
    A = _golden.dense_from(e)
    n = e["n"]
    d = sp.dense_ryser(A, n, 4)
    m1 = sp.Matrix.from_dense(A).compress(1)
    m2 = sp.Matrix.from_dense(A).compress(2)
    assert sp.sparse_ryser(m1.mat, m1.cptrs, m1.rows, m1.cvals, n, 4) == pytest.approx(d, rel=REL)
    assert sp.skipper(m2.mat, m2.rptrs, m2.cols, m2.cptrs, m2.rows, m2.cvals, n, 7) == pytest.approx(d, rel=REL)


def test_permanents_recorded_by_the_reference_authors(sp):
    """known-answer vectors: the twelve 32x32 Erdos-Renyi 0/1 matrices of the reference's SkipPer kit
    (revised_perman/sparyser/ErdosRenyi) with the permanents its authors recorded in
    revised_perman/sparyser/Results/*.out for seven algorithm / ordering variants each
    (tests/golden/erdos.json, made by tests/golden/make_erdos_golden.py).  The recorded variants agree
    with each other to ~1e-10; every exact engine here must land inside that band and on the
    long-double oracle value.  (For p = 0.40 and 0.50 the kit printed 0: those permanents exceed 2^64 and
    its integer conversion overflowed -- only the long-double value is checked there.)"""
    import _golden
    c = _golden.erdos()
    assert len(c) == 12
    known = 0
    for name, e in sorted(c.items()):
        n = e["n"]
        A = _golden.dense_from(e)
        rec = sorted(float(v) for v in e["recorded"].values() if float(v) > 0)
        if rec:
            assert len(rec) >= 5 and rec[-1] == pytest.approx(rec[0], rel=1e-9), name
            known += 1
        else:
            assert e["ld"] > 2.0 ** 64, name              # the kit's integer print overflowed
            rec = [e["ld"]]
        mid = rec[len(rec) // 2]
        m1 = sp.Matrix.from_dense(A).compress(1)
        m2 = sp.Matrix.from_dense(A).compress(2)
        st = sp.SpStats()
        results = {
            "dense": sp.dense_ryser(A, n, 4),
            "sparyser+sort": sp.sparse_ryser(m1.mat, m1.cptrs, m1.rows, m1.cvals, n, 4),
            "skipper+skiporder": sp.skipper(m2.mat, m2.rptrs, m2.cols, m2.cptrs, m2.rows, m2.cvals, n, 7, stats=st),
            "compressed": sp.permanent_compressed(A, sparse=True, preprocessing=1, algo_id=4),
        }
        assert st.visited < st.units                      # 0/1 rows: SkipPer really skips
        for how, got in results.items():
            assert got == pytest.approx(mid, rel=1e-9), (name, how, got, mid)
            assert got == pytest.approx(e["ld"], rel=1e-9), (name, how)
            assert rec[0] - 1e-9 * mid <= got <= rec[-1] + 1e-9 * mid, (name, how)
    assert known == 6
