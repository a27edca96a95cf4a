"""GPU parity tests of the dense Ryser path, all through the C-ABI (libsuperman_b200.so)."""
import math
import os

import numpy as np
import pytest

import _golden

pytestmark = pytest.mark.gpu

REL = 1e-9   # BASELINE.json north_star: "matching the reference's own double-precision result within 1e-9 relative"


def _rand(rng, n, p, kind):
    pat = rng.random((n, n)) < p
    pat[np.arange(n), rng.permutation(n)] = True          # permanent of the pattern is non-zero
    if kind == "bin":
        return pat.astype(float)
    if kind == "int":
        return pat * rng.integers(1, 6, (n, n)).astype(float)
    return pat * np.round(rng.uniform(0.01, 5, (n, n)), 6)


def _scale(A):
    """a magnitude of the Ryser terms: results whose true value is (near) zero are compared against this"""
    return float(np.prod(np.abs(A).sum(axis=1)))


@pytest.mark.parametrize("n", [1, 2, 3, 4, 6, 7, 8, 9, 12, 15, 16, 19, 22])
def test_parity_vs_oracle_all_ids(sp, oracle, n):
    rng = np.random.default_rng(100 + n)
    for kind in ("bin", "int", "dbl"):
        A = _rand(rng, n, 0.55, kind)
        want = oracle.perm_ld(A)
        for algo in (0, 1, 2, 3, 4, 5, 6):
            got = sp.dense_ryser(A, n, algo, gpu_num=1)
            assert got == pytest.approx(want, rel=REL, abs=1e-13 * _scale(A)), (n, kind, algo)
        if kind != "dbl" and n <= 20:
            exact = oracle.perm_i128(A.astype(int))
            got = sp.dense_ryser(A, n, 4)
            # FP64 Ryser is not integer-exact: exact after rounding while the
            # value is small, else 1e-12 relative
            assert abs(got - float(exact)) <= max(0.5, 1e-12 * float(exact)), (n, kind, got, exact)


def test_reference_named_wrappers(sp, oracle):
    rng = np.random.default_rng(3)
    n = 13
    A = _rand(rng, n, 0.6, "dbl")
    want = oracle.perm_ld(A)
    for fn in (sp.gpu_perman64_xglobal, sp.gpu_perman64_xlocal, sp.gpu_perman64_xshared,
               sp.gpu_perman64_xshared_coalescing, sp.gpu_perman64_xshared_coalescing_mshared):
        assert fn(A, n, 2048, 128) == pytest.approx(want, rel=REL)
    assert sp.gpu_perman64_xshared_coalescing_mshared_multigpu(A, n, 1, 2048, 128) == pytest.approx(want, rel=REL)
    assert sp.gpu_perman64_xshared_coalescing_mshared_multigpucpu_chunks(A, n, 1, False, 16, 2048, 128) == pytest.approx(want, rel=REL)


def test_register_and_shared_memory_kernels_agree(sp, oracle, monkeypatch):
    rng = np.random.default_rng(8)
    for n in (8, 17, 24):
        A = _rand(rng, n, 0.5, "dbl")
        want = oracle.perm_ld(A)
        for b in ("3", "4"):
            monkeypatch.setenv("SP_DENSE_LOWCOLS", b)
            assert sp.dense_ryser(A, n, 4) == pytest.approx(want, rel=REL)
        monkeypatch.delenv("SP_DENSE_LOWCOLS")
        monkeypatch.setenv("SP_DENSE_FORCE_SMEM", "1")
        assert sp.dense_ryser(A, n, 4) == pytest.approx(want, rel=REL)
        monkeypatch.delenv("SP_DENSE_FORCE_SMEM")


def test_ranges_ragged_and_empty(sp, oracle):
    """kernel-level (start, end) contract: arbitrary, unaligned, empty and single-index ranges"""
    rng = np.random.default_rng(21)
    n = 20
    A = _rand(rng, n, 0.5, "dbl")
    full = 1 << (n - 1)
    sc = _scale(A)
    for lo, hi in [(0, 0), (0, 1), (1, 2), (5, 5), (0, full), (1, full), (12345, 12346), (777, 300000),
                   (full - 3, full), (1 << 14, 1 << 15), ((1 << 14) - 1, (1 << 15) + 1)]:
        got = sp.dense_ryser_range(A, lo, hi, n)
        want = oracle.ryser_range_ld(A, lo, hi)
        assert got == pytest.approx(want, rel=1e-9, abs=1e-15 * sc), (lo, hi)
    cuts = [0, 1, 4097, full // 3, full // 2 + 9, full - 1, full]
    tot = sum(sp.dense_ryser_range(A, cuts[i], cuts[i + 1], n) for i in range(len(cuts) - 1))
    assert tot * sp.nw_factor(n) == pytest.approx(oracle.perm_ld(A), rel=REL)
    with pytest.raises(sp.SupermanError):
        sp.dense_ryser_range(A, 0, full + 1, n)
    with pytest.raises(sp.SupermanError):
        sp.dense_ryser_range(A, 9, 3, n)


def test_resident_handle_is_deterministic(sp):
    rng = np.random.default_rng(4)
    n = 26
    A = _rand(rng, n, 0.5, "dbl")
    with sp.DenseHandle(A, n) as h:
        a = h.run(0, 1 << (n - 1))
        b = h.run(0, 1 << (n - 1))
    assert a == b                       # fixed summation tree: bit-reproducible
    assert sp.dense_ryser(A, n, 4) == a * sp.nw_factor(n)


def test_golden_corpus(sp):
    """the reference's own corpus files named by BASELINE.json configs 0-2, against the long-double
    oracle values stored in tests/golden (and the reference's recorded double results)"""
    c = _golden.corpus()
    assert c
    for name, e in sorted(c.items()):
        A = _golden.dense_from(e)
        got = sp.dense_ryser(A, e["n"], 4)
        assert got == pytest.approx(e["ld"], rel=REL), name
        if e.get("ref_perman64") is not None:
            assert got == pytest.approx(e["ref_perman64"], rel=3e-8), name    # the serial chain is the noisy side
        if e.get("ld_binary") is not None:
            gotb = sp.dense_ryser((A != 0).astype(float), e["n"], 4)
            assert gotb == pytest.approx(e["ld_binary"], rel=REL), name


def test_golden_corpus_wide(sp, tmp_path):
    """30 more files of the reference corpus (int/, float/, double/; n = 30, 31; densities 0.1 ... 0.9),
    read back through the library's own reader: dense Ryser, SpaRyser + SortOrder and SkipPer +
    SkipOrder against the long-double oracle values of tests/golden/corpus_wide.json"""
    c = _golden.corpus_wide()
    assert len(c) >= 20
    for name, e in sorted(c.items()):
        p = tmp_path / name.replace("/", "_")
        _golden.write_matrix_file(e, p)
        n = e["n"]
        want = e["ld"]
        m = sp.Matrix.read(str(p))
        assert m.nov == n and m.type == e["type"]
        # a file without a perfect matching (density 0.1) has permanent 0: every method returns rounding
        # noise of its 2^(n-1) Ryser terms, the long-double oracle included -- bound it by the term size
        _, matching = sp.Matrix.read(str(p)).dm()
        if matching < n:
            a = np.abs(m.mat)
            noise = (1 << (n - 1)) * n * 1.2e-16 * float(np.prod(a.sum(axis=1) / 2 + a.max(axis=1)))
            check = lambda v: abs(v) <= noise and abs(want) <= noise
        else:
            check = lambda v: v == pytest.approx(want, rel=REL)
        assert check(sp.dense_ryser(m.mat, n, 4)), name
        m1 = sp.Matrix.read(str(p)).compress(1)
        assert check(sp.sparse_ryser(m1.mat, m1.cptrs, m1.rows, m1.cvals, n, 4)), name
        m2 = sp.Matrix.read(str(p)).compress(2)
        assert check(sp.skipper(m2.mat, m2.rptrs, m2.cols, m2.cptrs, m2.rows, m2.cvals, n, 7)), name


def test_small_golden_files_via_reader(sp, tmp_path):
    for idx, e in enumerate(_golden.small()):
        p = tmp_path / ("g%d.txt" % idx)
        _golden.write_matrix_file(e, p)
        m = sp.Matrix.read(str(p)).compress(0)
        assert sp.dense_ryser(m.mat, m.nov, 4) == pytest.approx(e["ld"], rel=REL)
        mb = sp.Matrix.read(str(p), binary=True).compress(0)
        assert round(sp.dense_ryser(mb.mat, mb.nov, 4)) == int(e["i128_binary"])


def test_errors(sp):
    with pytest.raises(sp.SupermanError):
        sp.dense_ryser(np.ones((4, 4)), 4, 9)          # "Unknown Algorithm ID"
    with pytest.raises(sp.SupermanError):
        sp.dense_ryser(np.ones((65, 65)), 65, 4)
    with pytest.raises(sp.SupermanError):
        sp.dense_ryser(np.ones((22, 22)), 22, 5, gpu_num=64)     # more devices than the box has


def test_large_n_paths(sp, oracle, monkeypatch):
    """49 <= n <= 64: register kernel at 2-3 blocks/SM, and the shared-memory-X kernel with > 48 KiB of
    dynamic shared memory (the reference cannot launch there): leading ranges
    against the oracle"""
    rng = np.random.default_rng(6)
    for n in (49, 52, 57, 64):
        A = _rand(rng, n, 0.3, "dbl")
        lo, hi = 3, 3 + (1 << 18)
        want = oracle.ryser_range_ld(A, lo, hi)
        got = sp.dense_ryser_range(A, lo, hi, n)                  # ragged head + register body + tail
        assert got == pytest.approx(want, rel=1e-9, abs=1e-14 * _scale(A))
        lo2, hi2 = 1 << 20, (1 << 20) + (1 << 19)
        want2 = oracle.ryser_range_ld(A, lo2, hi2)
        assert sp.dense_ryser_range(A, lo2, hi2, n) == pytest.approx(want2, rel=1e-9, abs=1e-14 * _scale(A))
        monkeypatch.setenv("SP_DENSE_FORCE_SMEM", "1")
        assert sp.dense_ryser_range(A, lo, hi, n) == pytest.approx(want, rel=1e-9, abs=1e-14 * _scale(A))
        monkeypatch.delenv("SP_DENSE_FORCE_SMEM")


# ---- BASELINE.json full sizes: size-independent properties ------------------------------------------
def test_n36_properties(sp):
    """n = 36 (2^35 Gray indices, ~0.14 s per permanent): the GPU value must be invariant under
    transposition and row / column permutations (each walks a different Gray sequence), scale as
    c^n under a scalar, be additive over a split of the index space and identical for the static
    and dynamic partitions."""
    import bench
    n = 36
    A = bench.synthetic_matrix(n, 0.5)
    st = sp._ffi.SpStats()
    p0 = sp.dense_ryser(A, n, 4, stats=st)
    assert st.units == 1 << 35 and st.path == 1
    assert math.isfinite(p0) and p0 > 0
    # the headline workload itself against the long-double oracle (tests/golden/bench36.json, made by
    # tests/golden/make_bench_golden.py: 2^35 indices in x87 long double, 17 min on 8 cores)
    import json
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bench36.json")))
    assert g["n"] == 36 and g["checksum"] == float(A.sum())
    assert p0 == pytest.approx(g["ld"], rel=REL)
    assert abs(p0 / g["ld"] - 1.0) < 2e-10           # observed 3e-11
    rng = np.random.default_rng(1)
    assert sp.dense_ryser(A.T.copy(), n, 4) == pytest.approx(p0, rel=REL)
    assert sp.dense_ryser(A[rng.permutation(n)][:, rng.permutation(n)].copy(), n, 4) == pytest.approx(p0, rel=REL)
    assert sp.dense_ryser(2.0 * A, n, 4) == pytest.approx(p0 * 2.0 ** n, rel=REL)
    D = np.diag(rng.uniform(0.5, 2.0, n))
    assert sp.dense_ryser(D @ A, n, 4) == pytest.approx(p0 * float(np.prod(np.diag(D))), rel=REL)
    full = 1 << 35
    cuts = [0, 12345678901, full // 2 + 17, full]
    tot = sum(sp.dense_ryser_range(A, cuts[i], cuts[i + 1], n) for i in range(3))
    assert tot * sp.nw_factor(n) == pytest.approx(p0, rel=REL)
    assert sp.dense_ryser(A, n, 6, gpu_num=1) == pytest.approx(p0, rel=REL)


def test_n40_known_answers(sp, oracle):
    """n = 40 (2^39 Gray indices, ~2.5 s per permanent), the other half of BASELINE config 4: no CPU oracle
    reaches this size, so the GPU value is checked against permanents known in closed form, computed in
    exact rational arithmetic from the very (dyadic) parameters the float64 matrix is built from
    (tests/_closed_forms.py): D1 (J - I) D2 -> derangements(40) prod d1 prod d2; u v^T + diag(d) -> a
    polynomial identity; a hidden block-diagonal matrix -> product of exact __int128 permanents.
    Observed relative errors 3e-13 .. 2e-10; the bar is the north star's 1e-9."""
    import _closed_forms as cf
    n = 40
    rng = np.random.default_rng(40)
    cases = [("derangement", cf.derangement_matrix(rng, n)),
             ("rank1+diag", cf.rank1_plus_diag(rng, n)),
             ("blocks 20+20 int", cf.block_diagonal(rng, oracle, [20, 20], "int")),
             ("blocks 13+13+14 bin", cf.block_diagonal(rng, oracle, [13, 13, 14], "bin"))]
    for name, (A, exact) in cases:
        st = sp._ffi.SpStats()
        got = sp.dense_ryser(A, n, 4, stats=st)
        assert st.units == 1 << 39 and st.path == 1, name
        assert abs(got / float(exact) - 1.0) < REL, (name, got, float(exact))
    # the same permanent along different Gray sequences and through the other ids
    A, exact = cases[1][1]
    ex = float(exact)
    assert abs(sp.dense_ryser(A.T.copy(), n, 4) / ex - 1.0) < REL
    assert abs(sp.dense_ryser(A[rng.permutation(n)][:, rng.permutation(n)].copy(), n, 5, gpu_num=1) / ex - 1.0) < REL
    assert abs(sp.dense_ryser(A, n, 6, gpu_num=1) / ex - 1.0) < REL
    # additive over a ragged split of the 2^39 indices
    full = 1 << 39
    cuts = [0, 123456789012, full // 2 + 4097, full]
    tot = sum(sp.dense_ryser_range(A, cuts[i], cuts[i + 1], n) for i in range(3))
    assert abs(tot * sp.nw_factor(n) / ex - 1.0) < REL


def test_quad_precision_mode(sp, oracle):
    """-q of the revised front-end (flags.calculation_quad, revised_perman/flags.h:61-64) = double-double
    arithmetic in the dense kernel: (1) on a generic real matrix it agrees with the long-double oracle to the
    oracle's own precision; (2) on ragged ranges it is additive; (3) on chesapeake (39 x 39, 0/1), where the
    FP64 sum is 2e-6 off because the permanent is 1e-12 of the terms it is the sum of, it returns the exact
    integer 13 173 481 190 272 (tests/golden/known_perman.json, Python-integer DP)."""
    import _golden
    rng = np.random.default_rng(77)
    n = 18
    A = _rand(rng, n, 0.6, "dbl")
    want = oracle.perm_ld(A)
    fp64 = sp.dense_ryser(A, n, 4)
    sp.set_precision(True)
    try:
        st = sp._ffi.SpStats()
        got = sp.dense_ryser(A, n, 4, stats=st)
        assert st.path == 8                                   # SPD_PATH_DENSE_DD
        assert got == pytest.approx(want, rel=5e-16)          # long double carries 64 bits, double-double 106
        assert abs(got / want - 1.0) <= abs(fp64 / want - 1.0) + 1e-16
        full = 1 << (n - 1)
        cuts = [0, 1, 12345, full // 2 + 3, full]
        tot = sum(sp.dense_ryser_range(A, cuts[i], cuts[i + 1], n) for i in range(4))
        assert tot * sp.nw_factor(n) == pytest.approx(want, rel=5e-16)
        for algo in (5, 6):
            assert sp.dense_ryser(A, n, algo, gpu_num=1) == pytest.approx(want, rel=5e-16)
        # orders around the block kernel's limits (n < 6 and ranges shorter than 2048 indices take the loop kernel)
        for m in (3, 5, 6, 7, 12, 13):
            Am = _rand(rng, m, 0.8, "dbl")
            assert sp.dense_ryser(Am, m, 4) == pytest.approx(oracle.perm_ld(Am), rel=5e-16), m
        e = _golden.known_perman()["chesapeake"]
        a = _golden.dense_from(e)
        exact = int(e["exact"])
        v = sp.dense_ryser(a, 39, 4, stats=st)
        assert st.units == 1 << 38 and st.path == 8
        assert round(v) == exact, (v, exact)
    finally:
        sp.set_precision(False)
    v64 = sp.dense_ryser(a, 39, 4)
    assert 1e-8 < abs(v64 / exact - 1.0) < 1e-4                # what FP64 Ryser delivers on this matrix


def test_multi_device_partitions(sp, oracle):
    ndev = sp.device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs")
    rng = np.random.default_rng(12)
    n = 27
    A = _rand(rng, n, 0.5, "dbl")
    want = sp.dense_ryser(A, n, 4)
    for g in range(2, min(ndev, 8) + 1):
        st = sp._ffi.SpStats()
        assert sp.dense_ryser(A, n, 5, gpu_num=g, stats=st) == pytest.approx(want, rel=1e-11)
        assert st.devices == g and sum(st.device_units[:g]) == 1 << (n - 1)
        assert sp.dense_ryser(A, n, 6, gpu_num=g, stats=st) == pytest.approx(want, rel=1e-11)
