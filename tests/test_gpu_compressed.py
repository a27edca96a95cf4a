"""sp_permanent_compressed (the revised front-end's -o / -u path: degree compression, d34 splits,
Sinkhorn scaling; host/sp_reduce.c) on the GPU engine against the CPU oracle.  The truth for n > 30
is the same recursion evaluated with the long-double oracle at small leaves (pinned against the
direct long-double permanent on the CPU in tests/test_host_reduce.py)."""
import os
import subprocess

import numpy as np
import pytest

from _compressed import banded, oracle_compressed
from test_host_reduce import sparse_matrix

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PERMAN = os.path.join(ROOT, "superman_b200", "perman")


@pytest.mark.parametrize("n,weights,sparse,algo,pre", [
    (33, "real", False, 4, 0), (34, "int", True, 4, 1), (36, "real", True, 7, 2), (38, "int", False, 4, 0),
    (40, "real", True, 4, 1)])
def test_compressed_matches_oracle(sp, oracle, n, weights, sparse, algo, pre):
    rng = np.random.default_rng(100 + n)
    a = banded(rng, n, weights)
    want = oracle_compressed(sp, oracle, a)
    st = sp.SpStats()
    got = sp.permanent_compressed(a, sparse=sparse, preprocessing=pre, algo_id=algo, stats=st)
    assert got == pytest.approx(want, rel=1e-9), (got, want)
    assert st.chunks >= 1 and st.error == 0
    # explicit threshold (-u 4) and no scaling at all agree too (looser without scaling)
    got_u = sp.permanent_compressed(a, sparse=sparse, preprocessing=pre, algo_id=algo, scaling_threshold=4.0)
    assert got_u == pytest.approx(want, rel=1e-9)
    # upstream's default (no scaling) runs, but FP64 Ryser on the merged matrices is not trustworthy
    got_plain = sp.permanent_compressed(a, sparse=sparse, preprocessing=pre, algo_id=algo, scaling_threshold=-1.0)
    assert np.isfinite(got_plain)
    # a deeper recursion (more, smaller leaves) gives the same permanent
    st2 = sp.SpStats()
    got_deep = sp.permanent_compressed(a, sparse=sparse, preprocessing=pre, algo_id=algo, leaf_nov=20, stats=st2)
    assert got_deep == pytest.approx(want, rel=1e-9)
    assert st2.chunks >= st.chunks


def test_compressed_equals_direct_on_a_matrix_without_structure(sp, oracle):
    rng = np.random.default_rng(9)
    n = 20
    a = (rng.random((n, n)) < 0.6) * rng.uniform(0.5, 2.0, (n, n))
    a[np.arange(n), np.arange(n)] = 1.0
    assert (a != 0).sum(axis=0).min() >= 5 and (a != 0).sum(axis=1).min() >= 5
    st = sp.SpStats()
    got = sp.permanent_compressed(a, stats=st)
    assert st.chunks == 1
    assert got == sp.dense_ryser(a)                       # untouched matrix: same kernel, same bits
    assert got == pytest.approx(oracle.perm_ld(a), rel=1e-10)
    # scaling only (leaf_nov < 0), as -u without -o
    got_u = sp.permanent_compressed(a, scaling_threshold=2.0, leaf_nov=-1)
    assert got_u == pytest.approx(oracle.perm_ld(a), rel=1e-10)


def test_compressed_zero_and_tiny(sp):
    z = np.ones((9, 9)); z[:, 4] = 0
    assert sp.permanent_compressed(z) == 0.0
    assert sp.permanent_compressed(np.eye(12) * 2.0) == 4096.0          # all d1 steps, 1x1 leaf
    tri = np.triu(np.ones((10, 10)))
    assert sp.permanent_compressed(tri, sparse=True, algo_id=4, preprocessing=1) == 1.0


def write_matrix(path, a, kind):
    n = a.shape[0]
    nz = [(i, j) for i in range(n) for j in range(n) if a[i, j] != 0]
    with open(path, "w") as f:
        f.write(f"{n} {len(nz)} {kind}\n")
        for i, j in nz:
            f.write(f"{i} {j} {int(a[i, j]) if kind == 'int' else repr(float(a[i, j]))}\n")


def run_cli(*args):
    env = dict(os.environ, PERMAN_PRECISION="17")
    r = subprocess.run([PERMAN, *args], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("Result17:")][-1]
    return float(line.split()[2]), r.stdout


def test_cli_compression_scaling_and_dm_flags(sp, oracle, tmp_path):
    rng = np.random.default_rng(36)
    a = banded(rng, 36, "int")
    want = oracle_compressed(sp, oracle, a)
    path = str(tmp_path / "banded36.txt")
    write_matrix(path, a, "int")
    for extra in (("-p", "4", "-o"), ("-s", "-p", "4", "-r", "1", "-o"), ("-s", "-p", "7", "-r", "2", "-o", "-u", "2"),
                  ("-p", "4", "--reduce", "--dm")):
        got, out = run_cli("-f", path, *extra)
        assert got == pytest.approx(want, rel=1e-9), (extra, got, want)
        assert "Compressed:" in out
        if "--dm" in extra:
            assert "DM: maximum matching 36 of 36" in out
    # -u alone: Sinkhorn scaling, no compression -- on a small dense matrix
    b = sparse_matrix(rng, 18, 8, 12, "real")
    pathb = str(tmp_path / "dense18.txt")
    write_matrix(pathb, b, "double")
    got, out = run_cli("-f", pathb, "-p", "4", "-u", "3")
    assert got == pytest.approx(oracle.perm_ld(b), rel=1e-10)
    assert "Compressed: 1 leaf" in out


def test_compressed_on_the_reference_corpus(sp):
    """the reference's own sparse files (tests/golden/corpus*.json: long-double permanents made by
    tests/golden/make_golden.py): -o shrinks them and lands on the same permanent"""
    import _golden
    seen = 0
    for name, entry in _golden.corpus().items():
        if "_0.20_" not in name and "_0.30_" not in name:
            continue
        a = _golden.dense_from(entry)
        n = entry["n"]
        for sparse, algo, pre in ((True, 4, 1), (True, 7, 2), (False, 4, 0)):
            st = sp.SpStats()
            got = sp.permanent_compressed(a, sparse=sparse, preprocessing=pre, algo_id=algo, stats=st)
            assert got == pytest.approx(entry["ld"], rel=1e-9), (name, sparse, algo, got, entry["ld"])
            assert st.units < (1 << (n - 1)), (name, st.units)        # fewer Gray indices than the direct run
        seen += 1
    assert seen >= 3


def test_leaves_are_spread_over_the_devices(sp, oracle):
    """multi-device ids (-p5 / -p6 / -p8): with at least as many pending leaves as devices every device
    computes whole leaves; the sum runs in leaf order, so the result does not depend on the split"""
    g = sp.device_count()
    if g < 2:
        pytest.skip("needs >= 2 GPUs")
    a = banded(np.random.default_rng(140), 40, "real")
    want = oracle_compressed(sp, oracle, a)
    one = sp.permanent_compressed(a, sparse=False, algo_id=4, leaf_nov=24)
    assert one == pytest.approx(want, rel=1e-9)
    for sparse, algo, pre in ((False, 5, 0), (False, 6, 0), (True, 5, 1), (True, 8, 2)):
        st = sp.SpStats()
        got = sp.permanent_compressed(a, sparse=sparse, preprocessing=pre, algo_id=algo, gpu_num=g, leaf_nov=24, stats=st)
        assert got == pytest.approx(want, rel=1e-9), (sparse, algo)
        assert st.chunks >= g and st.devices == g
        used = [d for d in range(g) if st.device_units[d] > 0]
        assert len(used) == g, (sparse, algo, used)              # every device took leaves
        again = sp.permanent_compressed(a, sparse=sparse, preprocessing=pre, algo_id=algo, gpu_num=g, leaf_nov=24)
        assert again == got                                       # whichever device took which leaf
    # fewer leaves than devices: each leaf is split over the devices by the id's own entry point
    st = sp.SpStats()
    few = sp.permanent_compressed(a, sparse=False, algo_id=5, gpu_num=g, leaf_nov=39, stats=st)
    assert few == pytest.approx(want, rel=1e-9)


def test_compressed_random_matrices_against_the_direct_permanent(sp, oracle):
    """end to end: C recursion + GPU leaves against the long-double permanent of the ORIGINAL matrix
    (no recursion on the oracle side), small leaves so that every kind of step happens many times"""
    rng = np.random.default_rng(2026)
    kinds = 0
    for trial in range(36):
        n = int(rng.integers(14, 25))
        if trial % 3 == 0:
            a = banded(rng, n, "int" if trial % 2 else "real")
        else:
            a = sparse_matrix(rng, n, 1 + trial % 3, 3 + trial % 3, "int" if trial % 2 else "real")
        if trial % 5 == 0:
            a = a.T.copy()
        want = oracle.perm_ld(a)
        sparse, algo, pre = ((False, 4, 0), (True, 4, 1), (True, 7, 2))[trial % 3]
        st = sp.SpStats()
        got = sp.permanent_compressed(a, sparse=sparse, preprocessing=pre, algo_id=algo, leaf_nov=int(rng.integers(6, 12)), stats=st)
        assert got == pytest.approx(want, rel=1e-9, abs=1e-9), (trial, n, got, want)
        kinds += st.chunks > 1
    assert kinds >= 10          # the recursion really split


def test_real_matrices_with_recorded_permanents(sp):
    """chesapeake (39x39, 340 entries) and will57 (57x57, 281 entries) from
    revised_perman/elektrik_matrices/known_perman, with the permanents the reference's SkipPer kit
    recorded (revised_perman/sparyser/RealResults/*.out, 260-1600 s of CPU time each) and the CPU
    long-double recursion value of tests/golden/make_known_perman_ld.py when present.

    What the numbers say (profiles/r01_real_known_perman.log): on chesapeake the kit's five variants
    spread over 1e-5, and so do the DIRECT FP64 Ryser runs here (dense 2.4e-6, SpaRyser 3.6e-7 away) --
    a 0/1 permanent of 1.3e13 is what is left of 2^38 terms of size up to 1e25, whoever sums them.  The
    -o path (balanced leaves of order <= 30) is the well-conditioned one: it returns the same value to
    14 digits for ten different reduction trees (transpose, permutations, leaf sizes).  will57 is only
    reachable through -o (2^56 indices directly): ~2300 leaves, 2 s on one B200, the same value to 15
    digits over ten reduction trees; the kit's two recorded values disagree with each other by 6 % and
    are 6.5-6.9 times larger -- and wrong: the exact integer permanents of both matrices (Python-integer
    transfer-matrix DP, tools/exact_permanent_dp.py) are 13 173 481 190 272 and 1 070 536 592 880 585 216,
    which the -o path reproduces to 1e-12."""
    import _golden
    d = _golden.known_perman()
    assert set(d) >= {"chesapeake", "will57"}

    e = d["chesapeake"]
    a = _golden.dense_from(e)
    n = e["n"]
    assert n == 39 and int((a != 0).sum()) == 340
    rec = sorted(float(r["perman"]) for r in e["recorded"].values())
    st = sp.SpStats()
    trees = [sp.permanent_compressed(a, sparse=True, preprocessing=1, algo_id=4, stats=st),
             sp.permanent_compressed(a, sparse=False, algo_id=4, leaf_nov=28),
             sp.permanent_compressed(a.T.copy(), sparse=True, preprocessing=1, algo_id=4, leaf_nov=33)]
    assert st.chunks > 100
    ref = trees[0]
    for v in trees:
        assert v == pytest.approx(ref, rel=1e-12)
    assert abs(ref - round(ref)) < 0.5                          # a 0/1 permanent below 2^53
    assert rec[0] * (1 - 2e-5) <= ref <= rec[-1] * (1 + 2e-5), (ref, rec)
    m1 = sp.Matrix.from_dense(a).compress(1)
    direct = sp.sparse_ryser(m1.mat, m1.cptrs, m1.rows, m1.cvals, n, 4)
    assert direct == pytest.approx(ref, rel=1e-5)              # direct FP64 Ryser: conditioning, see above
    # -u alone (Sinkhorn-balance the whole matrix, no compression) repairs the direct sum
    scaled = sp.permanent_compressed(a, sparse=True, preprocessing=1, algo_id=4, scaling_threshold=1.0, leaf_nov=-1)
    assert scaled == pytest.approx(ref, rel=1e-9)
    if "ld_recursion" in e:
        assert ref == pytest.approx(e["ld_recursion"], rel=1e-11)
    # the exact value: tools/exact_permanent_dp.py, a transfer-matrix DP in Python integers that shares
    # nothing with the library (checked against the __int128 Ryser oracle in tests/test_oracle.py)
    assert int(e["exact"]) == 13173481190272
    assert round(ref) == int(e["exact"])
    assert abs(ref / float(int(e["exact"])) - 1.0) < 1e-12

    e = d["will57"]
    a = _golden.dense_from(e)
    assert e["n"] == 57 and int((a != 0).sum()) == 281
    st = sp.SpStats()
    v1 = sp.permanent_compressed(a, sparse=True, preprocessing=1, algo_id=4, stats=st)
    v2 = sp.permanent_compressed(a, sparse=False, algo_id=4, leaf_nov=27)
    v3 = sp.permanent_compressed(a.T.copy(), sparse=True, preprocessing=2, algo_id=7, leaf_nov=32)
    assert st.chunks > 1000 and st.error == 0
    assert v2 == pytest.approx(v1, rel=1e-12) and v3 == pytest.approx(v1, rel=1e-12)
    if "ld_recursion" in e:
        assert v1 == pytest.approx(e["ld_recursion"], rel=1e-11)
    # exact: 1 070 536 592 880 585 216 (tools/exact_permanent_dp.py, 0.1 s in Python integers).  The two values
    # the reference's kit recorded for this matrix (7.39e18 and 6.95e18, 6 % apart from each other) are wrong.
    exact = int(e["exact"])
    assert exact == 1070536592880585216
    for v in (v1, v2, v3):
        assert abs(v / float(exact) - 1.0) < 1e-12, (v, exact)
    for r in e["recorded"].values():
        assert float(r["perman"]) / exact > 6.0


def test_extreme_row_and_column_scales(sp, oracle):
    """rows scaled by 1e+-120 and columns by 1e+-80: the direct Ryser sum overflows, the balanced one
    (-u, or -o on a matrix the compression touches) does not -- its Sinkhorn factors are undone in long
    double on the host"""
    rng = np.random.default_rng(77)
    n = 16
    base = sparse_matrix(rng, n, 4, 7, "real")
    rs = np.array([1e120 if i % 2 else 1e-120 for i in range(n)])
    cs = np.array([1e-80 if j % 2 else 1e80 for j in range(n)])
    a = base * rs[:, None] * cs[None, :]
    want = float(np.longdouble(oracle.perm_ld(base)) * np.prod(rs.astype(np.longdouble)) * np.prod(cs.astype(np.longdouble)))
    assert np.isfinite(want) and want > 0
    assert not np.isfinite(sp.dense_ryser(a)) or sp.dense_ryser(a) != pytest.approx(want, rel=1e-6)   # direct: hopeless
    got = sp.permanent_compressed(a, scaling_threshold=1.0, leaf_nov=-1)
    assert got == pytest.approx(want, rel=1e-10)
    got = sp.permanent_compressed(a, sparse=True, preprocessing=1, algo_id=4, scaling_threshold=2.0, leaf_nov=8)
    assert got == pytest.approx(want, rel=1e-10)
