"""GPU dev probe: SpaRyser / SkipPer kernels vs the CPU oracle (not a test)."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import superman_b200 as sp
from superman_b200._ffi import SpStats
from _oracle import Oracle
O = Oracle()
rng = np.random.default_rng(3)

def gen(n, p, kind):
    while True:
        pat = rng.random((n, n)) < p
        if kind == "bin": A = pat.astype(float)
        elif kind == "int": A = pat * rng.integers(1, 6, (n, n)).astype(float)
        else: A = pat * np.round(rng.uniform(0.01, 5, (n, n)), 6)
        if (A.sum(0) > 0).all() and (A.sum(1) > 0).all(): return A

bad = 0
for n in (7, 9, 12, 16, 20, 22):
    for p in (0.2, 0.35, 0.6):
        for kind in ("bin", "int", "dbl"):
            A = gen(n, p, kind)
            want = O.perm_ld(A)
            for pre in (0, 1, 2):
                m = sp.Matrix.from_dense(A).compress(pre)
                st1, st2 = SpStats(), SpStats()
                g1 = sp.sparse_ryser(m.mat, m.cptrs, m.rows, m.cvals, n, 4, stats=st1)
                g2 = sp.skipper(m.mat, m.rptrs, m.cols, m.cptrs, m.rows, m.cvals, n, 7, stats=st2)
                # zero permanents come out as rounding noise: compare against the size of the Ryser terms
                tol = max(1e-9 * abs(want), 1e-13 * float(np.prod(np.abs(A).sum(axis=1))))
                ok = abs(g1 - want) <= tol and abs(g2 - want) <= tol
                if not ok:
                    bad += 1
                    print("MISMATCH", n, p, kind, pre, want, g1, g2)
            print(n, p, kind, "want %.12e spa %.12e skip %.12e visited %d/%d path %d/%d c=%d" % (want, g1, g2, st2.visited, st2.units, st1.path, st2.path, st2.tile_log2))
print("bad", bad)

# ragged ranges
n = 18
A = gen(n, 0.3, "int")
m = sp.Matrix.from_dense(A).compress(1)
full = 1 << (n - 1)
cuts = [0, 1, 777, full // 3, full // 2 + 5, full]
for skipf in (False, True):
    tot = sum(sp.sparse_ryser_range(m.mat, m.cptrs, m.rows, m.cvals, cuts[i], cuts[i + 1], n, skipper=skipf) for i in range(len(cuts) - 1))
    print("ragged", skipf, tot * sp.nw_factor(n), O.perm_ld(A))

# throughput on the reference's config-3 inputs if present in tests/golden, else synthetic n=33 p=0.2
for kind in ("bin", "int", "dbl"):
    n = 33
    A = gen(n, 0.2, kind)
    for pre, name in ((1, "SortOrder"), (2, "SkipOrder")):
        m = sp.Matrix.from_dense(A).compress(pre)
        for fn, lab in ((lambda st: sp.sparse_ryser(m.mat, m.cptrs, m.rows, m.cvals, n, 4, stats=st), "SpaRyser"),
                        (lambda st: sp.skipper(m.mat, m.rptrs, m.cols, m.cptrs, m.rows, m.cvals, n, 7, stats=st), "SkipPer")):
            st = SpStats(); fn(st); v = fn(st)
            print("n=33 p=0.2 %s %-9s %-8s perm=%.10e kernel_ms=%.2f eff it/s=%.3e visited=%.3e (%.1f%%) c=%d" % (
                kind, name, lab, v, st.kernel_ms, (1 << 32) / (st.kernel_ms * 1e-3), st.visited, 100.0 * st.visited / (1 << 32), st.tile_log2))
    st = SpStats(); sp.dense_ryser(A, n, 4, stats=st); v = sp.dense_ryser(A, n, 4, stats=st)
    print("n=33 dense                      perm=%.10e kernel_ms=%.2f" % (v, st.kernel_ms))
