"""Latency of small calls through the C-ABI (plan create + run + destroy on a warm device)."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
import superman_b200 as sp
rng = np.random.default_rng(0)
for n in (10, 16, 20):
    A = (rng.random((n, n)) < 0.3) * rng.integers(1, 5, (n, n)).astype(float); A[np.arange(n), np.arange(n)] = 1
    m = sp.Matrix.from_dense(A).compress(1)
    for name, fn in (("dense", lambda: sp.dense_ryser(A, n, 4)),
                     ("sparse", lambda: sp.sparse_ryser(m.mat, m.cptrs, m.rows, m.cvals, n, 4)),
                     ("skipper", lambda: sp.skipper(m.mat, m.rptrs, m.cols, m.cptrs, m.rows, m.cvals, n, 7)),
                     ("sparse dyn", lambda: sp.sparse_ryser(m.mat, m.cptrs, m.rows, m.cvals, n, 6))):
        fn(); fn()
        t = time.perf_counter()
        for _ in range(50): fn()
        print("n=%d %-10s %.3f ms/call" % (n, name, (time.perf_counter() - t) / 50 * 1e3))
