"""Profiling driver (run plain, then under ncu): SpaRyser + SortOrder at BASELINE config 3 size."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
import superman_b200 as sp
n = 33
rng = np.random.default_rng(33000)
A = (rng.random((n, n)) < 0.2) * rng.integers(1, 6, (n, n)).astype(float)
A[np.arange(n), rng.permutation(n)] = 1.0
st = sp.SpStats()
m1 = sp.Matrix.from_dense(A).compress(1)
for _ in range(2):
    v1 = sp.sparse_ryser(m1.mat, m1.cptrs, m1.rows, m1.cvals, n, 4, stats=st)
print("SpaRyser+SortOrder  %.12e  kernel_ms %.3f" % (v1, st.kernel_ms))
