"""Small end-to-end pass over every kernel (for compute-sanitizer; sizes kept tiny)."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
import superman_b200 as sp
rng = np.random.default_rng(1)
for n in (9, 14, 17):
    A = (rng.random((n, n)) < 0.5) * rng.integers(1, 5, (n, n)).astype(float)
    A[np.arange(n), np.arange(n)] = 1.0
    d = sp.dense_ryser(A, n, 4)
    r = sp.dense_ryser_range(A, 3, (1 << (n - 1)) - 5, n)
    m = sp.Matrix.from_dense(A).compress(2)
    s = sp.sparse_ryser(m.mat, m.cptrs, m.rows, m.cvals, n, 4)
    k = sp.skipper(m.mat, m.rptrs, m.cols, m.cptrs, m.rows, m.cvals, n, 7)
    print(n, d, s, k)
g = sp.Matrix.grid(6, 6)
print(sp.rasmussen_sparse(g.rptrs, g.cols, g.cptrs, g.rows, g.nov, g.nnz, 2000, 1, seed=1))
print(sp.scaling_sparse(g.cptrs, g.rows, g.rptrs, g.cols, g.nov, g.nnz, 2000, 4, 5, 1, seed=1))
B = (rng.random((10, 10)) < 0.6).astype(float); B[np.arange(10), np.arange(10)] = 1
print(sp.scaling_dense(B * 1.5, 10, 500, 4, 5, 1, seed=1))
