"""Profiling driver (run plain, then under ncu): SpaRyser + SkipPer at BASELINE config 3 size
(n = 33, density 0.2, SortOrder / SkipOrder) and the two estimators on the 36x36 grid."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
import superman_b200 as sp
from superman_b200._ffi import SpStats
n = 33
rng = np.random.default_rng(33000)
A = (rng.random((n, n)) < 0.2) * rng.integers(1, 6, (n, n)).astype(float)
A[np.arange(n), rng.permutation(n)] = 1.0
st = SpStats()
m1 = sp.Matrix.from_dense(A).compress(1)
for _ in range(2):
    v1 = sp.sparse_ryser(m1.mat, m1.cptrs, m1.rows, m1.cvals, n, 4, stats=st)
print("SpaRyser+SortOrder  %.12e  kernel_ms %.3f" % (v1, st.kernel_ms))
m2 = sp.Matrix.from_dense(A).compress(2)
for _ in range(2):
    v2 = sp.skipper(m2.mat, m2.rptrs, m2.cols, m2.cptrs, m2.rows, m2.cvals, n, 7, stats=st)
print("SkipPer+SkipOrder   %.12e  kernel_ms %.3f visited %.3e of %.3e" % (v2, st.kernel_ms, st.visited, st.units))
g = sp.Matrix.grid(36, 36)
for _ in range(2):
    r = sp.rasmussen_sparse(g.rptrs, g.cols, g.cptrs, g.rows, g.nov, g.nnz, 100000, 1, seed=1, stats=st)
print("Rasmussen 36x36 x100000  %.6e  kernel_ms %.3f  trials/s %.3e" % (r, st.kernel_ms, st.units / (st.kernel_ms * 1e-3)))
for _ in range(2):
    s = sp.scaling_sparse(g.cptrs, g.rows, g.rptrs, g.cols, g.nov, g.nnz, 100000, 4, 5, 1, seed=1, stats=st)
print("Scaling   36x36 x100000 y4 z5  %.6e  kernel_ms %.3f  trials/s %.3e" % (s, st.kernel_ms, st.units / (st.kernel_ms * 1e-3)))
g8 = sp.Matrix.grid(8, 8)
r = sp.rasmussen_sparse(g8.rptrs, g8.cols, g8.cptrs, g8.rows, g8.nov, g8.nnz, 1000000, 1, seed=1, stats=st)
print("Rasmussen 8x8 x1e6  %.6e +- %.2e (exact 12988816) kernel_ms %.3f trials/s %.3e" % (r, st.std_error, st.kernel_ms, st.units / (st.kernel_ms * 1e-3)))
s = sp.scaling_sparse(g8.cptrs, g8.rows, g8.rptrs, g8.cols, g8.nov, g8.nnz, 1000000, 4, 5, 1, seed=1, stats=st)
print("Scaling   8x8 x1e6  %.6e +- %.2e (exact 12988816) kernel_ms %.3f trials/s %.3e" % (s, st.std_error, st.kernel_ms, st.units / (st.kernel_ms * 1e-3)))
