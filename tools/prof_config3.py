"""Profiling driver for BASELINE config 3 on bench.py's seeded matrices (run plain, then under ncu):
SpaRyser + SortOrder on the int and dbl matrices, SkipPer + SkipOrder on the bin matrix; plus the round-1
development matrix (rng 33000) for continuity with profiles/r01_ncu_level_engine_v2.txt."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
import bench, superman_b200 as sp
from superman_b200._ffi import SpStats
st = SpStats()
n = 33
def spa(A, pre=1):
    m = sp.Matrix.from_dense(A).compress(pre)
    for _ in range(2):
        v = sp.sparse_ryser(m.mat, m.cptrs, m.rows, m.cvals, n, 4, stats=st)
    return v, st.kernel_ms
def skip(A, pre=2):
    m = sp.Matrix.from_dense(A).compress(pre)
    for _ in range(2):
        v = sp.skipper(m.mat, m.rptrs, m.cols, m.cptrs, m.rows, m.cvals, n, 7, stats=st)
    return v, st.kernel_ms, st.visited / 2 ** 32
print("int SpaRyser+SortOrder", spa(bench.config3_matrix("int")))
print("dbl SpaRyser+SortOrder", spa(bench.config3_matrix("dbl")))
print("bin SkipPer+SkipOrder", skip(bench.config3_matrix("bin")))
rng = np.random.default_rng(33000)
A = (rng.random((n, n)) < 0.2) * rng.integers(1, 6, (n, n)).astype(float)
A[np.arange(n), rng.permutation(n)] = 1.0
print("r01 dev matrix SpaRyser+SortOrder", spa(A))
