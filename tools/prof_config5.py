"""Profiling driver for BASELINE config 5: both estimators on the 36x36 grid, 100 000 trials (run plain, then under ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import superman_b200 as sp
from superman_b200._ffi import SpStats
st = SpStats()
g = sp.Matrix.grid(36, 36)
for _ in range(2):
    r = sp.rasmussen_sparse(g.rptrs, g.cols, g.cptrs, g.rows, g.nov, g.nnz, 100000, 1, seed=0, stats=st)
print("rasmussen", r, st.kernel_ms, st.visited)
for _ in range(2):
    s = sp.scaling_sparse(g.cptrs, g.rows, g.rptrs, g.cols, g.nov, g.nnz, 100000, 4, 5, 1, seed=0, stats=st)
print("scaling", s, st.kernel_ms, st.visited)
