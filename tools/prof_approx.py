"""Profiling driver for the estimators (run plain, then under ncu): Rasmussen and the scaled estimator
on the 36x36 grid (BASELINE config 5) and the 8x8 grid, 2^17 trials each."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import superman_b200 as sp
from superman_b200._ffi import SpStats
st = SpStats()
T = int(os.environ.get("TRIALS", 1 << 17))
for gm in [int(x) for x in os.environ.get("GRIDS", "36,8").split(",")]:
    name = "%dx%d" % (gm, gm)
    g = sp.Matrix.grid(gm, gm)
    for _ in range(2):
        r = sp.rasmussen_sparse(g.rptrs, g.cols, g.cptrs, g.rows, g.nov, g.nnz, T, 1, seed=1, stats=st)
    print("Rasmussen %s x%d  %.6e  kernel_ms %.3f  trials/s %.3e" % (name, T, r, st.kernel_ms, st.units / (st.kernel_ms * 1e-3)))
    for _ in range(2):
        s = sp.scaling_sparse(g.cptrs, g.rows, g.rptrs, g.cols, g.nov, g.nnz, T, 4, 5, 1, seed=1, stats=st)
    print("Scaling   %s x%d y4 z5  %.6e  kernel_ms %.3f  trials/s %.3e" % (name, T, s, st.kernel_ms, st.units / (st.kernel_ms * 1e-3)))
