import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import bench, superman_b200 as sp
from superman_b200._ffi import SpStats
n = 36
A = bench.synthetic_matrix(n, 0.5)
for c in (8, 9, 10, 11):
    for g in (2, 4, 8, 16, 32, 64):
        os.environ["SP_DENSE_TILE_LOG2"] = str(c); os.environ["SP_DENSE_UNUSED"] = str(g)
        with sp.DenseHandle(A, n) as h:
            st = SpStats(); h.run(0, 1 << 35, st); v = h.run(0, 1 << 35, st)
        print("c=%2d gpb=%2d  %.3f ms  %.4e it/s  rel.err %.1e" % (c, g, st.kernel_ms, (1 << 35) / (st.kernel_ms * 1e-3), abs(v * sp.nw_factor(n) / 4.8452758461437975e43 - 1)))
