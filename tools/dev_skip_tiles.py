import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
import superman_b200 as sp
from superman_b200._ffi import SpStats
rng = np.random.default_rng(3)
n = 33
for kind in ("bin", "int"):
    pat = rng.random((n, n)) < 0.2
    pat[np.arange(n), rng.permutation(n)] = True
    A = pat.astype(float) if kind == "bin" else pat * rng.integers(1, 6, (n, n)).astype(float)
    for pre in (1, 2):
        m = sp.Matrix.from_dense(A).compress(pre)
        for c in (7, 8, 9, 10, 11, 12):
            for w in (2, 8):
                os.environ["SP_SPARSE_TILE_LOG2"] = str(c); os.environ["SP_SKIP_TILES_PER_LANE"] = str(w)
                st = SpStats()
                sp.skipper(m.mat, m.rptrs, m.cols, m.cptrs, m.rows, m.cvals, n, 7, stats=st)
                v = sp.skipper(m.mat, m.rptrs, m.cols, m.cptrs, m.rows, m.cvals, n, 7, stats=st)
                print("%s pre=%d c=%2d W=%d  %.3f ms visited %.1f%%  %.10e" % (kind, pre, c, w, st.kernel_ms, 100.0 * st.visited / st.units, v))
