"""BASELINE config 3 on bench.py's seeded matrices (bin / int / dbl): SpaRyser + SortOrder and SkipPer + SkipOrder,
best of three kernel times, the host model's FP64 instructions per index and the visited fraction.
SP_SPARSE_REORDER=0 switches the plan's own choice of the low columns off."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import bench, superman_b200 as sp
from superman_b200._ffi import SpStats
st = SpStats()
for kind in ("bin", "int", "dbl"):
    A = bench.config3_matrix(kind)
    for pre, skip in ((1, False), (2, True)):
        m = sp.Matrix.from_dense(A).compress(pre)
        if skip:
            f = lambda: sp.skipper(m.mat, m.rptrs, m.cols, m.cptrs, m.rows, m.cvals, 33, 7, stats=st)
        else:
            f = lambda: sp.sparse_ryser(m.mat, m.cptrs, m.rows, m.cvals, 33, 4, stats=st)
        f()
        ms = []
        for _ in range(3):
            v = f(); ms.append(st.kernel_ms)
        print("%s %-22s %.3f ms  model %.2f FP64 instr/index  visited %.3f  value %.12e"
              % (kind, "SkipPer+SkipOrder" if skip else "SpaRyser+SortOrder", min(ms), st.sq_scale, st.visited / 2 ** 32, v), flush=True)
