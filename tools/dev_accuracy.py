"""GPU dev probe: error of the dense kernel vs the long-double golden values as a function of the
tile length / low-column count (not a test)."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import superman_b200 as sp
import _golden
c = _golden.corpus()
for name in sorted(c):
    e = c[name]
    A = _golden.dense_from(e)
    n = e["n"]
    row = []
    for B in (3, 4):
        os.environ["SP_DENSE_LOWCOLS"] = str(B)
        for tl in (6, 8, 10, 12, 13, 14):
            os.environ["SP_DENSE_TILE_LOG2"] = str(tl)
            v = sp.dense_ryser(A, n, 4)
            vt = sp.dense_ryser(A.T.copy(), n, 4)
            row.append("B%d c%-2d %+.2e %+.2e" % (B, tl, v / e["ld"] - 1, vt / e["ld"] - 1))
    print(name, "n=%d" % n)
    for r in row: print("   ", r)
os.environ.pop("SP_DENSE_TILE_LOG2"); os.environ.pop("SP_DENSE_LOWCOLS")
import bench
A = bench.synthetic_matrix(36, 0.5)
rng = np.random.default_rng(0)
vals = []
for tl in (10, 12, 13):
    os.environ["SP_DENSE_TILE_LOG2"] = str(tl)
    v = [sp.dense_ryser(A, 36, 4), sp.dense_ryser(A.T.copy(), 36, 4)]
    for k in range(3):
        p, q = rng.permutation(36), rng.permutation(36)
        v.append(sp.dense_ryser(A[p][:, q].copy(), 36, 4))
    m = np.mean(v)
    print("n=36 c=%d" % tl, " ".join("%+.2e" % (x / m - 1) for x in v), "mean %.15e" % m)
