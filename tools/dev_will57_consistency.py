"""dev: is the -o value of will57 (n = 57) independent of the reduction tree?  transpose, random row /
column permutations and different leaf sizes take different d1 / d2 / d34 steps."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import _golden
import superman_b200 as sp
for name in ("chesapeake", "will57"):
    e = _golden.known_perman()[name]
    a = _golden.dense_from(e); n = e["n"]
    rng = np.random.default_rng(5)
    variants = [("as read", a), ("transposed", a.T.copy())]
    for k in range(3):
        variants.append(("rows/cols permuted #%d" % k, a[rng.permutation(n)][:, rng.permutation(n)].copy()))
    for label, m in variants:
        for leaf in (30, 33):
            st = sp.SpStats()
            t = time.perf_counter()
            v = sp.permanent_compressed(m, sparse=True, preprocessing=1, algo_id=4, leaf_nov=leaf, stats=st)
            print("%-10s %-24s leaf_nov %d: %.15e  %5d leaves  %.2f s" % (name, label, leaf, v, st.chunks, time.perf_counter() - t), flush=True)
