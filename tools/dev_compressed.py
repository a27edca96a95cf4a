"""dev: sp_permanent_compressed variants on the n=40 banded matrix against the CPU recursion"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import superman_b200 as sp
from _oracle import Oracle
from _compressed import oracle_compressed, banded

o = Oracle()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
a = banded(np.random.default_rng(100 + n), n, "real")
want = oracle_compressed(sp, o, a)
for sparse, algo, pre in [(False, 4, 0), (True, 4, 0), (True, 4, 1), (True, 4, 2), (True, 7, 2)]:
    for leaf in (20, 25, 30, 33, 35):
        for thr in (0.0, -1.0):
            st = sp.SpStats()
            got = sp.permanent_compressed(a, sparse=sparse, preprocessing=pre, algo_id=algo, leaf_nov=leaf,
                                          scaling_threshold=thr, stats=st)
            print(f"sparse={sparse} algo={algo} pre={pre} leaf={leaf} thr={thr}: rel {got / want - 1:+.3e} leaves {st.chunks} "
                  f"kernel {st.kernel_ms:.2f} ms wall {st.wall_ms:.2f} ms", flush=True)
