"""stress: the leaf-parallel -o driver (two host threads per device launching the same kernels with
different shared-memory sizes) repeated on will57; every run must return the same value"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import _golden
import superman_b200 as sp
a = _golden.dense_from(_golden.known_perman()["will57"])
g = sp.device_count()
vals = set()
for rep in range(int(os.environ.get("REPS", "6"))):
    for sparse, algo, pre, leaf in ((True, 4, 1, 30), (True, 7, 2, 28), (False, 4, 0, 27), (True, 5 if g > 1 else 4, 1, 26)):
        v = sp.permanent_compressed(a, sparse=sparse, preprocessing=pre, algo_id=algo, gpu_num=g, leaf_nov=leaf)
        vals.add(round(v / 1.070536592880585e18, 10))
print("runs ok; distinct normalised values:", sorted(vals))
