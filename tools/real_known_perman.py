"""Evidence: the real-world matrices with recorded permanents (tests/golden/known_perman.json) on the
GPU engine -- direct and through -o -- with wall times, next to the reference kit's recorded values
and CPU seconds."""
import os, sys, time, json
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import _golden
import superman_b200 as sp
d = _golden.known_perman()
g = sp.device_count()
out = []
def timed(label, fn):
    fn()
    t = time.perf_counter(); v = fn(); dt = time.perf_counter() - t
    return label, v, dt
for name in ("chesapeake", "will57"):
    e = d[name]; a = _golden.dense_from(e); n = e["n"]
    runs = []
    st = sp.SpStats()
    if n <= 40:
        m1 = sp.Matrix.from_dense(a).compress(1)
        runs.append(timed("SpaRyser + SortOrder, direct (2^%d indices)" % (n - 1), lambda: sp.sparse_ryser(m1.mat, m1.cptrs, m1.rows, m1.cvals, n, 4)))
        m2 = sp.Matrix.from_dense(a).compress(2)
        runs.append(timed("SkipPer + SkipOrder, direct", lambda: sp.skipper(m2.mat, m2.rptrs, m2.cols, m2.cptrs, m2.rows, m2.cvals, n, 7)))
        runs.append(timed("dense, direct", lambda: sp.dense_ryser(a, n, 4)))
    runs.append(timed("-o, SpaRyser leaves", lambda: sp.permanent_compressed(a, sparse=True, preprocessing=1, algo_id=4, stats=st)))
    leaves = st.chunks
    runs.append(timed("-o, dense leaves", lambda: sp.permanent_compressed(a, sparse=False, algo_id=4)))
    if g > 1:
        runs.append(timed("-o, SpaRyser leaves, %d GPUs" % g, lambda: sp.permanent_compressed(a, sparse=True, preprocessing=1, algo_id=5, gpu_num=g)))
    print("== %s: n = %d, %d entries, %d leaves with -o" % (name, n, int((a != 0).sum()), leaves))
    for label, v, dt in runs:
        print("   %-46s %.15e   %.3f s" % (label, v, dt))
    for log, r in sorted(e["recorded"].items()):
        print("   recorded %-37s %s   %.0f s (CPU, reference kit)" % (log, r["perman"], r["seconds"]))
    if "ld_recursion" in e:
        print("   CPU long-double recursion (leaves <= 23)        %.15e" % e["ld_recursion"])
    sys.stdout.flush()
