"""dev: direct vs compressed (-o) run on the sparse files of the reference corpus (golden fixtures)"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import _golden
import superman_b200 as sp
for name, e in _golden.corpus().items():
    if "_0.20_" not in name and "_0.30_" not in name:
        continue
    a = _golden.dense_from(e); n = e["n"]
    m = sp.Matrix.from_dense(a).compress(1)
    st = sp.SpStats()
    for _ in range(2):
        d = sp.sparse_ryser(m.mat, m.cptrs, m.rows, m.cvals, n, 4, stats=st)
    t_direct, w_direct = st.kernel_ms, st.wall_ms
    for _ in range(2):
        c = sp.permanent_compressed(a, sparse=True, preprocessing=1, algo_id=4, stats=st)
    print("%-18s n=%d direct %.3f ms (wall %.3f) rel %.1e | -o: %d leaf(es), %.2e indices, kernel %.3f ms wall %.3f ms rel %.1e"
          % (name, n, t_direct, w_direct, d / e["ld"] - 1, st.chunks, st.units, st.kernel_ms, st.wall_ms, c / e["ld"] - 1))
