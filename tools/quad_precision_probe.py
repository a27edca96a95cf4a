"""-q (double-double) dense mode next to FP64: time and accuracy on the synthetic n = 30 / 36 workloads (long-double
goldens) and on chesapeake (exact integer known).  Evidence for DESIGN.md 4.6; not a test."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import superman_b200 as sp, bench, _golden
from superman_b200._ffi import SpStats
st = SpStats()
peak = sp.fp64_peak(0, 100)
for n in (30, 32, 36):
    A = bench.synthetic_matrix(n, 0.5)
    v64 = sp.dense_ryser(A, n, 4, stats=st); ms64 = st.kernel_ms
    sp.set_precision(True)
    vq = sp.dense_ryser(A, n, 4, stats=st); msq = st.kernel_ms
    sp.set_precision(False)
    its = (1 << (n - 1)) / (msq * 1e-3)
    print("n=%d  fp64 %.3f ms rel.err %s | quad %.2f ms (%.1fx) rel.err %s  FP64 instr/index %d -> %.3f of the FP64 issue peak"
          % (n, ms64, bench.golden_rel_err(n, v64), msq, msq / ms64, bench.golden_rel_err(n, vq), 15 * n + 11, its * (15 * n + 11) / peak), flush=True)
e = _golden.known_perman()["chesapeake"]
a = _golden.dense_from(e)
exact = int(e["exact"])
v = sp.dense_ryser(a, 39, 4, stats=st)
print("chesapeake fp64  %.3f  rel.err %.2e  %.1f ms" % (v, abs(v / exact - 1), st.kernel_ms))
sp.set_precision(True)
v = sp.dense_ryser(a, 39, 4, stats=st)
sp.set_precision(False)
print("chesapeake quad  %.3f  rel.err %.2e  %.1f ms   (exact %d)" % (v, abs(v / exact - 1), st.kernel_ms, exact))
