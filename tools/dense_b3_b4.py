"""Dense register kernel with 3 vs 4 low columns (8 vs 16 running products per thread) for a list of
orders: kernel time and fraction of the measured FP64 issue peak.  NS=... selects the orders."""
import os, subprocess, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys
sys.path.insert(0, %r)
import bench, superman_b200 as sp
peak = max(sp.fp64_peak(0, 100) for _ in range(2))
for n in [int(x) for x in os.environ["NS"].split(",")]:
    A = bench.synthetic_matrix(n, 0.5)
    hi = min(1 << (n - 1), 1 << 33)
    with sp.DenseHandle(A, n) as h:
        st = sp.SpStats()
        h.run(0, hi, st)
        best = min(h.run(0, hi, st) and 0 or st.kernel_ms for _ in range(3))
    print("B=%%s n=%%d %%.3f ms frac %%.4f" %% (os.environ.get("SP_DENSE_LOWCOLS", "auto"), n, best, hi / (best * 1e-3) * (2 * n + 1) / peak), flush=True)
''' % R
for B in ("3", "4"):
    print(subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, SP_DENSE_LOWCOLS=B), capture_output=True, text=True).stdout)
