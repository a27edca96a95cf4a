"""dev: B = 3 vs B = 4 low-column count of the dense kernel at the orders where the default dips"""
import os, subprocess, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys
sys.path.insert(0, %r)
import bench, superman_b200 as sp
peak = max(sp.fp64_peak(0, 100) for _ in range(2))
for n in (33, 40, 41, 44, 55, 58, 62, 63):
    A = bench.synthetic_matrix(n, 0.5)
    hi = min(1 << (n - 1), 1 << 33)
    with sp.DenseHandle(A, n) as h:
        st = sp.SpStats(); h.run(0, hi, st); best = 1e30
        for _ in range(3):
            h.run(0, hi, st); best = min(best, st.kernel_ms)
    print("B=%%s n=%%d %%.3f ms frac %%.3f" %% (os.environ.get("SP_DENSE_LOWCOLS", "4"), n, best, hi / (best * 1e-3) * (2 * n + 1) / peak), flush=True)
''' % R
for B in ("4", "3"):
    print(subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, SP_DENSE_LOWCOLS=B), capture_output=True, text=True).stdout)
