"""One short dense run per order in NS (default 32,33) for ncu: 592*8 groups of 2^16 indices each."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, superman_b200 as sp
for n in [int(x) for x in os.environ.get("NS", "32,33").split(",")]:
    A = bench.synthetic_matrix(n, 0.5)
    hi = min(1 << (n - 1), int(os.environ.get("GROUPS", 592 * 8)) << 16)
    with sp.DenseHandle(A, n) as h:
        st = sp.SpStats()
        h.run(0, hi, st)
        print(n, hi, st.kernel_ms, flush=True)
