"""GPU dev probe: dense kernel correctness on small n + throughput sweep at n=36 (not a test)."""
import os, sys, time, json, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import superman_b200 as sp
from superman_b200._ffi import SpStats

def perm_dp(A):
    n = len(A)
    dp = {0: 0}
    dp = [0] * (1 << n); dp[0] = 1
    for mask in range(1, 1 << n):
        r = bin(mask).count("1") - 1
        s = 0; m = mask
        while m:
            j = (m & -m).bit_length() - 1
            s += dp[mask ^ (1 << j)] * A[r][j]
            m &= m - 1
        dp[mask] = s
    return dp[-1]

rng = np.random.default_rng(1)
print("devices", sp.device_count())
peak = sp.fp64_peak(0, 300)
print("fp64 peak instr/s %.4e (nominal 148*64*1.965e9=%.4e)" % (peak, 148*64*1.965e9))
for n in (2, 3, 5, 7, 8, 10, 12, 14, 16):
    A = (rng.random((n, n)) < 0.6) * rng.integers(1, 6, (n, n))
    want = perm_dp(A.tolist())
    for B in (3, 4):
        os.environ["SP_DENSE_LOWCOLS"] = str(B)
        got = sp.dense_ryser(A.astype(float), n, 4)
        os.environ["SP_DENSE_FORCE_SMEM"] = "1"
        got2 = sp.dense_ryser(A.astype(float), n, 4)
        del os.environ["SP_DENSE_FORCE_SMEM"]
        # ragged range through both kernels
        full = 1 << (n - 1)
        a, b = full // 3 + 1, full - full // 5
        parts = sp.dense_ryser_range(A.astype(float), 0, a, n) + sp.dense_ryser_range(A.astype(float), a, b, n) + sp.dense_ryser_range(A.astype(float), b, full, n)
        got3 = parts * sp.nw_factor(n)
        ok = all(abs(g - want) <= 1e-9 * max(1, abs(want)) for g in (got, got2, got3))
        print(n, B, want, got, got2, got3, "OK" if ok else "MISMATCH")

n = 36
A = (rng.random((n, n)) < 0.5) * rng.uniform(0.01, 5, (n, n))
res = {}
for B, tl in itertools.product((3, 4), (20, 21, 22, 23)):
    os.environ["SP_DENSE_LOWCOLS"] = str(B)
    os.environ["SP_DENSE_TILES_LOG2"] = str(tl)
    with sp.DenseHandle(A, n) as h:
        st = SpStats()
        lo, hi = 0, 1 << 33   # quarter of the n=36 space
        v = h.run(lo, hi, st)
        v = h.run(lo, hi, st)
        its = (hi - lo) / (st.kernel_ms * 1e-3)
        print("n=36 B=%d tiles_log2=%d c=%d ms=%.2f it/s=%.4e frac_of_measured=%.3f frac_nominal=%.3f val=%.15e" % (
            B, tl, st.tile_log2, st.kernel_ms, its, its * (2*n+1) / peak, its*(2*n+1)/(148*64*1.965e9), v))
for n in (32, 40):
    A = (rng.random((n, n)) < 0.5) * rng.uniform(0.01, 5, (n, n))
    for B in (3, 4):
        os.environ["SP_DENSE_LOWCOLS"] = str(B)
        os.environ["SP_DENSE_TILES_LOG2"] = "22"
        with sp.DenseHandle(A, n) as h:
            st = SpStats()
            hi = 1 << min(n - 1, 33)
            v = h.run(0, hi, st); v = h.run(0, hi, st)
            its = hi / (st.kernel_ms * 1e-3)
            print("n=%d B=%d c=%d ms=%.2f it/s=%.4e frac_of_measured=%.3f" % (n, B, st.tile_log2, st.kernel_ms, its, its*(2*n+1)/peak))
# full n=36 e2e
n = 36
A = (rng.random((n, n)) < 0.5) * rng.uniform(0.01, 5, (n, n))
os.environ["SP_DENSE_LOWCOLS"] = "4"
st = SpStats()
t = time.time(); v = sp.dense_ryser(A, n, 4, stats=st); t = time.time() - t
print("full n=36: perm=%.15e wall=%.3fs kernel_ms=%.2f it/s=%.4e" % (v, t, st.kernel_ms, (1 << 35) / (st.kernel_ms * 1e-3)))
