"""dev: LevelRyser parameter sweep (B, S0, S through the SP_* environment knobs) on the config-3 matrix,
next to the cost model's own choice.  Each configuration runs in a fresh process (the knobs are read
at plan creation)."""
import os, subprocess, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np
import superman_b200 as sp
n = 33
rng = np.random.default_rng(33000)
A = (rng.random((n, n)) < 0.2) * rng.integers(1, 6, (n, n)).astype(float)
A[np.arange(n), rng.permutation(n)] = 1.0
st = sp.SpStats()
m = sp.Matrix.from_dense(A).compress(int(os.environ.get("PRE", "1")))
best = 1e9
for _ in range(3):
    try:
        v = sp.sparse_ryser(m.mat, m.cptrs, m.rows, m.cvals, n, 4, stats=st)
    except Exception as e:
        print("ERR", str(e)[:80]); sys.exit(0)
    best = min(best, st.kernel_ms)
print("%%.3f ms  %%.12e" %% (best, v))
''' % R
for B in (0, 3, 4):
    for S in (0, 3, 4, 6):
        for S0 in (0, 1, 2, 3, 4):
            if (B == 0) != (S == 0) or (B == 0) != (S0 == 0) or S0 > S or S0 < S - 2:
                continue
            env = dict(os.environ)
            if B:
                env.update(SP_SPARSE_LOWCOLS=str(B), SP_LEVEL_SLOTS=str(S), SP_LEVEL_SLOTS0=str(S0), SP_SPARSE_ENGINE="2")
            out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True).stdout.strip()
            print("B=%d S0=%d S=%d: %s" % (B, S0, S, out), flush=True)
