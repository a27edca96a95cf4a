"""dev: compare LevelRyser kernel variants built side by side by tools/build_level_variants.sh.  Every variant is
loaded in its own process (SUPERMAN_B200_LIB); SpaRyser + SortOrder and SkipPer + SkipOrder on the development
matrix and on bench.py's three config-3 matrices: best kernel ms of 5 and the value (must agree between variants)."""
import os, subprocess, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np
import superman_b200 as sp
import bench
n = 33
rng = np.random.default_rng(33000)
A = (rng.random((n, n)) < 0.2) * rng.integers(1, 6, (n, n)).astype(float)
A[np.arange(n), rng.permutation(n)] = 1.0
mats = {"dev": A, "dev01": (A != 0).astype(float)}
for k in ("bin", "int", "dbl"): mats[k] = bench.config3_matrix(k)
extra = [int(x) for x in os.environ.get("EXTRA_N", "").split(",") if x]
for nn in extra:
    r2 = np.random.default_rng(nn)
    M = (r2.random((nn, nn)) < 0.2) * r2.integers(1, 6, (nn, nn)).astype(float)
    M[np.arange(nn), r2.permutation(nn)] = 1.0
    mats["n%%d" %% nn] = M
st = sp.SpStats()
out = []
for name, M in mats.items():
    nn = M.shape[0]
    for label, pre, skip in (("spa", 1, False), ("skip", 2, True)):
        m = sp.Matrix.from_dense(M).compress(pre)
        best = 1e9
        for _ in range(5):
            if skip: v = sp.skipper(m.mat, m.rptrs, m.cols, m.cptrs, m.rows, m.cvals, nn, 7, stats=st)
            else: v = sp.sparse_ryser(m.mat, m.cptrs, m.rows, m.cvals, nn, 4, stats=st)
            best = min(best, st.kernel_ms)
        out.append("%%s/%%s %%.3f ms %%.13e" %% (name, label, best, v))
print("\n".join(out))
''' % R
# arguments: variant[:ENV=VALUE[:ENV=VALUE...]]
names = sys.argv[1:] or sorted(os.listdir(os.path.join(R, "tools", "_bin")))
for spec in names:
    name, *kv = spec.split(":")
    lib = os.path.join(R, "tools", "_bin", name, "libsuperman_b200.so")
    if name == "main":
        lib = os.path.join(R, "superman_b200", "libsuperman_b200.so")
    if not os.path.exists(lib):
        continue
    env = dict(os.environ, SUPERMAN_B200_LIB=lib)
    env.update(dict(x.split("=", 1) for x in kv))
    name = spec
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print("== %s" % name)
    print(r.stdout.strip() or r.stderr[-2000:], flush=True)
