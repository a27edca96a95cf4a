"""In-process multi-GPU probe (-p5 static / -p6 dynamic) for dense n=36/40, SpaRyser/SkipPer n=36, approximations."""
import os, sys, time, json
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
import bench
import superman_b200 as sp
from superman_b200._ffi import SpStats
ng = sp.device_count()
res = []
for n in (36, 40):
    A = bench.synthetic_matrix(n, 0.5)
    base = None
    for g in [x for x in (1, 2, 4, 8) if x <= ng]:
        for algo, name in ((5, "static"), (6, "dynamic")):
            if n == 40 and g < ng and not (g == 1 and algo == 5):
                continue
            st = SpStats()
            if not (n == 40 and g == 1):
                sp.dense_ryser(A, n, algo, gpu_num=g, stats=st)   # warm-up (contexts)
            t = time.perf_counter(); v = sp.dense_ryser(A, n, algo, gpu_num=g, stats=st); dt = time.perf_counter() - t
            its = (1 << (n - 1)) / dt
            if base is None: base = its
            r = dict(case="dense", n=n, gpus=g, partition=name, seconds=dt, kernel_ms_max=st.kernel_ms, it_per_s=its, speedup=its / base,
                     chunks=st.chunks, value=v, dev_ms=[round(x, 2) for x in st.device_ms[:g]])
            res.append(r); print(r, flush=True)
n = 36
rng = np.random.default_rng(36000)
A = (rng.random((n, n)) < 0.2) * rng.integers(1, 6, (n, n)).astype(float)
A[np.arange(n), rng.permutation(n)] = 1.0
for pre, fn in ((1, "spa"), (2, "skip")):
    m = sp.Matrix.from_dense(A).compress(pre)
    for g in [x for x in (1, 8) if x <= ng]:
        st = SpStats()
        call = (lambda: sp.sparse_ryser(m.mat, m.cptrs, m.rows, m.cvals, n, 6, gpu_num=g, stats=st)) if fn == "spa" else \
               (lambda: sp.skipper(m.mat, m.rptrs, m.cols, m.cptrs, m.rows, m.cvals, n, 8, gpu_num=g, stats=st))
        call()
        t = time.perf_counter(); v = call(); dt = time.perf_counter() - t
        r = dict(case=fn, n=n, gpus=g, seconds=dt, visited=st.visited, units=st.units, value=v, chunks=st.chunks)
        res.append(r); print(r, flush=True)
g36 = sp.Matrix.grid(36, 36)
for g in [x for x in (1, 8) if x <= ng]:
    st = SpStats()
    sp.scaling_sparse(g36.cptrs, g36.rows, g36.rptrs, g36.cols, g36.nov, g36.nnz, 8 << 20, 4, 5, g, seed=1, stats=st)
    t = time.perf_counter(); v = sp.scaling_sparse(g36.cptrs, g36.rows, g36.rptrs, g36.cols, g36.nov, g36.nnz, 8 << 20, 4, 5, g, seed=1, stats=st); dt = time.perf_counter() - t
    r = dict(case="scaling 36x36 x8Mi y4 z5", gpus=g, seconds=dt, trials_per_s=st.units / dt, value=v, std_error=st.std_error)
    res.append(r); print(r, flush=True)
json.dump(res, open(os.path.join(R, "gpurun_out", "multi_gpu_probe.json"), "w"), indent=1)
