"""Times the UNMODIFIED reference GPU wrappers (oracle/_ref/libref_gpu.so, built from the reference's
own .cu files for sm_100a) on this box, next to libsuperman_b200 on the same inputs.  Reporting
tool only ("the kernel to beat"); not part of the product or the tests."""
import ctypes as C, os, sys, time, json
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
import bench
import superman_b200 as sp
from superman_b200._ffi import SpStats

so = os.path.join(R, "oracle", "_ref", "libref_gpu.so")
lib = C.CDLL(so)
dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
lib.ref_gpu_dense_multigpu.restype = C.c_double
lib.ref_gpu_dense_multigpu.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_int]
lib.ref_gpu_dense_chunks.restype = C.c_double
lib.ref_gpu_dense_chunks.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_int]
lib.ref_gpu_sparse_multigpu.restype = C.c_double
lib.ref_gpu_sparse_multigpu.argtypes = [dp, ip, ip, dp, C.c_int, C.c_int, C.c_int, C.c_int]
ngpu = sp.device_count()
out = []
for n in (32, 36):
    A = np.ascontiguousarray(bench.synthetic_matrix(n, 0.5))
    for g in sorted({1, min(2, ngpu), ngpu}):
        lib.ref_gpu_dense_multigpu(A.ctypes.data_as(dp), n, g, 2048, 128)          # warm-up (context, allocs)
        t = time.perf_counter(); ref = lib.ref_gpu_dense_multigpu(A.ctypes.data_as(dp), n, g, 2048, 128); tr = time.perf_counter() - t
        st = SpStats()
        sp.dense_ryser(A, n, 5, gpu_num=g, stats=st)
        t = time.perf_counter(); ours = sp.dense_ryser(A, n, 5, gpu_num=g, stats=st); to = time.perf_counter() - t
        out.append(dict(case="dense -p5", n=n, gpus=g, ref_s=tr, ours_s=to, speedup=tr / to, ref_value=ref, ours_value=ours,
                        rel_diff=abs(ref / ours - 1)))
        print(out[-1], flush=True)
n = 33
rng = np.random.default_rng(33000)
A = (rng.random((n, n)) < 0.2) * rng.integers(1, 6, (n, n)).astype(float)
A[np.arange(n), rng.permutation(n)] = 1.0
m = sp.Matrix.from_dense(A).compress(1)
mat, cp, ro, cv = np.ascontiguousarray(m.mat), m.cptrs, m.rows, m.cvals
lib.ref_gpu_sparse_multigpu(mat.ctypes.data_as(dp), cp.ctypes.data_as(ip), ro.ctypes.data_as(ip), cv.ctypes.data_as(dp), n, 1, 2048, 256)
t = time.perf_counter(); ref = lib.ref_gpu_sparse_multigpu(mat.ctypes.data_as(dp), cp.ctypes.data_as(ip), ro.ctypes.data_as(ip), cv.ctypes.data_as(dp), n, 1, 2048, 256); tr = time.perf_counter() - t
sp.sparse_ryser(m.mat, cp, ro, cv, n, 4)
t = time.perf_counter(); ours = sp.sparse_ryser(m.mat, cp, ro, cv, n, 4); to = time.perf_counter() - t
out.append(dict(case="SpaRyser -s -p5 -r1 (int, p=0.2)", n=n, gpus=1, ref_s=tr, ours_s=to, speedup=tr / to, ref_value=ref, ours_value=ours, rel_diff=abs(ref / ours - 1)))
print(out[-1], flush=True)
# config 5: estimators on grid graphs, 2^20 trials (one reference launch)
ip = C.POINTER(C.c_int)
lib.ref_gpu_rasmussen_chunks_sparse.restype = C.c_double
lib.ref_gpu_rasmussen_chunks_sparse.argtypes = [ip, ip, ip, ip, C.c_int, C.c_int, C.c_int, C.c_int]
lib.ref_gpu_scaling_chunks_sparse.restype = C.c_double
lib.ref_gpu_scaling_chunks_sparse.argtypes = [ip, ip, ip, ip, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
for gm, gn in ((8, 8), (36, 36)):
    g = sp.Matrix.grid(gm, gn)
    cp, ro, rp, co = g.cptrs, g.rows, g.rptrs, g.cols
    args = (cp.ctypes.data_as(ip), ro.ctypes.data_as(ip), rp.ctypes.data_as(ip), co.ctypes.data_as(ip), g.nov, g.nnz, 1 << 20, 1)
    trials = 1 << 20
    for name, rf, of in (("Rasmussen", lambda: lib.ref_gpu_rasmussen_chunks_sparse(*args),
                          lambda: sp.rasmussen_sparse(rp, co, cp, ro, g.nov, g.nnz, trials, 1, seed=1)),
                         ("scaling y4 z5", lambda: lib.ref_gpu_scaling_chunks_sparse(*args, 4, 5),
                          lambda: sp.scaling_sparse(cp, ro, rp, co, g.nov, g.nnz, trials, 4, 5, 1, seed=1))):
        rf()
        t = time.perf_counter(); ref = rf(); tr = time.perf_counter() - t
        of()
        t = time.perf_counter(); ours = of(); to = time.perf_counter() - t
        row = dict(case="%s %dx%d grid, 2^20 trials" % (name, gm, gn), gpus=1, ref_s=tr, ours_s=to, speedup=tr / to, ref_value=ref, ours_value=ours)
        if name.startswith("scaling"):
            row["note"] = ("NOT a baseline: the reference's scaled kernel reads `is_break` uninitialised "
                           "(gpu_approximation_sparse.cu:342-449), leaves its trial loop on the first step and returns 0; "
                           "its time is that of an empty kernel")
        out.append(row)
        print(out[-1], flush=True)
json.dump(out, open(os.path.join(R, "gpurun_out", "ref_gpu_time.json"), "w"), indent=1)
