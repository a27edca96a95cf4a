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
json.dump(out, open(os.path.join(R, "gpurun_out", "ref_gpu_time.json"), "w"), indent=1)
