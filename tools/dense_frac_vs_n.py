"""Roofline fraction of the dense Ryser kernel for every matrix order it is compiled for (n = 13..64):
2^33 Gray indices (the whole space when n <= 34) through the resident handle, CUDA-event time of the
launches, against the FP64 issue peak measured in the same process (spd_fp64_peak_instr_per_s).
Algorithmic work: 2n+1 FP64 instructions per Gray index."""
import json, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import bench
import superman_b200 as sp

peak = max(sp.fp64_peak(0, 200) for _ in range(3))
print("# measured FP64 issue peak %.4e thread-instr/s" % peak)
rows = []
NS = [int(x) for x in os.environ["NS"].split(",")] if os.environ.get("NS") else range(13, 65)
for n in NS:
    A = bench.synthetic_matrix(n, 0.5)
    hi = min(1 << (n - 1), 1 << 33)
    with sp.DenseHandle(A, n) as h:
        st = sp.SpStats()
        h.run(0, hi, st)
        best = 1e30
        for _ in range(3):
            h.run(0, hi, st)
            best = min(best, st.kernel_ms)
    its = hi / (best * 1e-3)
    frac = its * (2 * n + 1) / peak
    rows.append(dict(n=n, indices=hi, kernel_ms=best, it_per_s=its, frac=frac, path=st.path, tile_log2=st.tile_log2))
    print("n=%2d  indices 2^%d  %.3f ms  %.4e it/s  frac %.3f  tile 2^%d" % (n, hi.bit_length() - 1, best, its, frac, st.tile_log2), flush=True)
json.dump(dict(peak=peak, rows=rows), open(os.path.join(R, "gpurun_out", "dense_frac_vs_n.json"), "w"), indent=1)
