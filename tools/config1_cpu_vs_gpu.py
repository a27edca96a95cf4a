"""BASELINE config 1 (`perman -f double/30_0.50_0 -c -t<nproc>`): the reference's own CPU path
parallel_perman64 (algo.h:662, unmodified, oracle/_ref/libref.so) on this box's host cores, next to
the GPU engine on the same matrices (golden fixtures of the reference corpus; nothing is read from
the reference tree at run time).  SURVEY 8(d) 'CPU baseline beside it'."""
import json, os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import _golden
from _oracle import Reference
import superman_b200 as sp

ref = Reference()
threads = len(os.sched_getaffinity(0))
out = []
for name in ("int/30_0.50_0", "double/30_0.50_0"):
    e = _golden.corpus()[name]
    a = _golden.dense_from(e); n = e["n"]
    ref.parallel_perman64(a, threads)                      # warm the OpenMP pool
    t0 = time.perf_counter(); v_cpu = ref.parallel_perman64(a, threads); t_cpu = time.perf_counter() - t0
    st = sp.SpStats()
    sp.dense_ryser(a, algo_id=4, stats=st)
    t0 = time.perf_counter(); v_gpu = sp.dense_ryser(a, algo_id=4, stats=st); t_gpu = time.perf_counter() - t0
    its = float(1 << (n - 1))
    out.append({"file": name, "n": n, "host_threads": threads,
                "ref_cpu_parallel_perman64_s": t_cpu, "ref_cpu_it_per_s": its / t_cpu, "ref_cpu_value": v_cpu,
                "ref_cpu_rel_err_vs_long_double": v_cpu / e["ld"] - 1,
                "gpu_wall_s": t_gpu, "gpu_kernel_ms": st.kernel_ms, "gpu_it_per_s": its / t_gpu, "gpu_value": v_gpu,
                "gpu_rel_err_vs_long_double": v_gpu / e["ld"] - 1, "speedup_wall": t_cpu / t_gpu})
    print(json.dumps(out[-1]), flush=True)
os.makedirs(os.path.join(R, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(R, "gpurun_out", "config1_cpu_vs_gpu.json"), "w"), indent=1)
