#!/usr/bin/env python
"""bench.py -- headline benchmark: Gray-code iterations/s of the dense FP64 Ryser path at n=36.

    python bench.py --gpus N --steps K --warmup W            (N > 1: under torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

One "step" is one complete permanent of the seeded synthetic 36x36 density-0.50 double matrix
(BASELINE.json configs[3]; 2^35 Gray indices).  With N ranks the index space is cut into N
contiguous, 2^16-aligned slices, one per rank / GPU; there is no data-path collective -- each rank
leaves one double and rank 0 adds them in rank order (strong scaling of one permanent, as the
north star asks: "time-to-permanent ... at 1/2/4/8 B200").

  value      iterations/s with the matrix already resident in HBM: (K * 2^35) / sum of the
             per-step device times, the device time being CUDA events recorded by the library on
             the stream its kernels run on; max over ranks.
  e2e        the same metric through the C-ABI entry point a caller uses
             (sp_dense_ryser_range: host matrix in, one double out), host wall clock around the
             call, H2D of matrix + start vector and D2H of the result inside; max over ranks.
  roofline   FP64-issue roofline of the dominant kernel (ryser_reg_kernel): algorithmic work is
             2n+1 FP64 instructions per Gray index; peak is the DFMA issue rate
             measured on this GPU in this run (MEASURED_PEAKS.json has no FP64 entry).
  cpu_baseline  the reference's own OpenMP all-double range kernel cpu_perman64
             (gpu_exact_dense.cu:6-69, compiled unmodified into oracle/_ref/libref_gpu.so) on a
             bounded slice of the same workload, all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_DENSE = 36
DENSITY = 0.50
ALIGN_LOG2 = 16


def synthetic_matrix(n: int, density: float, instance: int = 0):
    """Seeded generator matching the reference corpus: entry non-zero with
    probability `density`, value uniform in (0, 5) rounded to 6 decimals; seed = 1000*n + instance."""
    import numpy as np
    rng = np.random.default_rng(1000 * n + instance)
    pat = rng.random((n, n)) < density
    val = np.round(rng.uniform(0.0, 5.0, (n, n)), 6)
    val[val == 0.0] = 1e-6
    return (pat * val).astype(np.float64)


def golden_rel_err(n: int, value: float):
    """relative difference to the committed long-double oracle value of this very workload
    (tests/golden/bench<n>.json, a number in a fixture file -- nothing of oracle/ runs here)"""
    p = os.path.join(ROOT, "tests", "golden", "bench%d.json" % n)
    if not os.path.exists(p):
        return None
    with open(p) as f:
        g = json.load(f)
    return abs(value / g["ld"] - 1.0)


def rank_slice(total: int, rank: int, world: int, align_log2: int = ALIGN_LOG2):
    """Contiguous slice of [0, total) for `rank`, boundaries rounded down to 2^align_log2
    (same rule as sp_sched_boundary in superman_b200/host/sp_sched.c)."""
    def b(i):
        if i <= 0:
            return 0
        if i >= world:
            return total
        q, r = divmod(total, world)
        v = q * i + min(i, r)
        return (v >> align_log2) << align_log2
    return b(rank), b(rank + 1)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.time(), line.strip()))
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self, t0: float, t1: float) -> dict:
        sm, smax, reasons, power = [], [], set(), []
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.2:
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1])); smax.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power),
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_rate(mat, n: int, seconds: float, threads: int | None = None):
    """Times the reference's cpu_perman64 (all-double OpenMP range kernel) on a leading slice of
    the n x n workload sized for about `seconds` of work.  Returns (iters_per_s, cores, sample, kind)."""
    import ctypes as C
    import numpy as np
    so = os.path.join(ROOT, "oracle", "_ref", "libref_gpu.so")
    a = np.ascontiguousarray(mat, dtype=np.float64)
    if os.path.exists(so):
        lib = C.CDLL(so)
        lib.ref_cpu_perman64.restype = C.c_double
        lib.ref_cpu_perman64.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_longlong, C.c_longlong, C.c_int]
        lib.ref_gpu_max_threads.restype = C.c_int
        # all host cores this process may use; torchrun exports OMP_NUM_THREADS=1, which would otherwise
        # pin the reference's OpenMP region to one thread (cpu_perman64 takes the count explicitly)
        cores = threads or len(os.sched_getaffinity(0)) or lib.ref_gpu_max_threads()

        def run(count):
            t = time.perf_counter()
            lib.ref_cpu_perman64(a.ctypes.data_as(C.POINTER(C.c_double)), n, 1, 1 + count, cores)
            return time.perf_counter() - t
        kind = "reference"
        what = "cpu_perman64 (gpu_exact_dense.cu:6-69, unmodified, oracle/_ref/libref_gpu.so)"
    else:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from _oracle import Oracle   # the C restatement: the one other place bench may run oracle/
        orc = Oracle()
        cores = 1

        def run(count):
            t = time.perf_counter()
            orc.ryser_range_f64(a, 1, 1 + count)
            return time.perf_counter() - t
        kind = "port"
        what = "oracle/oracle.c orc_ryser_range_f64 (serial restatement of cpu_perman64)"
    probe = 1 << 24
    dt = run(probe)
    rate = probe / max(dt, 1e-9)
    count = 1
    while count * 2 <= rate * seconds and count * 2 <= (1 << (n - 1)) - 1:
        count *= 2
    count = max(count, probe)
    dt = run(count)
    sample = f"Gray indices [1, 1+2^{count.bit_length() - 1}) of the n={n} workload, {what}"
    return count / dt, cores, sample, kind, count, dt


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    mat = synthetic_matrix(N_DENSE, DENSITY)
    # size one step at ~4 s of CPU work, then time W + K steps of it
    rate, cores, sample, kind, count, _ = cpu_reference_rate(mat, N_DENSE, 4.0)
    import ctypes as C
    import numpy as np
    so = os.path.join(ROOT, "oracle", "_ref", "libref_gpu.so")
    a = np.ascontiguousarray(mat, dtype=np.float64)
    if kind == "reference":
        lib = C.CDLL(so)
        lib.ref_cpu_perman64.restype = C.c_double
        lib.ref_cpu_perman64.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_longlong, C.c_longlong, C.c_int]
        step = lambda: lib.ref_cpu_perman64(a.ctypes.data_as(C.POINTER(C.c_double)), N_DENSE, 1, 1 + count, cores)
    else:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from _oracle import Oracle
        orc = Oracle()
        step = lambda: orc.ryser_range_f64(a, 1, 1 + count)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = args.steps * count / dt
    line = {
        "impl": "reference", "metric": "gray_code_iterations_per_second", "value": value, "unit": "iterations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"dense Ryser n={N_DENSE} density {DENSITY} FP64, one full permanent (2^{N_DENSE - 1} Gray indices) per step, "
                               f"seeded synthetic matrix (seed {1000 * N_DENSE}); BASELINE.json configs[3]",
                   "sample_per_step": sample},
        "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def run_in_process_multi_gpu(args):
    """`python bench.py --gpus N` WITHOUT torchrun: the library's own static multi-GPU split
    (sp_dense_ryser id 5 = the reference's -p5 -dN: one host thread per device, rank-order host sum)."""
    import torch
    import superman_b200 as sp
    from superman_b200._ffi import SpStats
    n = args.n
    mat = synthetic_matrix(n, DENSITY)
    total = 1 << (n - 1)
    if sp.device_count() < args.gpus:
        print(f"bench.py: --gpus {args.gpus} but {sp.device_count()} device(s) visible", file=sys.stderr)
        return 2
    peak = sp.fp64_peak(0, 200)
    st = SpStats()
    for _ in range(args.warmup):
        sp.dense_ryser(mat, n, 5, gpu_num=args.gpus, stats=st)
    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.3)
    torch.cuda.synchronize()
    t0w, t0 = time.time(), time.perf_counter()
    dev_ms, launches, perm = 0.0, 0, 0.0
    for _ in range(args.steps):
        perm = sp.dense_ryser(mat, n, 5, gpu_num=args.gpus, stats=st)
        dev_ms += st.kernel_ms            # max over devices of their CUDA-event times
        launches += st.launches
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    t1w = time.time()
    time.sleep(0.2)
    sampler.stop()
    units = args.steps * total
    value = units / (dev_ms * 1e-3)
    achieved = (total / args.gpus) * (2 * n + 1) / ((dev_ms / args.steps) * 1e-3)
    line = {
        "metric": "gray_code_iterations_per_second", "value": value, "unit": "iterations/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"dense Ryser n={n} density {DENSITY} FP64, one full permanent (2^{n - 1} Gray indices) per step, "
                               f"seeded synthetic matrix (seed {1000 * n}); BASELINE.json configs[3]",
                   "partition": f"in-process static split over {args.gpus} devices (sp_dense_ryser id 5 = -p5 -d{args.gpus}), one host thread per device",
                   "l2": "working set is 10 KB per device; the kernel is FP64-issue bound"},
        "seconds_per_permanent": dev_ms / args.steps * 1e-3, "permanent": perm,
        "roofline": {"bound": "fp64_issue", "achieved": achieved / 1e9, "peak": peak / 1e9, "unit": "Ginstr/s (thread-level FP64)",
                     "frac": achieved / peak, "traffic": None, "kernel": "spb::ryser_reg_kernel"},
        "cpu_baseline": None,
        "e2e": {"value": units / wall, "unit": "iterations/s", "ms_per_step": 1e3 * wall / args.steps,
                "h2d_bytes_per_step": 8 * (n * n + n) * args.gpus, "d2h_bytes_per_step": 8 * args.gpus,
                "api": "sp_dense_ryser(host mat, nov, 5, gpu_num, ...) -> double"},
        "gpu_launches": launches, "clocks": sampler.summary(t0w, t1w),
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=N_DENSE, help="matrix order (default 36, the headline config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3   # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        print("bench.py: no CUDA device; superman_b200 has no CPU fallback", file=sys.stderr)
        return 2
    torch.cuda.set_device(local_rank)
    if world == 1 and args.gpus > 1:
        return run_in_process_multi_gpu(args)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # rank 0 must print exactly one line: keep NCCL's version banner off stdout
        # (NCCL prints it at VERSION level and above, i.e. also at WARN)
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "NONE"
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))

    import superman_b200 as sp
    from superman_b200._ffi import SpStats

    n = args.n
    mat = synthetic_matrix(n, DENSITY)
    total = 1 << (n - 1)
    lo, hi = rank_slice(total, rank, world)
    dev = local_rank

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather_sum(x: float) -> float:
        """one double per rank, added in rank order on the host (gpu_exact_dense.cu:769-771)"""
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        parts = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        s = 0.0
        for p in parts:
            s += float(p.item())
        return s

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    # measured FP64 issue peak of this GPU (roofline denominator)
    peak = sp.fp64_peak(dev, 200)

    handle = sp.DenseHandle(mat, n, device=dev)
    st = SpStats()
    # ---- resident arm: W warm-up + K timed steps ---------------------------------------------------
    for _ in range(args.warmup):
        handle.run(lo, hi, st)
    sampler = ClockSampler(dev)
    sampler.start()
    time.sleep(0.3)
    barrier()
    t_region0 = time.time()
    w0 = time.perf_counter()
    dev_ms, launches, partial = 0.0, 0, 0.0
    for _ in range(args.steps):
        flush_buf.zero_()                      # L2 flush between timed iterations (outside the events)
        torch.cuda.synchronize()
        partial = handle.run(lo, hi, st)       # CUDA events on the library's stream bracket the launches
        dev_ms += st.kernel_ms
        launches += st.launches
    barrier()
    wall_resident = time.perf_counter() - w0
    t_region1 = time.time()
    path, tile_log2 = st.path, st.tile_log2
    handle.close()
    dev_ms_max = allmax(dev_ms)
    perm_resident = sp.nw_factor(n) * gather_sum(partial)

    # ---- end-to-end arm: host buffers through the C-ABI, K timed steps -----------------------------
    st2 = SpStats()
    for _ in range(args.warmup):
        sp.dense_ryser_range(mat, lo, hi, n, device=dev, stats=st2)
    barrier()
    e0 = time.perf_counter()
    e2e_launches = 0
    part2 = 0.0
    for _ in range(args.steps):
        part2 = sp.dense_ryser_range(mat, lo, hi, n, device=dev, stats=st2)
        e2e_launches += st2.launches
    barrier()
    e2e_s = allmax(time.perf_counter() - e0)
    perm_e2e = sp.nw_factor(n) * gather_sum(part2)
    time.sleep(0.2)
    sampler.stop()
    clocks = sampler.summary(t_region0, t_region1)

    units = args.steps * total
    value = units / (dev_ms_max * 1e-3)
    e2e_value = units / e2e_s
    launches_total = gather_sum(float(launches + e2e_launches))

    # dominant-kernel roofline: one launch of ryser_reg_kernel covers this rank's slice
    per_launch_units = hi - lo
    per_launch_s = (dev_ms / args.steps) * 1e-3
    achieved = per_launch_units * (2 * n + 1) / per_launch_s
    roofline = {
        "bound": "fp64_issue", "achieved": achieved / 1e9, "peak": peak / 1e9, "unit": "Ginstr/s (thread-level FP64)",
        "frac": achieved / peak,
        # dram__bytes_read.sum + dram__bytes_write.sum of one ryser_reg_kernel<36,4,128,4> launch over 2^35
        # indices (ncu --set full, profiles/r01_ncu_dense.txt): 87 296 B read + 0 B written
        "traffic": 87296 if (n == 36 and world == 1) else None,
        "algorithmic": f"{2 * n + 1} FP64 instr per Gray index x {per_launch_units} indices per launch (SURVEY 8(d)); "
                       f"executed count is {2 * n} (the 1.0*x0 multiply is elided)",
        "peak_source": "measured in this run: spd_fp64_peak_instr_per_s (8 independent DFMA chains/thread, best of 3); "
                       "MEASURED_PEAKS.json has no FP64 entry; nominal 148 SM x 64 lanes x 1.965 GHz = 18612 Ginstr/s",
        "kernel": "spb::ryser_reg_kernel<36,4,128,4>" if n == 36 else "spb::ryser_reg_kernel",
        "hbm_note": "working set is n^2+n doubles (10.4 KB) staged in shared memory: DRAM traffic ~0, not HBM-bound",
    }

    line = None
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            rate, cores, sample, kind, _, dt = cpu_reference_rate(mat, n, 12.0)
            cpu = {"value": rate, "unit": "iterations/s", "cores": cores, "kind": kind, "sample": sample,
                   "seconds": dt}
        line = {
            "metric": "gray_code_iterations_per_second", "value": value, "unit": "iterations/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"dense Ryser n={n} density {DENSITY} FP64, one full permanent (2^{n - 1} Gray indices) per step, "
                            f"seeded synthetic matrix (seed {1000 * n}); BASELINE.json configs[3]",
                "partition": f"static, {world} contiguous 2^{ALIGN_LOG2}-aligned slice(s), one per GPU; rank-order host sum of one double per GPU",
                "l2": "256 MiB device memset between timed steps (working set is 10 KB; the kernel is FP64-issue bound)",
                "tile_log2": tile_log2, "kernel_path": path,
            },
            "seconds_per_permanent": dev_ms_max / args.steps * 1e-3,
            "permanent": perm_resident, "permanent_e2e": perm_e2e,
            "permanent_rel_err_vs_long_double_golden": golden_rel_err(n, perm_resident),
            "wall_ms_per_step_resident_incl_flush": 1e3 * wall_resident / args.steps,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "iterations/s", "ms_per_step": 1e3 * e2e_s / args.steps,
                    "h2d_bytes_per_step": 8 * (n * n + n), "d2h_bytes_per_step": 8,
                    "api": "sp_dense_ryser_range(host mat, nov, device, start, end) -> double"},
            "gpu_launches": int(launches_total),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
