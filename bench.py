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


def config3_matrix(kind: str, n: int = 33, density: float = 0.2):
    """Seeded n = 33, p = 0.2 matrix of BASELINE config 3 in the three input types the reference corpus has:
    'bin' (the -b flag: every entry 1), 'int' (1..5) and 'dbl' (uniform (0.01, 5), 6 decimals); one entry per
    row and column is forced so that the permanent is not structurally zero."""
    import numpy as np
    rng = np.random.default_rng(1000 * n + {"bin": 1, "int": 2, "dbl": 3}[kind])
    pat = rng.random((n, n)) < density
    pat[np.arange(n), rng.permutation(n)] = True
    if kind == "bin":
        return pat.astype(np.float64)
    if kind == "int":
        return (pat * rng.integers(1, 6, (n, n))).astype(np.float64)
    return (pat * np.round(rng.uniform(0.01, 5.0, (n, n)), 6)).astype(np.float64)


class _StdoutToStderr:
    """The reference's wrappers print `kernel0 in ...` with std::cout; rank 0's stdout must carry exactly one
    JSON line, so file descriptor 1 is pointed at stderr while they run."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def _best_of(fn, st, reps=3, warm=1):
    """(best kernel_ms, best wall_ms, last value) of `reps` timed calls after `warm` untimed ones"""
    v = None
    for _ in range(warm):
        v = fn()
    k, w = [], []
    for _ in range(reps):
        t = time.perf_counter()
        v = fn()
        w.append(1e3 * (time.perf_counter() - t))
        k.append(st.kernel_ms)
    return min(k), min(w), v


def extra_configs(sp, world: int, peak: float, dev: int = 0) -> dict:
    """The other BASELINE.json configs, measured after the headline's timed region on rank 0 (the other ranks
    are parked at the closing barrier): config 2 (n = 32 -p4), config 3 (SpaRyser + SortOrder and SkipPer +
    SkipOrder at n = 33, p = 0.2, per input type), config 4's n = 40, config 5 (both estimators on the 36 x 36
    grid, -x100000 -y4 -z5), the library's own in-process multi-GPU ids 5 / 6 over `world` devices, and the
    unmodified reference GPU wrappers (oracle/_ref/libref_gpu.so) on the same inputs as the kernels to beat."""
    import ctypes as C
    import numpy as np
    from superman_b200._ffi import SpStats
    nominal = 148 * 64 * 1.965e9
    out = {}
    st = SpStats()

    def dense_block(n, algo=4, reps=3, warm=1):
        A = synthetic_matrix(n, DENSITY)
        kms, wms, v = _best_of(lambda: sp.dense_ryser(A, n, algo, stats=st), st, reps, warm)
        its = (1 << (n - 1)) / (kms * 1e-3)
        return {"workload": f"dense Ryser n={n} density {DENSITY} FP64 seeded synthetic, -p{algo}, one full permanent",
                "kernel_ms": kms, "wall_ms": wms, "value": its, "unit": "iterations/s", "permanent": v,
                "rel_err_vs_long_double_golden": golden_rel_err(n, v),
                "roofline": {"bound": "fp64_issue", "frac": its * (2 * n + 1) / peak,
                             "frac_executed_instr": its * (2 * n) / peak, "frac_vs_nominal": its * (2 * n + 1) / nominal,
                             "unit": "fraction of thread-level FP64 instr/s", "peak": peak / 1e9, "nominal": nominal / 1e9}}

    out["config2_n32_p4"] = dense_block(32)
    out["config4_n40_p4"] = dense_block(40, reps=1, warm=0)

    # ---- config 3 ----
    c3 = {}
    for kind in ("bin", "int", "dbl"):
        A = config3_matrix(kind)
        n = A.shape[0]
        for pre, label, skip in ((1, "sparyser_sortorder", False), (2, "skipper_skiporder", True)):
            m = sp.Matrix.from_dense(A).compress(pre)
            if skip:
                fn = lambda: sp.skipper(m.mat, m.rptrs, m.cols, m.cptrs, m.rows, m.cvals, n, 7, stats=st)
            else:
                fn = lambda: sp.sparse_ryser(m.mat, m.cptrs, m.rows, m.cvals, n, 4, stats=st)
            kms, wms, v = _best_of(fn, st)
            eff = (1 << (n - 1)) / (kms * 1e-3)
            # sparse exact paths: the FP64 instructions per index the chosen kernel configuration executes
            # (host/sp_level.c counts them from the packing; ncu's executed counts agree within 3 %)
            model = st.sq_scale
            vf = st.visited / float(1 << (n - 1))
            c3[f"{kind}_{label}"] = {
                "kernel_ms": kms, "wall_ms": wms, "effective_iterations_per_s": eff, "visited": int(st.visited),
                "visited_frac": vf, "permanent": v,
                "fp64_instr_per_index": model,
                # SkipPer: only the evaluated blocks run the hot slots, so this is the share of the FP64 pipe that
                # did useful work, not a distance from a bound
                "roofline_frac": (eff * model * (vf if skip else 1.0) / peak) if model > 0 else None}
    out["config3_n33_p0.2"] = {
        "workload": "seeded 33x33 density 0.2 (bench.config3_matrix), -s -p4 -r1 (SpaRyser + SortOrder) and -s -p7 -r2 "
                    "(SkipPer + SkipOrder); effective it/s = 2^32 / kernel time; roofline_frac = FP64 instructions per "
                    "index of the chosen LevelRyser configuration x evaluated indices / time / measured FP64 issue peak "
                    "(ncu: profiles/r02_ncu_level_engine.txt)", **c3}

    # ---- config 5 ----
    g = sp.Matrix.grid(36, 36)
    int_peak = sp.int_peak(dev, 100)
    ipt = {}
    try:
        with open(os.path.join(ROOT, "profiles", "estimator_inst_per_trial.json")) as f:
            ipt = json.load(f)
    except OSError:
        pass
    c5 = {}
    for label, fn in (("rasmussen_p1", lambda: sp.rasmussen_sparse(g.rptrs, g.cols, g.cptrs, g.rows, g.nov, g.nnz, 100000, 1, seed=0, stats=st)),
                      ("scaling_p2_y4_z5", lambda: sp.scaling_sparse(g.cptrs, g.rows, g.rptrs, g.cols, g.nov, g.nnz, 100000, 4, 5, 1, seed=0, stats=st))):
        kms, wms, v = _best_of(fn, st)
        tps = 100000 / (kms * 1e-3)
        blk = {"kernel_ms": kms, "wall_ms": wms, "value": tps, "unit": "trials/s", "estimate": v, "std_error": st.std_error,
               "survivors": int(st.visited), "trials": int(st.units)}
        per = ipt.get(label)
        if per:
            blk["int_issue_roofline"] = {"thread_instr_per_trial": per["thread_instr_per_trial"], "source": per["source"],
                                         "achieved": tps * per["thread_instr_per_trial"] / 1e9, "peak": int_peak / 1e9,
                                         "unit": "Ginstr/s (thread-level)", "frac": tps * per["thread_instr_per_trial"] / int_peak}
        c5[label] = blk
    out["config5_grid36x36_x100000"] = {"workload": "-a -i -m36 -n36 -x100000 -y4 -z5 (nov 648, nnz 2520), Philox seed 0; exact value "
                                                    "3.0597e159 (Kasteleyn); nearly every trial of either estimator dies on this "
                                                    "pattern, in the reference as here (survivors reported)",
                                        "int_peak_measured_ginstr_s": int_peak / 1e9, **c5}

    # ---- in-process multi-GPU ids over `world` devices (what `perman -p5/-p6 -dN` runs) ----
    gdev = max(1, min(world, sp.device_count()))
    A36 = synthetic_matrix(N_DENSE, DENSITY)
    inproc = {"devices": gdev}
    for algo, label in ((5, "p5_static"), (6, "p6_dynamic")):
        kms, wms, v = _best_of(lambda: sp.dense_ryser(A36, N_DENSE, algo, gpu_num=gdev, stats=st), st)
        inproc[label] = {"kernel_ms_max_over_devices": kms, "wall_ms": wms, "value": (1 << (N_DENSE - 1)) / (wms * 1e-3),
                         "unit": "iterations/s (host wall clock of the library call)", "chunks": st.chunks, "permanent": v,
                         "rel_err_vs_long_double_golden": golden_rel_err(N_DENSE, v)}
    out["in_process_n36"] = inproc

    # ---- the unmodified reference GPU wrappers on the same B200, in a child process: the reference checks no
    # CUDA call, and its sparse kernel faults ("misaligned address") whenever nov + 1 + nnz is odd (doubles
    # placed after an odd number of ints in shared memory, gpu_exact_sparse.cu:467-476) -- a fault must not
    # poison this process's context ----
    so = os.path.join(ROOT, "oracle", "_ref", "libref_gpu.so")
    if os.path.exists(so):
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--reference-gpu-child"], capture_output=True,
                               text=True, timeout=300)
            ref = json.loads(r.stdout.strip().splitlines()[-1])
            ref["ours"] = {"dense_n32_wall_ms": out["config2_n32_p4"]["wall_ms"],
                           "sparyser_kernel_ms": {k: c3[f"{k}_sparyser_sortorder"]["kernel_ms"] for k in ("bin", "int", "dbl")},
                           "rasmussen_grid36x36_ms_per_2^20_trials": c5["rasmussen_p1"]["kernel_ms"] * (1 << 20) / 100000}
            out["reference_gpu"] = ref
        except Exception as e:   # a reporting extra must not take the headline down
            out["reference_gpu"] = {"error": str(e)[:200]}
    return out


def reference_gpu_child():
    """Times the unmodified reference GPU wrappers (oracle/_ref/libref_gpu.so); prints one JSON line."""
    import ctypes as C
    import numpy as np
    import superman_b200 as sp      # host-side reader / orderings only: nothing of ours runs on the GPU here
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_gpu.so"))
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    lib.ref_gpu_dense_multigpu.restype = C.c_double
    lib.ref_gpu_dense_multigpu.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.ref_gpu_sparse_multigpu.restype = C.c_double
    lib.ref_gpu_sparse_multigpu.argtypes = [dp, ip, ip, dp, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.ref_gpu_rasmussen_chunks_sparse.restype = C.c_double
    lib.ref_gpu_rasmussen_chunks_sparse.argtypes = [ip, ip, ip, ip, C.c_int, C.c_int, C.c_int, C.c_int]
    ref = {"note": "unmodified reference .cu files compiled for sm_100a (oracle/Makefile), launch geometry as RunAlgo "
                   "(2048 x 128 / 256); wall seconds of the wrapper call, second of two calls; the reference keeps X in "
                   "float, so its values on double-valued input are only ~3 digits (SURVEY 8(c))"}

    def timed(fn):
        with _StdoutToStderr():
            fn()
            t = time.perf_counter()
            v = fn()
            dt = time.perf_counter() - t
        return dt, v
    for n in (32, 36):
        A = np.ascontiguousarray(synthetic_matrix(n, DENSITY))
        s_, v = timed(lambda: lib.ref_gpu_dense_multigpu(A.ctypes.data_as(dp), n, 1, 2048, 128))
        ref[f"dense_n{n}"] = {"wall_s": s_, "value": v, "iterations_per_s": (1 << (n - 1)) / s_}
    g = sp.Matrix.grid(36, 36)
    s_, v = timed(lambda: lib.ref_gpu_rasmussen_chunks_sparse(g.cptrs.ctypes.data_as(ip), g.rows.ctypes.data_as(ip), g.rptrs.ctypes.data_as(ip),
                                                              g.cols.ctypes.data_as(ip), g.nov, g.nnz, 100000, 1))
    ref["rasmussen_grid36x36"] = {"wall_s": s_, "value": v, "trials": 1 << 20,
                                  "note": "the reference runs 1024 x 1024 trials per launch whatever -x says"}
    sparse = {}
    for kind in ("bin", "int", "dbl"):
        m = sp.Matrix.from_dense(config3_matrix(kind)).compress(1)
        if (33 + 1 + m.nnz) % 2:
            sparse[kind] = {"skipped": f"nov + 1 + nnz = {34 + m.nnz} is odd: the reference kernel faults with a misaligned "
                                       "shared-memory address on this input"}
            continue
        mat = np.ascontiguousarray(m.mat)
        s_, v = timed(lambda: lib.ref_gpu_sparse_multigpu(mat.ctypes.data_as(dp), m.cptrs.ctypes.data_as(ip), m.rows.ctypes.data_as(ip),
                                                          m.cvals.ctypes.data_as(dp), 33, 1, 2048, 256))
        sparse[kind] = {"wall_s": s_, "value": v}
    ref["sparyser_n33_sortorder"] = sparse
    print(json.dumps(ref), flush=True)
    return 0


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
                   "sample_per_step": sample,
                   "rate_note": "rate extrapolated from a slice: one step = the sample above, not a whole 2^35-index permanent "
                                "(the CPU would need ~50 s per permanent); iterations/s is independent of the slice"},
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
    ap.add_argument("--no-configs", action="store_true", help="skip the block with the other BASELINE configs")
    ap.add_argument("--reference-gpu-child", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.reference_gpu_child:
        return reference_gpu_child()
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
        # host-side (gloo) group for the closing barrier: while rank 0 measures the in-process multi-GPU ids on
        # every device, the other ranks must wait on the CPU -- an NCCL barrier would leave a spinning kernel on
        # their GPUs, which the in-process run shares
        park_group = dist.new_group(backend="gloo")

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
        # the same fraction on the 2n instructions the kernel executes per index, and against the nominal peak
        "frac_executed_instr": achieved / peak * (2 * n) / (2 * n + 1),
        "frac_vs_nominal": achieved / (148 * 64 * 1.965e9),
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
        if not args.no_configs:
            sampler2 = ClockSampler(dev)
            sampler2.start()
            c0 = time.time()
            try:
                line["configs"] = extra_configs(sp, world, peak, dev)
            except Exception as e:     # the headline line is printed whatever happens to the extras
                line["configs"] = {"error": str(e)[:300]}
            c1 = time.time()
            time.sleep(0.15)
            sampler2.stop()
            line["configs"]["clocks"] = sampler2.summary(c0, c1)
            line["configs"]["seconds"] = c1 - c0
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier(group=park_group)
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
