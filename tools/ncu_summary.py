"""Reads a .ncu-rep (ncu -i ... --page raw --csv) and prints the metrics DESIGN.md / VERDICT care about."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__thread_inst_executed_per_inst_executed.pct",
        "sass__inst_executed_register_spilling", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__sass_average_branch_targets_threads_uniform.pct", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg.per_second"]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print("-" * 100)
    for k in keys:
        if k in idx:
            print("%-70s %-12s %s" % (k, units[idx[k]], r[idx[k]]))
    # top warp-stall reasons (warps stalled per issue-active cycle)
    stalls = []
    for h, i in idx.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    if stalls:
        print("%-70s %s" % ("top stalls (warps per issue-active cycle)", ", ".join("%s %.2f" % (n, v) for v, n in stalls[:6])))
    for k in ("smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_lsu.sum",
              "smsp__inst_executed_pipe_xu.sum", "smsp__inst_executed_pipe_fp64.sum", "smsp__warps_eligible.avg.per_cycle_active",
              "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"):
        if k in idx:
            print("%-70s %-12s %s" % (k, units[idx[k]], r[idx[k]]))
