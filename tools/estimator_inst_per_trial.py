"""Reads an ncu report of tools/prof_config5.py (BASELINE config 5: both estimators on the 36 x 36 grid, 100 000
trials per launch) and writes profiles/estimator_inst_per_trial.json: executed thread-level instructions per
trial of each kernel (smsp__inst_executed.sum x smsp__thread_inst_executed_per_inst_executed.ratio / trials).
bench.py multiplies them by its measured trials/s to quote the estimators against the measured integer issue
rate.  usage: estimator_inst_per_trial.py <report.ncu-rep> [trials]"""
import csv, json, os, subprocess, sys
rep = sys.argv[1]
trials = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
out = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    key = "rasmussen_p1" if "rasmussen_mid" in name else "scaling_p2_y4_z5" if "approx_kernel" in name else None
    if key is None:
        continue
    inst = float(r[ix["smsp__inst_executed.sum"]])
    ratio = float(r[ix["smsp__thread_inst_executed_per_inst_executed.ratio"]])
    out[key] = {"kernel": name.split("(")[0], "warp_instr_per_trial": inst / trials, "threads_per_instr": ratio,
                "thread_instr_per_trial": inst * ratio / trials,
                "issue_active_pct": float(r[ix["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
                "kernel_ms_under_ncu": float(r[ix["gpu__time_duration.sum"]]) * (1e-6 if "nsecond" in rows[1][ix["gpu__time_duration.sum"]] else 1.0),
                "source": "ncu --set full, %s, %d trials per launch" % (os.path.basename(rep), trials)}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "estimator_inst_per_trial.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
