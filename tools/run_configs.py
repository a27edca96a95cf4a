"""Runs BASELINE.json's five configs through the `perman` executable (the drop-in surface) on this
box and prints what it printed, with wall times.  Inputs: the reference corpus matrices stored in
tests/golden/*.json (written back to the reference's text format) and bench.py's seeded synthetic
n = 36 / 40 matrices.  Evidence only (profiles/): nothing here is a test."""
import json, os, re, subprocess, sys, tempfile, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import bench, _golden
import superman_b200 as sp

EXE = os.path.join(R, "superman_b200", "perman")
tmp = tempfile.mkdtemp()
ngpu = sp.device_count()


def write_dense(A, path, typ="double"):
    n = A.shape[0]
    nz = [(i, j, A[i, j]) for i in range(n) for j in range(n) if A[i, j] != 0]
    with open(path, "w") as f:
        f.write("%d %d %s\n" % (n, len(nz), typ))
        for i, j, v in nz:
            f.write("%d %d %r\n" % (i, j, float(v)) if typ != "int" else "%d %d %d\n" % (i, j, int(v)))


def run(label, *args, want=None):
    t = time.perf_counter()
    r = subprocess.run([EXE, *args], capture_output=True, text=True, env=dict(os.environ, PERMAN_PRECISION="17"))
    dt = time.perf_counter() - t
    out = [l for l in r.stdout.splitlines() if not l.startswith("0 ") and not l.startswith("1 ") and "GRID" not in l]
    print("## %s\n$ perman %s      [process wall %.3f s, rc %d]" % (label, " ".join(args).replace(tmp, "<tmp>"), dt, r.returncode))
    for l in out[:14]:
        print("   " + l)
    if r.stderr.strip():
        print("   stderr: " + r.stderr.strip()[:300])
    m = re.search(r"^Result17: \S+ (\S+)", r.stdout, flags=re.M)
    if want is not None and m:
        print("   -> relative difference to the long-double oracle value %.17g: %.2e" % (want, abs(float(m.group(1)) / want - 1)))
    print(flush=True)


c = _golden.corpus()
files = {}
for name, e in c.items():
    p = os.path.join(tmp, name.replace("/", "_"))
    _golden.write_matrix_file(e, p)
    files[name] = (p, e)

print("# devices visible: %d\n" % ngpu)
print("# config 1 (CPU -c) is the reference's algo.h path: not provided by this build (no CPU fallback)")
run("config 1 refusal", "-f", files["double/30_0.50_0"][0], "-c")
print("# config 2: dense Ryser n=32 density 0.50 FP64, -p4")
for name in ("double/32_0.50_0", "int/32_0.50_0"):
    if name in files:
        run(name, "-f", files[name][0], "-g", "-p4", want=files[name][1]["ld"])
print("# config 3: SpaRyser n=33 density 0.20 + SortOrder (-s -p4 -r1), SkipPer + SkipOrder (-s -p7 -r2)")
for name in ("int/33_0.20_0", "double/33_0.20_0"):
    if name in files:
        run(name + " SpaRyser+SortOrder", "-f", files[name][0], "-s", "-p4", "-r1", want=files[name][1]["ld"])
        run(name + " SkipPer+SkipOrder", "-f", files[name][0], "-s", "-p7", "-r2", want=files[name][1]["ld"])
        if name.startswith("int/"):
            run(name + " -b SkipPer+SkipOrder", "-f", files[name][0], "-b", "-s", "-p7", "-r2", want=files[name][1].get("ld_binary"))
print("# config 4: dense Ryser n=36/40 synthetic FP64, static (-p5) and dynamic (-p6) across 1/2/4/8 GPUs")
g36 = None
p = os.path.join(R, "tests", "golden", "bench36.json")
if os.path.exists(p):
    g36 = json.load(open(p))["ld"]
for n in (36, 40):
    path = os.path.join(tmp, "synthetic_%d_0.50" % n)
    write_dense(bench.synthetic_matrix(n, 0.5), path)
    for g in [x for x in (1, 2, 4, 8) if x <= ngpu]:
        if n == 40 and g not in (1, ngpu):
            continue
        for algo in ("-p5", "-p6"):
            if n == 40 and g == 1 and algo == "-p6":
                continue
            run("n=%d %s -d%d" % (n, algo, g), "-f", path, algo, "-d", str(g), want=g36 if n == 36 else None)
print("# config 5: Rasmussen / scaling on the 36x36 grid graph, -a -i -m36 -n36 -x100000 -y4 -z5")
for algo in ("-p1", "-p2", "-p3", "-p4"):
    run("grid 36x36 " + algo, "-a", "-i", "-m", "36", "-n", "36", "-x", "100000", "-y", "4", "-z", "5", algo, "-d", str(ngpu))
print("# (for a pattern where the estimators survive: 8x8 grid, exact 12988816)")
for algo in ("-p1", "-p2"):
    run("grid 8x8 " + algo, "-a", "-i", "-m", "8", "-n", "8", "-x", "1000000", "-y", "4", "-z", "5", algo)
