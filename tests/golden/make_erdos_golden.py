"""Generates tests/golden/erdos.json: the twelve 32x32 Erdos-Renyi 0/1 matrices of the reference's
SkipPer experiment kit (revised_perman/sparyser/ErdosRenyi/erdos_32.<p>.<k>.mtx, p = 0.20 ... 0.50,
k = 0, 1, 2) together with the permanents the reference's AUTHORS recorded for them
(revised_perman/sparyser/Results/erdos_n32p<p>a<algo>s<sort>x<k>.single.out, "Overall perman is: ...",
seven algorithm / ordering variants per matrix) -- the only known-answer vectors the reference tree
holds -- and the long-double oracle value (oracle/oracle.c).  Run in the build container:
    python tests/golden/make_erdos_golden.py
Nothing here is read at test time except the JSON file."""
import glob, json, os, re, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import numpy as np
from _oracle import Oracle

KIT = "/root/reference/revised_perman/sparyser"
O = Oracle()
out = {}
for p in ("0.20", "0.30", "0.40", "0.50"):
    for k in (0, 1, 2):
        path = os.path.join(KIT, "ErdosRenyi", "erdos_32.%s.%d.mtx" % (p, k))
        with open(path) as f:
            rows, cols, nnz = (int(x) for x in f.readline().split())
            pairs = [tuple(int(x) - 1 for x in line.split()[:2]) for line in f if line.strip()]
        assert rows == cols == 32 and len(pairs) == nnz
        A = np.zeros((rows, rows))
        for i, j in pairs:
            A[i, j] = 1.0
        recorded = {}
        for log in sorted(glob.glob(os.path.join(KIT, "Results", "erdos_n32p%sa*x%d.single.out" % (p, k)))):
            m = re.search(r"Overall perman is: (\S+)", open(log).read())
            if m:
                recorded[os.path.basename(log)] = m.group(1)
        t0 = time.time()
        ld = O.perm_ld(A)
        name = "erdos_32.%s.%d" % (p, k)
        out[name] = {"n": rows, "type": "int", "triples": [[i, j, 1] for i, j in pairs], "recorded": recorded, "ld": ld}
        print(name, "ld %.17g" % ld, "recorded", sorted(set(recorded.values()))[:3], "%.0fs" % (time.time() - t0), flush=True)
json.dump(out, open(os.path.join(HERE, "erdos.json"), "w"))
print(len(out), "matrices")
