"""Long-double oracle value of the HEADLINE workload (bench.py's seeded synthetic n=36 matrix).
Takes ~1.5 h on 8 cores; output: tests/golden/bench36.json (value + the matrix seed it belongs to)."""
import json, os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import bench
from _oracle import Oracle
n = int(sys.argv[1]) if len(sys.argv) > 1 else 36
A = bench.synthetic_matrix(n, bench.DENSITY)
t = time.time()
v = Oracle().perm_ld(A)
json.dump({"n": n, "density": bench.DENSITY, "seed": 1000 * n, "ld": v, "seconds": time.time() - t,
           "checksum": float(A.sum())}, open(os.path.join(HERE, "bench%d.json" % n), "w"))
print(n, "%.17g" % v, time.time() - t)
