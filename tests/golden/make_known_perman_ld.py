"""Adds `ld_recursion` to tests/golden/known_perman.json: the permanent of each matrix from the CPU
restatement of the compressed recursion (tests/_compressed.py: the library's pinned host reduction steps,
numpy Sinkhorn, long-double oracle at leaves of order <= 23).  CPU only, long: ~20 min for chesapeake
(12 822 leaves), ~80 min for will57 (48 975 leaves).  Usage:
    python tests/golden/make_known_perman_ld.py chesapeake [will57]"""
import json, os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import superman_b200 as sp
import _golden
from _oracle import Oracle
from _compressed import oracle_compressed

path = os.path.join(HERE, "known_perman.json")
O = Oracle()
for name in sys.argv[1:]:
    d = json.load(open(path))
    a = _golden.dense_from(d[name])
    t0 = time.time()
    v = oracle_compressed(sp, O, a, leaf_nov=22, mode="full", max_leaf=24)
    d = json.load(open(path))
    d[name]["ld_recursion"] = v
    d[name]["ld_recursion_leaf_nov"] = 22
    json.dump(d, open(path, "w"))
    print(name, "%.17g" % v, "%.0f s" % (time.time() - t0), flush=True)
