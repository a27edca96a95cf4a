"""Generates tests/golden/corpus_wide.json: a wider slice of the reference's own corpus (int/, float/,
double/; n = 30, 31; densities 0.10 ... 0.90; instance 1) as `i j val` triples with the long-double
oracle permanent `ld` (oracle/oracle.c, pinned by tests/test_oracle.py).  Run in the build container
(where /root/reference exists):  python tests/golden/make_corpus_wide.py
Nothing here is read at test time except the JSON file."""
import json, os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from _oracle import Oracle, Reference
from make_golden import triples  # noqa: E402

REF = "/root/reference"
O, R = Oracle(), Reference()
out = {}
for typ in ("int", "double", "float"):
    for n in (30, 31):
        for p in ("0.10", "0.30", "0.50", "0.70", "0.90"):
            name = "%s/%d_%s_1" % (typ, n, p)
            path = os.path.join(REF, name)
            if not os.path.exists(path):
                continue
            A, hdr_nnz, t = R.read_matrix(path)
            t0 = time.time()
            out[name] = {"n": A.shape[0], "type": t, "header_nnz": hdr_nnz, "triples": triples(A, t), "ld": O.perm_ld(A)}
            print(name, out[name]["ld"], "%.1fs" % (time.time() - t0), flush=True)
json.dump(out, open(os.path.join(HERE, "corpus_wide.json"), "w"))
print(len(out), "files")
