"""Generates the golden fixtures under tests/golden/ from the reference tree.

Run in the build container (where /root/reference and oracle/_ref/libref.so exist):
    python tests/golden/make_golden.py [--big]
Outputs (committed):
    corpus.json   matrices of the reference's own corpus named by BASELINE.json's configs, as
                  `i j val` triples, with
                    - `ld`: the long-double oracle permanent (oracle/oracle.c, pinned by test_oracle.py),
                    - `ref_*`: values returned by the UNMODIFIED reference functions (algo.h) through
                      oracle/_ref/libref.so: perman64 (all-double), parallel_perman64 (float X),
                      parallel_perman64_sparse / parallel_skip_perman64_w after SortOrder / SkipOrder
    small.json    small seeded matrices (n <= 18) with reference outputs of every host-side function
                  (reader, CRS/CCS, SortOrder, SkipOrder) and reference permanents
    grid.json     grid-graph patterns from gridGraph2compressed with Kasteleyn counts
Nothing here is read at test time except the JSON files.
"""
import json, os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import numpy as np
from _oracle import Oracle, Reference

REF = "/root/reference"
O, R = Oracle(), Reference()
BIG = "--big" in sys.argv


def triples(A, typ):
    out = []
    n = A.shape[0]
    for i in range(n):
        for j in range(n):
            if A[i, j] != 0:
                out.append([i, j, int(A[i, j]) if typ == "int" else float(A[i, j])])
    return out


def comp_dict(c):
    return {k: (c[k].tolist() if hasattr(c[k], "tolist") else c[k]) for k in ("cptrs", "rows", "cvals", "rptrs", "cols", "rvals", "nnz")} | {"mat": c["mat"].reshape(-1).tolist()}


def corpus():
    names = ["int/30_0.50_0", "double/30_0.50_0", "int/30_0.20_0", "double/30_0.20_0", "float/30_0.30_0"]
    if BIG:
        names += ["int/32_0.50_0", "double/32_0.50_0", "int/33_0.20_0", "double/33_0.20_0", "int/32_0.20_0"]
    out = {}
    for name in names:
        path = os.path.join(REF, name)
        A, hdr_nnz, typ = R.read_matrix(path)
        n = A.shape[0]
        t = time.time()
        e = {"n": n, "type": typ, "header_nnz": hdr_nnz, "triples": triples(A, typ)}
        e["ld"] = O.perm_ld(A)
        e["ref_perman64"] = R.perman64(A) if n <= 30 else None          # serial, all double (algo.h:1031)
        e["ref_parallel_perman64_float_x"] = R.parallel_perman64(A, 8)   # float X (algo.h:664): wrong on double/
        Ab, _, _ = R.read_matrix(path, generic=False)
        e["ld_binary"] = O.perm_ld(Ab) if n <= 30 or BIG else None
        for pre, key in ((1, "sort"), (2, "skip")):
            c = R.compress(A, pre)
            e["ref_sparse_" + key] = R.sparse(c, threads=8)               # algo.h:568
            e["ref_skipper_" + key] = R.skipper(c, threads=8, balanced=True)  # algo.h:885
            e["colcount_" + key] = np.diff(c["cptrs"]).tolist()
        out[name] = e
        print(name, "ld=%.17g" % e["ld"], "%.1fs" % (time.time() - t), flush=True)
    return out


def small():
    rng = np.random.default_rng(20261018)
    out = []
    for n, p, typ in [(5, 0.6, "int"), (8, 0.5, "int"), (11, 0.4, "double"), (12, 0.3, "int"), (14, 0.5, "double"),
                      (16, 0.25, "int"), (17, 0.35, "double"), (18, 0.5, "int"), (18, 0.2, "int"), (13, 0.45, "float")]:
        while True:
            pat = rng.random((n, n)) < p
            if pat.sum(0).min() > 0 and pat.sum(1).min() > 0:
                break
        if typ == "int":
            A = pat * rng.integers(1, 6, (n, n)).astype(float)
        elif typ == "float":
            A = (pat * np.round(rng.uniform(0.01, 5, (n, n)), 3)).astype(np.float32).astype(float)
        else:
            A = pat * np.round(rng.uniform(0.01, 5, (n, n)), 6)
        e = {"n": n, "type": typ, "triples": triples(A, typ)}
        e["ld"] = O.perm_ld(A)
        e["ref_perman64"] = R.perman64(A)
        e["i128"] = str(O.perm_i128(A.astype(int))) if typ == "int" else None
        for pre in (0, 1, 2):
            c = R.compress(A, pre)
            e["compress_%d" % pre] = comp_dict(c)
            if pre:
                e["ref_sparse_%d" % pre] = R.sparse(c, threads=1)
                e["ref_skipper_%d" % pre] = R.skipper(c, threads=1, balanced=False)
        Ab = (A != 0).astype(float)
        e["ld_binary"] = O.perm_ld(Ab)
        e["i128_binary"] = str(O.perm_i128(Ab.astype(int)))
        out.append(e)
    return out


def grids():
    out = []
    for m, n in [(2, 2), (4, 4), (6, 6), (4, 6), (3, 4), (8, 8), (5, 8), (36, 36)]:
        mat, nnz, cptrs, rows, rptrs, cols = R.grid_graph(m, n)
        out.append({"m": m, "n": n, "nov": int(mat.shape[0]), "nnz": int(nnz), "cptrs": cptrs.tolist(), "rows": rows.tolist(),
                    "rptrs": rptrs.tolist(), "cols": cols.tolist(), "kasteleyn": O.kasteleyn(m, n)})
    return out


def real():
    """the reference's real/ matrices that are small enough for the exact paths: inputs, and the
    long-double permanent where it finishes in minutes (ibm32, n = 32)"""
    out = {}
    for name, want_ld in (("real/ibm32.mtxzero", True), ("real/cage5_c2.mtxzero", False)):
        A, hdr_nnz, typ = R.read_matrix(os.path.join(REF, name))
        e = {"n": int(A.shape[0]), "type": typ, "header_nnz": hdr_nnz, "triples": triples(A, typ)}
        e["ld"] = O.perm_ld(A) if want_ld else None
        for pre, key in ((1, "sort"), (2, "skip")):
            e["colcount_" + key] = np.diff(R.compress(A, pre)["cptrs"]).tolist()
        out[name] = e
        print(name, e["n"], e["ld"], flush=True)
    return out


if __name__ == "__main__":
    if "--real" in sys.argv:
        json.dump(real(), open(os.path.join(HERE, "real.json"), "w"))
        sys.exit(0)
    json.dump(small(), open(os.path.join(HERE, "small.json"), "w"))
    json.dump(grids(), open(os.path.join(HERE, "grid.json"), "w"))
    json.dump(corpus(), open(os.path.join(HERE, "corpus_big.json" if BIG else "corpus.json"), "w"))
    print("done")
