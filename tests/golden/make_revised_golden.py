"""Generates tests/golden/revised.json: inputs and outputs of the revised front-end's structural
preprocessing (revised_perman/util.h d1compress / d2compress / d34compress / scalesk), produced by the
UNMODIFIED reference through oracle/_ref/libref_revised.so (built by `make -C oracle` where
/root/reference is present).  Run from the repo root:  python tests/golden/make_revised_golden.py
Python floats round-trip exactly through JSON, so the fixtures are bit-exact."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from _oracle import RevisedReference  # noqa: E402


def pattern(rng, n, deg_lo, deg_hi, weights):
    """n x n non-negative matrix with a full diagonal (perfect matching) and a few entries per row"""
    a = np.zeros((n, n))
    for i in range(n):
        k = int(rng.integers(deg_lo, deg_hi + 1))
        cols = set(rng.choice(n, size=k, replace=False).tolist()) | {i}
        for j in cols:
            a[i, j] = float(rng.integers(1, 6)) if weights == "int" else round(float(rng.uniform(0.1, 5.0)), 6)
    return a


def main():
    rev = RevisedReference()
    rng = np.random.default_rng(20261018)
    cases = []
    while len([c for c in cases if c["op"] == "d1"]) < 6:
        n = int(rng.integers(5, 12))
        a = pattern(rng, n, 0, 2, "int" if len(cases) % 2 else "real")
        if rev.min_nnz(a) != 1:
            continue
        out = rev.d1compress(a)
        cases.append({"op": "d1", "nov": n, "mat": a.reshape(-1).tolist(), "out": out.reshape(-1).tolist()})
    while len([c for c in cases if c["op"] == "d2"]) < 6:
        n = int(rng.integers(5, 12))
        a = pattern(rng, n, 1, 3, "int" if len(cases) % 2 else "real")
        if rev.min_nnz(a) != 2:
            continue
        out = rev.d2compress(a)
        cases.append({"op": "d2", "nov": n, "mat": a.reshape(-1).tolist(), "out": out.reshape(-1).tolist()})
    for deg in (3, 4):
        while len([c for c in cases if c["op"] == "d34" and c["min_deg"] == deg]) < 5:
            n = int(rng.integers(7, 13))
            a = pattern(rng, n, deg - 1, deg + 1, "int" if len(cases) % 2 else "real")
            if rev.min_nnz(a) != deg:
                continue
            first, second = rev.d34compress(a, deg)
            cases.append({"op": "d34", "min_deg": deg, "nov": n, "mat": a.reshape(-1).tolist(),
                          "out": first.reshape(-1).tolist(), "out2": second.reshape(-1).tolist()})
    for thr in (1.0, 2.0, 8.0, 20.0):
        for _ in range(2):
            n = int(rng.integers(6, 14))
            a = pattern(rng, n, 2, 5, "real")
            rv, cv = rev.scalesk(a, thr)
            cases.append({"op": "scalesk", "threshold": thr, "nov": n, "mat": a.reshape(-1).tolist(),
                          "rv": rv.tolist(), "cv": cv.tolist()})
    with open(os.path.join(HERE, "revised.json"), "w") as f:
        json.dump({"source": "revised_perman/util.h via oracle/ref_shim_revised.cpp", "cases": cases}, f)
    print(len(cases), "cases")


if __name__ == "__main__":
    main()
