"""Loaders for the committed fixtures under tests/golden/ (made by tests/golden/make_golden.py)."""
import json
import os

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    p = os.path.join(HERE, name)
    if not os.path.exists(p):
        return None
    with open(p) as f:
        return json.load(f)


def dense_from(entry) -> np.ndarray:
    n = entry["n"]
    a = np.zeros((n, n))
    for i, j, v in entry["triples"]:
        a[i, j] = v
    return a


def write_matrix_file(entry, path):
    """the reference's text format: `nov nnz type` then `i j val` lines (0-based)"""
    with open(path, "w") as f:
        f.write("%d %d %s\n" % (entry["n"], len(entry["triples"]), entry["type"]))
        for i, j, v in entry["triples"]:
            if entry["type"] == "int":
                f.write("%d %d %d\n" % (i, j, v))
            else:
                f.write("%d %d %r\n" % (i, j, v))


def small():
    return _load("small.json")


def grids():
    return _load("grid.json")


def corpus():
    """n=30 corpus entries, plus the n=32/33 ones when corpus_big.json was generated"""
    out = dict(_load("corpus.json") or {})
    out.update(_load("corpus_big.json") or {})
    return out


def corpus_wide():
    """a wider slice of the reference corpus (n = 30, 31; all three types; densities 0.1 ... 0.9)"""
    return _load("corpus_wide.json") or {}


def erdos():
    """the 12 Erdos-Renyi matrices of the reference's SkipPer kit with the authors' recorded permanents"""
    return _load("erdos.json") or {}


def known_perman():
    """real-world pattern matrices (chesapeake 39x39, will57 57x57) with the permanents recorded by the
    reference's SkipPer kit and, when generated, the CPU long-double recursion value"""
    return _load("known_perman.json") or {}


def real():
    return _load("real.json") or {}
