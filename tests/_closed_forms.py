"""Matrices whose permanent is known in closed form, for known-answer tests at orders no CPU oracle can
reach (n = 40: 2^39 Gray indices).  Every generator returns (A, exact) with `exact` a Python int or
Fraction computed in exact arithmetic from the same (dyadic, exactly representable) parameters the
float64 matrix is built from, so that A's permanent is `exact` to the last bit of A's entries.

  derangement_matrix   D1 (J - I) D2          perm = D(n) * prod(d1) * prod(d2),  D(n) = n-th derangement number
  rank1_plus_diag      u v^T + diag(d)        perm = sum_k (n-k)! * [t^k] prod_i (d_i t + u_i v_i)
                                              (expand over the set S of rows that take their diagonal entry d_i;
                                              the rest is the permanent of a rank-1 block, (n-|S|)! prod u_i v_i)
  block_diagonal       P (B_1 + ... + B_m) Q  perm = prod perm(B_i), each B_i small enough for the exact
                                              __int128 Ryser oracle; P, Q random permutations hide the blocks
"""
from __future__ import annotations

from fractions import Fraction
from math import factorial

import numpy as np


def dyadic(rng, lo, hi, bits, size):
    """random multiples of 2^-bits in [lo, hi): exactly representable, products of two stay exact"""
    q = 1 << bits
    return rng.integers(int(lo * q), int(hi * q), size).astype(np.float64) / q


def derangements(n: int) -> int:
    a, b = 1, 0           # D(0), D(1)
    for k in range(2, n + 1):
        a, b = b, (k - 1) * (a + b)
    return b if n >= 1 else 1


def derangement_matrix(rng, n, scaled=True):
    d1 = dyadic(rng, 0.5, 2.0, 6, n) if scaled else np.ones(n)
    d2 = dyadic(rng, 0.5, 2.0, 6, n) if scaled else np.ones(n)
    A = np.outer(d1, d2) * (1.0 - np.eye(n))
    exact = Fraction(derangements(n))
    for x in list(d1) + list(d2):
        exact *= Fraction(float(x))
    return A, exact


def rank1_plus_diag(rng, n):
    u = dyadic(rng, 0.5, 1.5, 6, n)
    v = dyadic(rng, 0.5, 1.5, 6, n)
    d = dyadic(rng, -1.0, 3.0, 6, n)
    A = np.outer(u, v) + np.diag(d)
    poly = [Fraction(1)]                               # coefficients of prod_i (d_i t + w_i), lowest first
    for i in range(n):
        di, wi = Fraction(float(d[i])), Fraction(float(u[i])) * Fraction(float(v[i]))
        nxt = [Fraction(0)] * (len(poly) + 1)
        for k, ck in enumerate(poly):
            nxt[k] += ck * wi
            nxt[k + 1] += ck * di
        poly = nxt
    exact = sum(ck * factorial(n - k) for k, ck in enumerate(poly))
    return A, exact


def block_diagonal(rng, oracle, sizes, kind="int", density=0.6, shuffle=True):
    n = sum(sizes)
    A = np.zeros((n, n))
    exact = 1
    o = 0
    for m in sizes:
        while True:
            pat = rng.random((m, m)) < density
            pat[np.arange(m), rng.permutation(m)] = True
            B = pat.astype(np.int64) if kind == "bin" else pat * rng.integers(1, 4, (m, m))
            p = oracle.perm_i128(B.astype(int))
            if p != 0:
                break
        A[o:o + m, o:o + m] = B
        exact *= int(p)
        o += m
    if shuffle:
        A = A[rng.permutation(n)][:, rng.permutation(n)]
    return np.ascontiguousarray(A, dtype=np.float64), exact


def perm_banded_exact(A, w):
    """exact permanent (Python ints) of an integer matrix whose row i only has entries in columns
    [i - w, i + w]: rows in order, state = set of used columns inside the sliding window"""
    n = len(A)
    states = {0: 1}                      # bit b of the key <-> column (i - w + b) of the current row i
    for i in range(n):
        nxt = {}
        for used, cnt in states.items():
            for b in range(2 * w + 1):
                c = i - w + b
                if c < 0 or c >= n or (used >> b) & 1 or A[i][c] == 0:
                    continue
                u2 = used | (1 << b)
                if not (u2 & 1) and i - w >= 0:
                    continue             # column i - w leaves the window unused: no later row can take it
                key = u2 >> 1
                nxt[key] = nxt.get(key, 0) + cnt * int(A[i][c])
        states = nxt
    return sum(states.values())       # n rows took n distinct columns and none was dropped unused: all are used


def dense_row_over_band(n, active):
    """n x n 0/1 matrix: row 0 all ones (more than 255 entries for n >= 257), rows 1 .. active-1 tridiagonal,
    the other rows identity.  perm = sum_j perm(minor(0, j)), every minor banded: exact by perm_banded_exact."""
    A = np.zeros((n, n), dtype=np.int64)
    A[0, :] = 1
    for i in range(1, active):
        for c in (i - 1, i, i + 1):
            if c < active:
                A[i, c] = 1
    for i in range(active, n):
        A[i, i] = 1
    exact = 0
    for j in range(active):              # row 0 on an identity column leaves that identity row without a column
        M = np.delete(np.delete(A[:active, :active], 0, axis=0), j, axis=1)
        exact += perm_banded_exact(M.tolist(), 2)
    return A.astype(np.float64), exact
