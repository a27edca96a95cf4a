"""CPU-side restatement of the compressed recursion (revised_perman/main.cpp:993-1092) with the
long-double oracle at the leaves: the truth the GPU driver sp_permanent_compressed is compared with.
The reduction steps are the library's host functions (pinned bit-exact against the reference in
tests/test_host_reduce.py); leaf balancing is an independent numpy Sinkhorn."""
import numpy as np


def sinkhorn(a, tol=1e-4, sweeps=2000):
    """(B, rv, cv) with B = diag(rv) a diag(cv) doubly stochastic to `tol`"""
    b = np.array(a, dtype=np.float64)
    n = b.shape[0]
    rv = np.ones(n); cv = np.ones(n)
    for _ in range(sweeps):
        c = b.sum(axis=0); c[c == 0] = 1.0
        b /= c; cv /= c
        r = b.sum(axis=1); r[r == 0] = 1.0
        b /= r[:, None]; rv /= r
        if np.abs(b.sum(axis=0) - 1).max() < tol:
            break
    return b, rv, cv


def leaf_perm(oracle, mat, mode):
    """permanent of a leaf with the long-double oracle; mode: 'none' | 'one' (one sweep) | 'full'"""
    n = mat.shape[0]
    if n == 1:
        return float(mat[0, 0])
    if mode == "none":
        return oracle.perm_ld(mat)
    b, rv, cv = sinkhorn(mat, sweeps=1 if mode == "one" else 2000)
    # without total support the factors drift apart (rv -> huge, cv -> tiny): undo them in long double,
    # whose exponent range cannot overflow on the way
    p = np.longdouble(oracle.perm_ld(b))
    for i in range(n):
        p /= np.longdouble(cv[i])
        p /= np.longdouble(rv[i])
    return float(p)


def oracle_compressed(sp, oracle, a, leaf_nov=14, mode="full", max_leaf=24):
    def total(a):
        m = sp.Matrix.from_dense(a)
        f = m.reduce()
        if f == 0.0:
            return 0.0
        d = m.min_degree()
        if m.nov > leaf_nov and d in (3, 4):
            other = m.split34(d)
            return f * (total(m.mat) + total(other.mat))
        assert m.nov <= max_leaf, "oracle leaf too large"
        return f * leaf_perm(oracle, m.mat, mode)

    return total(a)


def banded(rng, n, weights):
    """rows with 3-4 entries near the diagonal: every reduction step applies somewhere"""
    a = np.zeros((n, n))
    for i in range(n):
        for j in {i, (i + 1) % n, (i + int(rng.integers(2, 5))) % n, (i + n - 1) % n if rng.random() < 0.5 else i}:
            a[i, j] = float(rng.integers(1, 4)) if weights == "int" else round(float(rng.uniform(0.2, 3.0)), 6)
    return a
