"""CPU: host logic in C (reader, CRS/CCS, SortOrder, SkipOrder, grid, scheduler partition), the
C-ABI symbol table, and the no-CPU-fallback rule."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import _golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol(sp):
    from superman_b200 import _ffi
    names = set()
    for hdr in ("superman_b200.h", "superman_b200_device.h"):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(spd?_[a-z0-9_]+)\s*\(", text))
        names |= set(re.findall(r"\b(read_calculate_return|matlab_calculate_return_[a-z]+)\s*\(", text))
    assert len(names) > 40
    missing = [n for n in sorted(names) if not hasattr(_ffi.lib, n)]
    assert not missing, missing


def test_reader_and_preprocessing_match_reference_golden(sp, tmp_path):
    for idx, e in enumerate(_golden.small()):
        path = tmp_path / ("m%d.txt" % idx)
        _golden.write_matrix_file(e, path)
        for pre in (0, 1, 2):
            g = e["compress_%d" % pre]
            m = sp.Matrix.read(str(path)).compress(pre)
            assert m.nov == e["n"] and m.type == e["type"] and m.nnz == g["nnz"]
            assert m.mat.reshape(-1).tolist() == g["mat"], (idx, pre)
            assert m.cptrs.tolist() == g["cptrs"] and m.rows.tolist() == g["rows"]
            assert m.rptrs.tolist() == g["rptrs"] and m.cols.tolist() == g["cols"]
            assert m.cvals.tolist() == g["cvals"] and m.rvals.tolist() == g["rvals"]
        b = sp.Matrix.read(str(path), binary=True)
        assert set(np.unique(b.mat)) <= {0.0, 1.0}
        assert (b.mat != 0).sum() == len(e["triples"])


def test_reader_edge_cases(sp, tmp_path):
    p = tmp_path / "edge.txt"
    p.write_text("3 4 int\n0 0 2\n\nnot a line\n1 1 3.9\n2 2 4\n7 7 1\n0 2 5 trailing\n")
    m = sp.Matrix.read(str(p)).compress(0)
    want = np.array([[2, 0, 5], [0, 3, 0], [0, 0, 4]], float)   # "3.9" read as int -> 3; (7,7) out of range ignored
    assert np.array_equal(m.mat, want)
    assert m.header_nnz == 4 and m.nnz == 4
    (tmp_path / "bad.txt").write_text("3 4 complex\n0 0 1\n")
    with pytest.raises(sp.SupermanError):
        sp.Matrix.read(str(tmp_path / "bad.txt"))
    with pytest.raises(sp.SupermanError):
        sp.Matrix.read(str(tmp_path / "missing.txt"))
    (tmp_path / "empty.txt").write_text("")
    with pytest.raises(sp.SupermanError):
        sp.Matrix.read(str(tmp_path / "empty.txt"))
    # float files are rounded to float first (ReadMatrix<float>)
    (tmp_path / "f.txt").write_text("1 1 float\n0 0 0.1\n")
    assert sp.Matrix.read(str(tmp_path / "f.txt")).mat[0, 0] == float(np.float32(0.1))


def test_matrix_market_reader(sp, tmp_path):
    p = tmp_path / "g.mtx"
    p.write_text("%%MatrixMarket matrix coordinate real general\n% a comment\n3 3 4\n1 1 2.5\n2 3 -1\n3 2 4\n1 3 0.5\n")
    m = sp.Matrix.read(str(p)).compress(0)
    assert np.array_equal(m.mat, np.array([[2.5, 0, 0.5], [0, 0, -1], [0, 4, 0]]))
    assert m.type == "double" and m.header_nnz == 4 and m.nnz == 3          # CRS/CCS keep entries > 0 (util.h:537)
    b = sp.Matrix.read(str(p), binary=True)
    assert np.array_equal(b.mat, np.array([[1, 0, 1], [0, 0, 1], [0, 1, 0]], float))
    p = tmp_path / "s.mtx"
    p.write_text("%%MatrixMarket matrix coordinate pattern symmetric\n4 4 4\n1 1\n2 1\n4 3\n3 3\n")
    m = sp.Matrix.read(str(p))
    want = np.zeros((4, 4)); want[0, 0] = want[1, 0] = want[0, 1] = want[3, 2] = want[2, 3] = want[2, 2] = 1
    assert np.array_equal(m.mat, want) and m.type == "int"
    p = tmp_path / "i.mtx"
    p.write_text("%%MatrixMarket matrix coordinate integer general\n2 2 2\n1 2 3\n2 1 4\n")
    assert np.array_equal(sp.Matrix.read(str(p)).mat, np.array([[0, 3], [4, 0]], float))
    for bad in ("%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n",
                "%%MatrixMarket matrix coordinate complex general\n2 2 1\n1 1 1 0\n",
                "%%MatrixMarket matrix coordinate real general\n2 3 1\n1 1 1\n",
                "%%MatrixMarket matrix coordinate real general\n"):
        q = tmp_path / "bad.mtx"
        q.write_text(bad)
        with pytest.raises(sp.SupermanError):
            sp.Matrix.read(str(q))


def test_degree_compression_preserves_the_permanent(sp, oracle):
    rng = np.random.default_rng(11)
    shrunk = 0
    for trial in range(40):
        n = int(rng.integers(4, 15))
        pat = rng.random((n, n)) < rng.choice([0.15, 0.25, 0.4])
        pat[np.arange(n), rng.permutation(n)] = True
        A = pat * rng.integers(1, 4, (n, n)).astype(float)
        want = oracle.perm_ld(A)
        m = sp.Matrix.from_dense(A)
        f = m.reduce()
        k = m.nov
        assert 1 <= k <= n
        shrunk += n - k
        R = m.mat
        got = f * (oracle.perm_ld(R) if k > 1 else R[0, 0])
        assert got == pytest.approx(want, rel=1e-12, abs=1e-9), (trial, n, k)
        if k > 2:   # nothing of degree <= 2 is left
            assert ((R != 0).sum(axis=0) >= 3).all() and ((R != 0).sum(axis=1) >= 3).all()
        m.compress(1)                      # CRS / CCS can be rebuilt on the reduced matrix
        assert m.nnz == int((R > 0).sum())
    assert shrunk > 40
    # empty row -> 0; grid graphs (degrees 2..4) collapse a long way
    Z = np.ones((5, 5)); Z[3, :] = 0
    m = sp.Matrix.from_dense(Z)
    assert m.reduce() == 0.0 and m.nov == 1
    g = sp.Matrix.grid(4, 6)
    f = g.reduce()
    assert f * (oracle.perm_ld(g.mat) if g.nov > 1 else g.mat[0, 0]) == pytest.approx(281.0, rel=1e-12)
    assert g.nov < 12


def test_grid_matches_reference_golden(sp):
    for g in _golden.grids():
        m = sp.Matrix.grid(g["m"], g["n"])
        assert m.nov == g["nov"] and m.nnz == g["nnz"]
        assert m.cptrs.tolist() == g["cptrs"] and m.rows.tolist() == g["rows"]
        assert m.rptrs.tolist() == g["rptrs"] and m.cols.tolist() == g["cols"]
    with pytest.raises(sp.SupermanError):
        sp.Matrix.grid(3, 5)


def test_scheduler_partition(sp):
    from superman_b200 import _ffi
    f = _ffi.lib.sp_sched_boundary
    f.restype = C.c_ulonglong
    f.argtypes = [C.c_ulonglong, C.c_ulonglong, C.c_ulonglong, C.c_ulonglong, C.c_int]
    for lo, hi, parts, align in [(0, 1 << 35, 8, 14), (0, 1 << 35, 3, 14), (1, (1 << 20) + 7, 5, 4), (0, 100, 7, 0), (5, 5, 4, 3)]:
        b = [f(lo, hi, parts, i, align) for i in range(parts + 1)]
        assert b[0] == lo and b[-1] == hi
        assert all(b[i] <= b[i + 1] for i in range(parts))
        for x in b[1:-1]:
            assert x == lo or x % (1 << align) == 0
    # bench.py's rank_slice uses the same rule
    import bench
    for world in (1, 2, 4, 8):
        sl = [bench.rank_slice(1 << 35, r, world) for r in range(world)]
        assert sl[0][0] == 0 and sl[-1][1] == 1 << 35
        assert all(sl[i][1] == sl[i + 1][0] for i in range(world - 1))
        assert [s[0] for s in sl] == [f(0, 1 << 35, world, r, bench.ALIGN_LOG2) for r in range(world)]
    g = _ffi.lib.sp_dynamic_chunks
    g.restype = C.c_ulonglong
    g.argtypes = [C.c_int, C.c_int, C.c_int]
    assert g(36, 29, 1) == 128 and g(40, 29, 8) == 2048      # gpu_exact_dense.cu:786-793
    assert g(33, 30, 1) == 8 and g(30, 29, 8) >= 2


def test_no_cpu_fallback(sp):
    """without a GPU every compute entry point must fail loudly, never compute on the host"""
    if sp.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(sp.SupermanError):
        sp.dense_ryser(np.ones((4, 4)), 4)
    m = sp.Matrix.from_dense(np.ones((4, 4))).compress(0)
    with pytest.raises(sp.SupermanError):
        sp.sparse_ryser(m.mat, m.cptrs, m.rows, m.cvals, 4)
    with pytest.raises(sp.SupermanError):
        sp.rasmussen_sparse(m.rptrs, m.cols, m.cptrs, m.rows, 4, m.nnz, 10)
    with pytest.raises(sp.SupermanError):
        sp.fp64_peak(0, 10)
    # the -o driver too, even for matrices that compress away completely on the host
    for a in (np.eye(6) * 2.0, np.triu(np.ones((7, 7))), np.ones((5, 5))):
        with pytest.raises(sp.SupermanError):
            sp.permanent_compressed(a)


def test_product_does_not_link_the_oracle():
    so = os.path.join(ROOT, "superman_b200", "libsuperman_b200.so")
    out = subprocess.run(["nm", "-D", so], capture_output=True, text=True).stdout
    assert "orc_" not in out
    ldd = subprocess.run(["ldd", so], capture_output=True, text=True).stdout
    assert "liboracle" not in ldd and "libref" not in ldd
    for dirpath, _, files in os.walk(os.path.join(ROOT, "superman_b200")):
        for fn in files:
            if fn.endswith((".py", ".c", ".cu", ".h", ".cuh")):
                assert "oracle" not in open(os.path.join(dirpath, fn)).read().replace("the oracle restates", "").replace("with the C oracle", "").replace("the oracle", ""), fn


def test_cli_argument_errors():
    exe = os.path.join(ROOT, "superman_b200", "perman")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "Option -f is a required argument." in r.stderr      # main.cu:472-475
    r = subprocess.run([exe, "-f", "x", "-p", "-3"], capture_output=True, text=True)
    assert r.returncode == 1 and "requires an argument" in r.stderr                     # main.cu:381-384
    r = subprocess.run([exe, "-f", "x", "-c"], capture_output=True, text=True)
    assert r.returncode == 1 and "CPU-only" in r.stderr
    r = subprocess.run([exe, "--bogus"], capture_output=True, text=True)
    assert r.returncode == 1


def test_wide_corpus_files_read_back_exactly(sp, tmp_path):
    """30 files of the reference corpus (tests/golden/corpus_wide.json) through the reader and the
    three orderings: values, type, nnz and CRS/CCS sizes as the file says"""
    c = _golden.corpus_wide()
    assert len(c) >= 20
    for name, e in sorted(c.items()):
        p = tmp_path / name.replace("/", "_")
        _golden.write_matrix_file(e, p)
        A = _golden.dense_from(e)
        for pre in (0, 1, 2):
            m = sp.Matrix.read(str(p)).compress(pre)
            assert m.nov == e["n"] and m.type == e["type"] and m.header_nnz == e["header_nnz"]
            assert m.nnz == int((A > 0).sum()) == len(m.rows) == len(m.cols)
            if pre == 0:
                assert np.array_equal(m.mat, A)
            else:   # a row / column permutation of the same entries
                assert np.array_equal(np.sort(m.mat, axis=None), np.sort(A, axis=None))


def test_libconnect_exports_the_reference_names():
    """superman_b200/libConnect.so: the file and the four symbols the reference's bindings load
    (interface_connector.c:61-231, superPython.py:6-7)"""
    so = os.path.join(ROOT, "superman_b200", "libConnect.so")
    out = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
    for sym in ("connect", "read_calculate_return", "matlab_calculate_return_int", "matlab_calculate_return_double"):
        assert re.search(r" T %s$" % sym, out, flags=re.M), sym
    lib = C.CDLL(so)               # loads (RTLD_LOCAL) and resolves libsuperman_b200.so through its rpath
    lib.read_calculate_return.restype = C.c_double
    v = lib.read_calculate_return(b"/nonexistent", 5, 1, 1, 1, 1)
    assert v != v                  # NaN: no such file (and no compute without a GPU)


class _LevelImage(C.Structure):
    _fields_ = [("B", C.c_int), ("S0", C.c_int), ("S", C.c_int), ("R", C.c_int),
                ("NC", C.c_int), ("NCP", C.c_int), ("HSP", C.c_int),
                ("colT_hot", C.POINTER(C.c_double)), ("lowR", C.POINTER(C.c_double)), ("dcold", C.POINTER(C.c_double)),
                ("xb_hot", C.POINTER(C.c_double)), ("xb_cold", C.POINTER(C.c_double)), ("cold_start", C.POINTER(C.c_int)),
                ("instr_per_index", C.c_double), ("skip_long_tiles", C.c_int)]


class _LevelPlan(C.Structure):
    _fields_ = [("n", C.c_int), ("skip", C.c_int), ("mat_t", C.POINTER(C.c_double)),
                ("xbase", C.c_double * 64), ("level_sorted", C.c_int * 64), ("img", _LevelImage),
                ("owned", C.c_void_p * 6)]


def _replay(X0, cols):
    """Ryser / Nijenhuis-Wilf sum over all Gray indices for rows with start values X0[r] and column entries
    cols[k][r]: sum_i (-1)^i prod_r (X0[r] + sum_k gray(i)_k cols[k][r])."""
    nk = cols.shape[0]
    i = np.arange(1 << nk, dtype=np.int64)
    g = i ^ (i >> 1)
    bits = ((g[:, None] >> np.arange(nk)[None, :]) & 1).astype(np.float64)        # [2^nk, nk]
    X = X0[None, :] + bits @ cols
    return float(np.sum(np.where(i & 1, -1.0, 1.0) * np.prod(X, axis=1)))


def test_sparse_planner_packs_every_row_exactly_once(sp, oracle):
    """host/sp_level.c on the CPU: whatever column order, row order, engine and slot configuration the planner
    chooses, the packed images it would upload (hot slots, register-cold rows, cold rows; free slots neutral or
    holding a cold row) must describe the same matrix -- replaying the Nijenhuis-Wilf sum from the images gives the
    oracle's permanent.  Also the structural promises the kernel relies on (level order, cold_start, low-column
    image = the first B columns of the hot image, instruction count)."""
    from superman_b200 import _ffi
    build = _ffi.lib.sp_level_plan_build
    build.restype = C.c_int
    build.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_int, C.c_int, C.POINTER(_LevelPlan)]
    free = _ffi.lib.sp_level_plan_free
    free.argtypes = [C.POINTER(_LevelPlan)]
    rng = np.random.default_rng(2024)
    engines = set()
    for trial in range(40):
        n = int(rng.integers(7, 15))                                 # n = 14: the SkipPer tile-length sampling runs too
        dens = float(rng.choice([0.15, 0.25, 0.4, 0.7]))
        A = (rng.random((n, n)) < dens) * rng.integers(1, 5, (n, n)).astype(np.float64)
        A[np.arange(n), rng.permutation(n)] = 1.0
        want = oracle.perm_ld(A)
        dmat_t = np.ascontiguousarray(A.T)                           # dmat_t[k*n + j] = A[j][k]
        xbase = np.ascontiguousarray(A[:, n - 1] - A.sum(axis=1) / 2.0)
        for skip in (0, 1):
            for flags in (0, 1):                                      # 1 = SP_PLAN_WHOLE_SPACE (may reorder columns)
                plan = _LevelPlan()
                rc = build(dmat_t.ctypes.data_as(C.POINTER(C.c_double)), xbase.ctypes.data_as(C.POINTER(C.c_double)),
                           n, skip, flags, C.byref(plan))
                assert rc == 0
                try:
                    lv = list(plan.level_sorted[:n])
                    assert lv == sorted(lv)
                    mt = np.ctypeslib.as_array(plan.mat_t, shape=(n, n)).copy()     # [k][j]
                    xb = np.array(plan.xbase[:n])
                    for j in range(n):                                # level = first flippable column with an entry
                        nz = np.nonzero(mt[:n - 1, j])[0]
                        assert lv[j] == (int(nz[0]) if len(nz) else n)
                    got = _replay(xb, mt[:n - 1, :]) * sp.nw_factor(n)
                    assert got == pytest.approx(want, rel=1e-12), (trial, skip, flags, "row / column order")
                    im = plan.img
                    engines.add(im.B)
                    if im.B == 0:
                        continue
                    B, S0, S, R, NC, NCP, HSP = im.B, im.S0, im.S, im.R, im.NC, im.NCP, im.HSP
                    HS = S0 + (B - 1) * S
                    LB = B + (B & 1)
                    assert HS + R <= HSP and B in (3, 4) and 1 <= S0 <= S
                    hot = np.ctypeslib.as_array(im.colT_hot, shape=(n - 1, HSP))
                    low = np.ctypeslib.as_array(im.lowR, shape=(HS, LB))
                    cold = np.ctypeslib.as_array(im.dcold, shape=(n - 1, NCP))
                    xh = np.ctypeslib.as_array(im.xb_hot, shape=(HSP,))
                    xc = np.ctypeslib.as_array(im.xb_cold, shape=(NCP,))
                    cs = np.ctypeslib.as_array(im.cold_start, shape=(n - B + 2,))
                    assert np.array_equal(low[:, :B], hot[:B, :HS].T)
                    # a slot of level L (or a register-cold / cold row) has no entry in a column it does not expect
                    for s in range(HS):
                        L = 0 if s < S0 else 1 + (s - S0) // S
                        assert not hot[:L, s].any()
                    assert not hot[:B, HS:HS + R].any() and not cold[:B, :NC].any()
                    assert list(cs) == sorted(cs) and cs[0] == 0 and cs[-1] == NC
                    for i in range(n - B + 1):                        # rows [cs[i], ...) have level >= B + i
                        assert not cold[B:B + i, cs[i]:NC].any()
                    got = _replay(np.concatenate([xh[:HS + R], xc[:NC]]),
                                  np.concatenate([hot[:, :HS + R], cold[:, :NC]], axis=1)) * sp.nw_factor(n)
                    assert got == pytest.approx(want, rel=1e-12), (trial, skip, flags, B, S0, S, R)
                    assert 2.0 < im.instr_per_index < 2.0 * n + 2
                finally:
                    free(C.byref(plan))
    assert engines & {3, 4}, engines                                   # the LevelRyser packing was exercised
