"""GPU: the `perman` executable end to end -- flags, id dispatch, output lines (main.cu)."""
import os
import re
import subprocess

import numpy as np
import pytest

import _golden

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "superman_b200", "perman")


def run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([EXE, *args], capture_output=True, text=True, env=e, timeout=600)
    return r


def result_of(out, key="Result17"):
    m = re.search(r"^%s: (\S+) (\S+)" % key, out, flags=re.M)
    assert m, out
    return m.group(1), float(m.group(2))


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp("mats")
    out = []
    for idx, e in enumerate(_golden.small()):
        p = d / ("m%d.txt" % idx)
        _golden.write_matrix_file(e, p)
        out.append((str(p), e))
    return out


def test_dense_ids_and_result_line(files):
    path, e = files[7]          # n = 18 int
    names = {0: "gpu_perman64_xlocal", 1: "gpu_perman64_xlocal", 2: "gpu_perman64_xshared",
             3: "gpu_perman64_xshared_coalescing", 4: "gpu_perman64_xshared_coalescing_mshared",
             5: "gpu_perman64_xshared_coalescing_mshared_multigpu",
             6: "gpu_perman64_xshared_coalescing_mshared_multigpucpu_chunks"}
    for algo, name in names.items():
        r = run("-f", path, "-p", str(algo), "-d", "1", env={"PERMAN_PRECISION": "17"})
        assert r.returncode == 0, r.stderr
        # Result: <name> <6 significant digits> in <seconds>   (main.cu:58)
        m = re.search(r"^Result: (\S+) (\S+) in (\S+)$", r.stdout, flags=re.M)
        assert m and m.group(1) == name
        assert m.group(2) == "%g" % e["ld"]
        assert "kernel" in r.stdout
        assert result_of(r.stdout)[1] == pytest.approx(e["ld"], rel=1e-9)
    r = run("-f", path, "-p", "9")
    assert r.returncode == 0 and "Unknown Algorithm ID" in r.stdout     # main.cu:75
    r = run("-f", path, "-g", "-c", "-p", "4", env={"PERMAN_PRECISION": "17"})   # -c with -g is accepted (main.cu:66)
    assert result_of(r.stdout)[1] == pytest.approx(e["ld"], rel=1e-9)
    r = run("-f", path, "extra_arg", "-p", "4")
    assert "Non-option argument extra_arg" in r.stdout                  # main.cu:477-480


def test_default_algo_and_binary_flag(files):
    path, e = files[3]          # n = 12 int
    r = run("--file", path, env={"PERMAN_PRECISION": "17"})             # defaults: -p1, GPU (main.cu:335,482)
    name, v = result_of(r.stdout)
    assert name == "gpu_perman64_xlocal" and v == pytest.approx(e["ld"], rel=1e-9)
    r = run("-f", path, "-b", "-p", "4", env={"PERMAN_PRECISION": "17"})
    assert round(result_of(r.stdout)[1]) == int(e["i128_binary"])


def test_sparse_ids_with_preprocessing(files):
    for path, e in files[5:9]:
        for pre in ("0", "1", "2"):
            for algo, name in ((4, "gpu_perman64_xshared_coalescing_mshared_sparse"),
                               (6, "gpu_perman64_xshared_coalescing_mshared_multigpucpu_chunks_sparse"),
                               (7, "gpu_perman64_xshared_coalescing_mshared_skipper"),
                               (8, "gpu_perman64_xshared_coalescing_mshared_multigpucpu_chunks_skipper")):
                r = run("-f", path, "-s", "-r", pre, "-p", str(algo), "-d", "1", env={"PERMAN_PRECISION": "17"})
                assert r.returncode == 0, r.stderr
                got_name, v = result_of(r.stdout)
                assert got_name == name
                assert v == pytest.approx(e["ld"], rel=1e-9), (path, pre, algo)


def test_grid_approximations():
    for algo, name in ((1, "gpu_perman64_rasmussen_sparse"), (2, "gpu_perman64_approximation_sparse"),
                       (3, "gpu_perman64_rasmussen_multigpucpu_chunks"),
                       (4, "gpu_perman64_approximation_multigpucpu_chunks_sparse")):
        r = run("-a", "-i", "-m", "6", "-n", "6", "-x", "20000", "-y", "4", "-z", "5", "-p", str(algo), "-d", "1",
                env={"PERMAN_PRECISION": "17"})
        assert r.returncode == 0, r.stderr
        assert re.search(r"^Result: %s \S+ in \S+$" % name, r.stdout, flags=re.M)
        assert re.search(r"^Try: \S+ \S+ in \S+$", r.stdout, flags=re.M)      # main.cu:267
        assert "------------GRID--------------" in r.stdout                   # main.cu:309
        v = result_of(r.stdout)[1]
        se = float(re.search(r"^StdError: \S+ (\S+)", r.stdout, flags=re.M).group(1))
        assert abs(v - 6728.0) < 6 * se
    r = run("-a", "-i", "-m", "3", "-n", "5", "-p", "1")
    assert "one of the grid dimensions should be positive." in r.stdout       # main.cu:405 (sic)


def test_dense_approximation_on_file(files):
    path, e = files[3]
    for algo, name in ((1, "gpu_perman64_rasmussen"), (2, "gpu_perman64_approximation")):
        r = run("-f", path, "-b", "-a", "-p", str(algo), "-x", "50000", env={"PERMAN_PRECISION": "17"})
        assert r.returncode == 0, r.stderr
        got_name, v = result_of(r.stdout)
        assert got_name == name
        assert v == pytest.approx(float(e["i128_binary"]), rel=0.25)


def test_reduce_flag(files):
    for path, e in files[5:9]:
        for extra in ((), ("-s", "-r", "1", "-p", "4"), ("-s", "-r", "2", "-p", "7")):
            r = run("-f", path, "--reduce", *(extra or ("-p", "4")), env={"PERMAN_PRECISION": "17"})
            assert r.returncode == 0, r.stderr
            assert re.search(r"^Compressed: \d+ leaf matrix\(es\), \d+ Gray indices$", r.stdout, flags=re.M)
            assert result_of(r.stdout)[1] == pytest.approx(e["ld"], rel=1e-9)


def test_revised_front_end_flags(files):
    """-k reps / -l device / -o compression / -u scaling / ignored precision flags (revised_perman/main.cpp:1298-1325)"""
    path, e = files[6]
    r = run("-f", path, "-p", "4", "-k", "3", "-l", "0", "-h", "-w", "-q", "-v", "-e", "4", "-u", "2",
            env={"PERMAN_PRECISION": "17"})
    assert r.returncode == 0, r.stderr
    vals = [float(x) for x in re.findall(r"^Result17: \S+ (\S+)", r.stdout, flags=re.M)]
    assert len(vals) == 3 and all(v == pytest.approx(e["ld"], rel=1e-9) for v in vals)
    assert vals[0] == vals[1] == vals[2]                  # bit-reproducible
    r = run("-f", path, "-s", "-p", "4", "-o", env={"PERMAN_PRECISION": "17"})
    assert "Compressed: " in r.stdout and result_of(r.stdout)[1] == pytest.approx(e["ld"], rel=1e-9)
    r = run("-f", path, "-p", "4", "-l", "99")
    assert r.returncode == 1 and "device" in r.stderr


def test_matrixmarket_real_matrix_through_the_cli(tmp_path):
    """will57 (57x57 pattern matrix of revised_perman/elektrik_matrices/known_perman) written back as a
    MatrixMarket file: `perman -s -p4 -r1 -o` reads it, compresses it and prints the permanent the
    library returns for ten different reduction trees (tests/test_gpu_compressed.py)"""
    e = _golden.known_perman()["will57"]
    p = tmp_path / "will57.mtx"
    with open(p, "w") as f:
        f.write("%%MatrixMarket matrix coordinate pattern general\n% written by the test\n")
        f.write("%d %d %d\n" % (e["n"], e["n"], len(e["triples"])))
        for i, j, _ in e["triples"]:
            f.write("%d %d\n" % (i + 1, j + 1))
    r = run("-f", str(p), "-s", "-p", "4", "-r", "1", "-o", env={"PERMAN_PRECISION": "17"})
    assert r.returncode == 0, r.stderr
    assert re.search(r"^Compressed: \d{4,} leaf matrix\(es\)", r.stdout, flags=re.M)
    name, v = result_of(r.stdout)
    assert name == "gpu_perman64_xshared_coalescing_mshared_sparse"
    assert v == pytest.approx(1.070536592880585e18, rel=1e-12)
