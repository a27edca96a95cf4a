"""Generates tests/golden/known_perman.json: the real-world pattern matrices of
revised_perman/elektrik_matrices/known_perman for which the reference's SkipPer kit recorded permanents
(revised_perman/sparyser/RealResults/<name>.a<algo>s<sort>.out, "Overall perman is: ..."), as 0-based
`i j 1` triples (symmetric files expanded) with the recorded values.  Run in the build container:
    python tests/golden/make_known_perman_golden.py
Nothing here is read at test time except the JSON file."""
import glob, json, os, re
HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/revised_perman"
out = {}
for name in ("chesapeake", "will57"):
    with open(os.path.join(REF, "elektrik_matrices", "known_perman", name + ".mtx")) as f:
        banner = f.readline().lower().split()
        symmetric = "symmetric" in banner
        line = f.readline()
        while line.startswith("%"):
            line = f.readline()
        rows, cols, nnz = (int(x) for x in line.split())
        ent = set()
        for line in f:
            if not line.strip():
                continue
            i, j = (int(x) - 1 for x in line.split()[:2])
            ent.add((i, j))
            if symmetric:
                ent.add((j, i))
    recorded = {}
    for log in sorted(glob.glob(os.path.join(REF, "sparyser", "RealResults", name + ".mtx.a*.out"))):
        m = re.search(r"Overall perman is: (\S+) in (\S+)", open(log).read())
        if m:
            recorded[os.path.basename(log)] = {"perman": m.group(1), "seconds": float(m.group(2))}
    out[name] = {"n": rows, "type": "int", "symmetric": symmetric, "file_nnz": nnz,
                 "triples": [[i, j, 1] for i, j in sorted(ent)], "recorded": recorded}
    print(name, rows, cols, nnz, len(ent), recorded)
json.dump(out, open(os.path.join(HERE, "known_perman.json"), "w"))
