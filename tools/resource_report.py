"""Registers / stack (spill) bytes / shared memory of every kernel in the shipped library, read back from
the built file with cuobjdump --dump-resource-usage; `make` rewrites profiles/resource_usage.txt with
it after every link, so the committed table cannot go stale.  STACK > 0 means a spill."""
import re, subprocess, sys
lib, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "--dump-resource-usage", lib], capture_output=True, text=True).stdout
pairs = re.findall(r"Function (\S+):\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", txt)
names = subprocess.run(["c++filt"], input="\n".join(p[0] for p in pairs), capture_output=True, text=True).stdout.splitlines()
rows = []
for (m, reg, stack, shared, local), d in zip(pairs, names):
    d = re.sub(r"\(.*\)$", "", d).replace("void spb::", "").replace("spb::", "")
    rows.append((d, int(reg), int(stack), int(shared), int(local)))
def key(r):
    m = re.match(r"(\w+)<(.*)>", r[0])
    return (m.group(1), [int(x) if x.strip().lstrip("-").isdigit() else x for x in m.group(2).split(",")]) if m else (r[0], [])
rows.sort(key=key)
spills = [r for r in rows if r[2] or r[4]]
with open(out, "w") as f:
    f.write("# %s: %d kernels, %d with a stack frame (spill)\n" % (lib, len(rows), len(spills)))
    f.write("# kernel                                                            REG  STACK  SHARED(static)\n")
    for d, reg, stack, shared, local in rows:
        f.write("%-70s %4d %6d %8d\n" % (d, reg, stack + local, shared))
print("%s: %d kernels, %d with a stack frame" % (out, len(rows), len(spills)))
