"""Per-kernel digest of `ncu --page source --csv`: total stall samples, samples per opcode class and the
hottest SASS instructions.  usage: ncu_source_hot.py report.ncu-rep [top]"""
import csv, io, subprocess, sys, re
from collections import Counter
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = raw.split('"Kernel Name",')
for b in blocks[1:]:
    lines = b.splitlines()
    name = lines[0].strip('",')
    rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
    tot = 0; byop = Counter(); insts = []; stall = Counter(); exe = Counter()
    for r in rows[1:]:
        if len(r) < len(hdr): continue
        s = int(r[ix["# Samples"]] or 0); tot += s
        op = re.sub(r"^@!?U?P\d+\s+", "", r[ix["Source"]].strip()).split()[0].split(".")[0]
        byop[op] += s; exe[op] += int(r[ix["Instructions Executed"]] or 0)
        insts.append((s, r[ix["Address"]][-5:], r[ix["Source"]].strip()))
        for h in hdr:
            if h.startswith("stall_") and "Not Issued" not in h:
                stall[h] += int(r[ix[h]] or 0)
    print("=" * 100); print(name[:110]); print("samples", tot)
    print("stalls:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / max(tot, 1)) for k, v in stall.most_common(8)))
    print("by opcode (samples%, executed):", ", ".join("%s %.1f%% %d" % (k, 100 * v / max(tot, 1), exe[k]) for k, v in byop.most_common(10)))
    for s, a, src in sorted(insts, reverse=True)[:top]:
        print("  %5.2f%%  %s  %s" % (100 * s / max(tot, 1), a, src[:90]))
