#!/usr/bin/env python
"""per-source-line warp-stall samples of one kernel from an .ncu-rep (needs -lineinfo + --import-source on)
    python scripts/ncu_source_lines.py <rep> <kernel regex> [top N]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg, tot, instr = {}, 0, {}
hdr = None
for r in rows:
    if len(r) > 6 and r[0] == "Line No":
        hdr = r
        si, ii = r.index("# Samples"), r.index("Instructions Executed")
        continue
    if hdr is None or len(r) <= si or r[0] == "":
        continue  # sass rows repeat under their source line
    try:
        n, ne = int(r[si]), int(r[ii])
    except ValueError:
        continue
    key = (int(r[0]), r[1].strip()[:120])
    agg[key] = agg.get(key, 0) + n
    instr[key] = instr.get(key, 0) + ne
    tot += n
print("total samples", tot)
for (ln, src), n in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    print("%6d %5.1f%%  inst %8d  L%-5d %s" % (n, 100.0 * n / max(tot, 1), instr[(ln, src)], ln, src))
