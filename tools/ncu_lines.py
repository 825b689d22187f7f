#!/usr/bin/env python
"""Per-source-line instruction and stall-sample shares of one kernel from an ncu report.
usage: ncu_lines.py report.ncu-rep [top]"""
import csv, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
fname, hdr, out = "", None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
        ii, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    elif hdr and r[0].isdigit() and len(r) > ii and r[ii].isdigit():
        out.append((int(r[ii]), int(r[isamp]) if r[isamp].isdigit() else 0, "%s:%s" % (fname, r[0]), r[1].strip()))
tot_i = sum(o[0] for o in out) or 1
tot_s = sum(o[1] for o in out) or 1
print("total warp-instructions %d, samples %d" % (tot_i, tot_s))
out.sort(reverse=True)
for i, s, loc, src in out[:top]:
    print("%5.1f%% instr %5.1f%% samples  %-22s %s" % (100.0 * i / tot_i, 100.0 * s / tot_s, loc, src[:90]))
