#!/usr/bin/env python3
"""Per CUDA source line cost of a kernel from an .ncu-rep (needs -lineinfo + --import-source on):
tools/ncu_lines.py report.ncu-rep [min-share]   -> file:line, warp-instr share, stall-sample share, source text"""
import csv, io, subprocess, sys
rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None
rows = []
hdr = None
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; iE = hdr.index("Instructions Executed"); iN = hdr.index("# Samples"); continue
    if hdr and r[0].isdigit() and r[2] in ("", "-"):      # a CUDA source line (aggregated over its SASS)
        try:
            rows.append((cur, int(r[0]), int(r[iE] or 0), int(r[iN] or 0), r[1].strip()))
        except ValueError:
            pass
tot = sum(x[2] for x in rows); totS = sum(x[3] for x in rows)
print("total warp-instr %.4g samples %d" % (tot, totS))
for f, l, e, s, t in rows:
    if e > thr * tot or s > thr * totS:
        print("%-14s %4d  instr %5.2f%%  samples %5.2f%%  | %s" % (f, l, 100.0 * e / max(tot, 1), 100.0 * s / max(totS, 1), t[:110]))
