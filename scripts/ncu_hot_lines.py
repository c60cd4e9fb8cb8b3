#!/usr/bin/env python
"""Per-source-line share of executed warp instructions and stall samples from
`ncu -i X.ncu-rep --page source --print-source sass,cuda --csv` (stdin or file argument)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
fname, hdr, agg = "", None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
        iE, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
    elif hdr and r[0].isdigit():
        num = lambda v: float(v) if v.replace(".", "").isdigit() else 0.0
        agg.append((fname, int(r[0]), num(r[iE]), num(r[iS]), r[1].strip()))
tE = sum(a[2] for a in agg) or 1.0
tS = sum(a[3] for a in agg) or 1.0
print("total warp instructions %.0f, samples %.0f" % (tE, tS))
for f, ln, e, s, src in agg:
    if 100 * e / tE >= thresh or 100 * s / tS >= thresh:
        print("%-22s %4d  inst %5.1f%%  samples %5.1f%%  %s" % (f, ln, 100 * e / tE, 100 * s / tS, src[:100]))
