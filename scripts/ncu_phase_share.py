#!/usr/bin/env python
"""Instruction / stall-sample share per line range of one source file.
usage: ncu_phase_share.py sass_cuda.csv file.cu name:lo-hi [name:lo-hi ...] [envs]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
target = sys.argv[2]
ranges = [(a.split(":")[0], int(a.split(":")[1].split("-")[0]), int(a.split(":")[1].split("-")[1])) for a in sys.argv[3:] if ":" in a]
envs = [float(a) for a in sys.argv[3:] if ":" not in a]
envs = envs[0] if envs else 1.0
num = lambda v: float(v) if v.replace(".", "").isdigit() else 0.0
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
        agg.append((fname, int(r[0]), num(r[iE]), num(r[iS])))
tE = sum(a[2] for a in agg) or 1.0
tS = sum(a[3] for a in agg) or 1.0
for n, lo, hi in ranges:
    e = sum(x[2] for x in agg if x[0] == target and lo <= x[1] <= hi)
    s = sum(x[3] for x in agg if x[0] == target and lo <= x[1] <= hi)
    print("%-22s inst %5.1f%% (%7.0f/env)  samples %5.1f%%" % (n, 100 * e / tE, e / envs, 100 * s / tS))
e = sum(x[2] for x in agg if x[0] != target)
s = sum(x[3] for x in agg if x[0] != target)
print("%-22s inst %5.1f%% (%7.0f/env)  samples %5.1f%%" % ("other files", 100 * e / tE, e / envs, 100 * s / tS))
