#!/usr/bin/env python
"""Bucket executed warp-instructions / stall samples of an ncu cuda,sass source CSV by source line ranges.
usage: ncu_ranges.py src.csv file.cu 100-200:name 201-300:name ..."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
fname = sys.argv[2]
ranges = []
for a in sys.argv[3:]:
    r, n = a.split(":")
    lo, hi = r.split("-")
    ranges.append((int(lo), int(hi), n))
cur = None; hdr = None
tot_i = tot_s = 0.0
acc = {n: [0.0, 0.0] for _, _, n in ranges}
acc["(other files)"] = [0.0, 0.0]; acc["(unbucketed)"] = [0.0, 0.0]
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not r[0].isdigit(): continue
    ci = {h: i for i, h in enumerate(hdr)}
    try: ins = float(r[ci["Instructions Executed"]]); smp = float(r[ci["# Samples"]])
    except Exception: continue
    tot_i += ins; tot_s += smp
    if cur != fname: acc["(other files)"][0] += ins; acc["(other files)"][1] += smp; continue
    ln = int(r[0])
    for lo, hi, n in ranges:
        if lo <= ln <= hi: acc[n][0] += ins; acc[n][1] += smp; break
    else: acc["(unbucketed)"][0] += ins; acc["(unbucketed)"][1] += smp
for n, (i, s) in acc.items():
    print(f"{n:28s} inst {100*i/tot_i:5.1f}%  samples {100*s/tot_s:5.1f}%")
print("total warp-instr", tot_i)
