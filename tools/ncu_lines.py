#!/usr/bin/env python
"""Summarise an ncu report per CUDA source line: share of stall samples and of executed instructions.

    ncu -i prof.ncu-rep --page source --print-source cuda,sass --csv > src.csv
    python tools/ncu_lines.py src.csv [min_pct]
"""
import csv
import sys

path = sys.argv[1]
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.reader(open(path)))
cur_file = None
hdr = None
lines = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or r[0] in ("Function Name",) or not r[0].isdigit():
        continue
    ci = {h: i for i, h in enumerate(hdr)}
    # first "Source" col is the CUDA text
    def num(name):
        try:
            return float(r[ci[name]])
        except Exception:
            return 0.0
    stalls = {h[6:]: num(h) for h in hdr if h.startswith("stall_") and "Not Issued" not in h}
    lines.append((cur_file, int(r[0]), r[1].strip(), num("# Samples"), num("Instructions Executed"),
                  num("L1 Wavefronts Shared"), num("L1 Wavefronts Shared Ideal"), stalls))
ts = sum(l[3] for l in lines) or 1
ti = sum(l[4] for l in lines) or 1
print(f"total samples {ts:.0f}  total warp-instructions {ti:.0f}")
agg = {}
for l in lines:
    for k, v in l[7].items():
        agg[k] = agg.get(k, 0) + v
print("stall mix:", ", ".join(f"{k} {100 * v / ts:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for f, ln, src, s, ins, wf, wfi, st in lines:
    if 100 * s / ts >= minpct or 100 * ins / ti >= minpct:
        top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        tops = " ".join(f"{k}:{100 * v / max(s, 1):.0f}%" for k, v in top if v)
        conf = f" smem_wf {wf:.0f}/{wfi:.0f}" if wf else ""
        print(f"{f}:{ln:<4d} samp {100 * s / ts:5.1f}% inst {100 * ins / ti:5.1f}%  [{tops}]{conf}  {src[:90]}")
